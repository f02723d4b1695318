"""
Maximum mean discrepancy between sets of latent means on the GPU -- the computational half of
ava/plotting/mmd_plots.py (the matplotlib / t-SNE plotting half is out of scope).

Same functions and arguments as the reference:

    estimate_median_sigma(latent, n=10000, seed=42)            mmd_plots.py:450-476
    _estimate_mmd2(latent, i1, i2, sigma, max_n, seed)         mmd_plots.py:255-295
    _estimate_mmd2_linear_time(latent, i1, i2, sigma)          mmd_plots.py:298-312
    _calculate_mmd2(dc, condition_from_fn, ...)                mmd_plots.py:337-434

plus ``mmd2_matrix(latent, condition, ...)``, the array-level core of `_calculate_mmd2`: every
pair of conditions in ONE pass over the Gram matrix (`ava_b200_mmd_block_sums`) instead of an
O(n^2) Python double loop per pair.  All random choices (`np.random.seed/shuffle/randint`) stay
on the host in the reference's order, so subsampling and the bandwidth pairs are bit-identical;
arithmetic is float64 on the device.  No CPU fallback.
"""
import os

import numpy as np
import torch

from .. import _lib
from .._lib import call

EPSILON = 1e-8   # mmd_plots.py:34


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("ava_b200 mmd_plots requires a CUDA device; no CPU fallback")
    _lib.lib()
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _latent_dev(latent):
    if torch.is_tensor(latent):
        return latent.to(device=_dev(), dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(latent, dtype=np.float64)).to(_dev())


def _pairs(x, ia, ib, A=0.0, mode=0):
    """out[k] = ||x[ia[k]] - x[ib[k]]||^2 (mode 0) or exp(A * that) (mode 1), float64 on device."""
    ia = torch.from_numpy(np.ascontiguousarray(ia, dtype=np.int64)).to(x.device)
    ib = torch.from_numpy(np.ascontiguousarray(ib, dtype=np.int64)).to(x.device)
    out = torch.empty(ia.numel(), dtype=torch.float64, device=x.device)
    call("ava_b200_pair_kernel", x.data_ptr(), x.shape[1], ia.data_ptr(), ib.data_ptr(), ia.numel(),
         float(A), int(mode), out.data_ptr(), _stream())
    return out


def estimate_median_sigma(latent, n=10000, seed=42):
    """Median pairwise distance of `n` random pairs, for use as a kernel bandwidth
    (mmd_plots.py:450-476).  Pair indices are drawn on the host in the reference's order
    (i1, i2 alternately from np.random.randint), distances on the device."""
    N = len(latent)
    np.random.seed(seed)
    draws = np.array([np.random.randint(N) for _ in range(2 * n)], dtype=np.int64)
    np.random.seed(None)
    x = _latent_dev(latent)
    arr = _pairs(x, draws[0::2], draws[1::2]).cpu().numpy()
    return np.sqrt(np.median(arr) + EPSILON)


def _block_sums(x, seg, n_seg, A):
    seg_d = torch.from_numpy(np.ascontiguousarray(seg, dtype=np.int32)).to(x.device)
    S = torch.empty(n_seg, n_seg, dtype=torch.float64, device=x.device)
    call("ava_b200_mmd_block_sums", x.data_ptr(), x.shape[0], x.shape[1], seg_d.data_ptr(), n_seg,
         float(A), S.data_ptr(), _stream())
    return S.cpu().numpy()


def _mmd2_from_sums(S, counts, a, b):
    n1, n2 = counts[a], counts[b]
    term_1 = (S[a, a] - n1) / (n1 * (n1 - 1))      # 2/(n(n-1)) * sum_{i<j} k  (k_ii = 1)
    term_2 = (S[b, b] - n2) / (n2 * (n2 - 1))
    term_3 = 2.0 * S[a, b] / (n1 * n2)
    return term_1 + term_2 - term_3


def _estimate_mmd2(latent, i1, i2, sigma=None, max_n=None, seed=None):
    """Quadratic-time unbiased MMD^2 estimate between latent[i1] and latent[i2]
    (Gretton et al. 2012; mmd_plots.py:255-295, same arguments, `i1`/`i2` are shuffled in place
    when subsampled, as in the reference)."""
    if sigma is None:
        sigma = estimate_median_sigma(latent)
    A = -0.5 / (sigma ** 2)
    n1, n2 = len(i1), len(i2)
    if max_n is not None:
        np.random.seed(seed)
        n1, n2 = min(max_n, n1), min(max_n, n2)
        if n1 < len(i1):
            np.random.shuffle(i1)
            i1 = i1[:n1]
        if n2 < len(i2):
            np.random.shuffle(i2)
            i2 = i2[:n2]
        np.random.seed(None)
    x = _latent_dev(latent)
    idx = torch.from_numpy(np.concatenate([np.asarray(i1, dtype=np.int64),
                                           np.asarray(i2, dtype=np.int64)])).to(x.device)
    sub = x.index_select(0, idx).contiguous()
    seg = np.concatenate([np.zeros(n1, np.int32), np.ones(n2, np.int32)])
    S = _block_sums(sub, seg, 2, A)
    return _mmd2_from_sums(S, [n1, n2], 0, 1)


def _estimate_mmd2_linear_time(latent, i1, i2, sigma=None):
    """Linear-time MMD^2 estimate (mmd_plots.py:298-312)."""
    if sigma is None:
        sigma = estimate_median_sigma(latent)
    A = -0.5 / (sigma ** 2)
    n = min(len(i1), len(i2))
    m = n // 2
    assert m > 0
    i1, i2 = np.asarray(i1, dtype=np.int64), np.asarray(i2, dtype=np.int64)
    x1, y1 = i1[0:2 * m:2], i2[0:2 * m:2]
    x2, y2 = i1[1:2 * m:2], i2[1:2 * m:2]
    x = _latent_dev(latent)
    k = _pairs(x, np.concatenate([x1, y1, x1, x2]), np.concatenate([x2, y2, y2, y1]), A, 1)
    k = k.view(4, m)
    h = k[0] + k[1] - k[2] - k[3]
    return float(h.sum().item()) / m


def mmd2_matrix(latent, condition, alg='quadratic', max_n=None, sigma=None):
    """MMD^2 between every pair of conditions: the array-level core of `_calculate_mmd2`
    (mmd_plots.py:381-420).  Returns ``(mmd2 [n,n], all_conditions)``."""
    assert alg in ['linear', 'quadratic']
    condition = np.asarray(condition)
    all_conditions = np.unique(condition)
    n = len(all_conditions)
    result = np.zeros((n, n))
    if sigma is None:
        sigma = estimate_median_sigma(latent)
    groups = [np.argwhere(condition == c).flatten() for c in all_conditions]
    if alg == 'linear':
        for i in range(n - 1):
            for j in range(i + 1, n):
                result[i, j] = result[j, i] = _estimate_mmd2_linear_time(latent, groups[i], groups[j],
                                                                         sigma=sigma)
        return result, all_conditions
    if max_n is not None and any(len(g) > max_n for g in groups):
        # the reference reshuffles a condition's indices for every pair it takes part in
        # (seed=None): pair by pair, same draw order
        for i in range(n - 1):
            for j in range(i + 1, n):
                i1 = np.argwhere(condition == all_conditions[i]).flatten()
                i2 = np.argwhere(condition == all_conditions[j]).flatten()
                result[i, j] = result[j, i] = _estimate_mmd2(latent, i1, i2, sigma=sigma, max_n=max_n)
        return result, all_conditions
    # one pass over the Gram matrix for all pairs of conditions
    A = -0.5 / (sigma ** 2)
    x = _latent_dev(latent)
    order = np.concatenate(groups) if n else np.zeros(0, np.int64)
    seg = np.concatenate([np.full(len(g), k, np.int32) for k, g in enumerate(groups)]) if n else \
        np.zeros(0, np.int32)
    sub = x.index_select(0, torch.from_numpy(order.astype(np.int64)).to(x.device)).contiguous()
    S = _block_sums(sub, seg, max(n, 1), A)
    counts = [len(g) for g in groups]
    for i in range(n - 1):
        for j in range(i + 1, n):
            result[i, j] = result[j, i] = _mmd2_from_sums(S, counts, i, j)
    return result, all_conditions


def _calculate_mmd2(dc, condition_from_fn, mmd2_fn=None, condition_fn=None, parallel=False,
                    alg='quadratic', max_n=None, sigma=None, verbose=True):
    """Helper function for calculating MMD^2 from a DataContainer (mmd_plots.py:337-434; same
    arguments; `parallel` is accepted and ignored -- the GPU pass replaces the joblib pool)."""
    assert alg in ['linear', 'quadratic']
    assert mmd2_fn is not None
    if verbose:
        print("Estimating an MMD matrix...")
        print("\talg:", alg)
        print("\tmax_n:", max_n)
    latent = dc.request('latent_means')
    audio_fns = dc.request('audio_filenames')
    condition = np.array([condition_from_fn(str(i)) for i in audio_fns], dtype='int')
    if sigma is None:
        sigma = estimate_median_sigma(latent)
    if verbose:
        print("\tconditions found:", len(np.unique(condition)))
        print("\tsigma:", sigma)
    result, all_conditions = mmd2_matrix(latent, condition, alg=alg, max_n=max_n, sigma=sigma)
    if mmd2_fn is not None:
        if verbose:
            print("\tSaving MMD^2 to:", mmd2_fn)
        np.save(mmd2_fn, result)
    if condition_fn is not None:
        if verbose:
            print("\tSaving conditions to:", condition_fn)
        np.save(condition_fn, all_conditions)
    if verbose:
        print("\tDone.")
    return result, all_conditions


def _mmd2_to_mmd(mmd2):
    """mmd_plots.py:488-: clip negative estimates, take the square root."""
    return np.sqrt(np.clip(mmd2, 0.0, None))
