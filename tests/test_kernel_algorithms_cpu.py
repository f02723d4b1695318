"""CPU: numpy restatements of two device algorithms, operation for operation, against their
references -- so that the arithmetic the kernels are written to perform is pinned even where no
GPU is present.  (The kernels themselves are compared with the same references in the -m gpu
tests: test_device_time_tables_bit_exact, test_latent_pca_matches_sklearn_golden.)"""
import importlib

import numpy as np

PKG = "autoencoded-vocal-analysis_b200"


def _emulate_time_tables(grid0, K, base, tstart, tstop, n_t):
    """csrc/spec.cu::time_tables_kernel: every operation a separately rounded float64 op."""
    step = (tstop - tstart) / np.float64(n_t - 1)
    j = np.arange(n_t, dtype=np.float64)
    T = (j[None, :] * step[:, None]) + tstart[:, None]
    T[:, -1] = tstop
    dt = base[1] - base[0]
    g0 = grid0[:, None]
    first, last = base[0] + g0, base[K - 1][:, None] + g0
    km2 = (K - 2)[:, None]
    f = np.minimum(np.maximum(np.floor((T - first) / dt), 0.0), km2.astype(np.float64))
    i = f.astype(np.int64)
    for _ in range(2):
        i = np.where(base[i] + g0 > T, i - 1, i)
        i = np.clip(i, 0, km2)
        i = np.where((base[i + 1] + g0 <= T) & (i < km2), i + 1, i)
    lo, hi = base[i] + g0, base[i + 1] + g0
    w = (T - lo) / (hi - lo)
    bad = (T < first) | (T > last)
    return np.where(bad, -1, i).astype(np.int32), np.where(bad, 0.0, w), T


def test_device_time_table_arithmetic_equals_host_tables():
    """The device builds target times as fl(fl(j*step) + start) with the last one = stop, and
    brackets them against fl(base[k] + grid0): identical, bit for bit, to np.linspace + the
    host `bracket` (which follows ava/models/window_vae_dataset.py:231-235 and the interp2d fill
    rule) -- including windows at the file start, past the file end, and too-short segments."""
    pre = importlib.import_module(PKG + ".preprocessing.utils")
    fs, nperseg, hop, wl, sh = 32000, 512, 256, 0.12, 0.05
    rng = np.random.default_rng(2)
    for lo_t, hi_t, flen_s in ((0.0, 600.0, 600.05), (9.7, 10.05, 10.0), (0.0, 0.06, 3.0)):
        n = 1500
        onsets = rng.uniform(lo_t, hi_t, n)
        offsets = onsets + wl
        t1, t2 = np.maximum(0, onsets - sh), offsets + sh
        s1, s2 = np.rint(t1 * fs).astype(np.int64), np.rint(t2 * fs).astype(np.int64)
        flen = np.full(n, int(flen_s * fs))
        seg = np.minimum(flen, s2) - np.maximum(0, s1)
        short = (seg < nperseg) | (s1 >= flen)
        seg = np.where(short, 0, seg)
        K = np.where(short, 3, pre.num_frames(np.maximum(seg, nperseg), nperseg, hop))
        kmax = int(K.max())
        base = np.arange(nperseg / 2, nperseg / 2 + kmax * hop, hop) / float(fs) - (nperseg / 2) / fs
        tt = np.linspace(onsets, offsets, 128, axis=-1)
        want_i, want_w = pre.bracket(np.maximum(0.0, t1), base, K, tt)
        got_i, got_w, T = _emulate_time_tables(np.maximum(0.0, t1), K, base, onsets, offsets, 128)
        assert np.array_equal(T, tt)
        assert np.array_equal(got_i, want_i) and np.array_equal(got_w, want_w)


def _round_robin_pairs(D, r):
    n = D + (D & 1)
    out = [(r % (n - 1), n - 1)]
    for i in range(1, n // 2):
        out.append(((r + i) % (n - 1), (r - i + (n - 1)) % (n - 1)))
    return [(min(p, q), max(p, q)) for p, q in out if max(p, q) < D]


def _jacobi_round_robin(C, max_sweeps=40):
    """csrc/pca.cu::pca_eigh_kernel: the D/2 disjoint pairs of a round rotate together (columns,
    then rows, then the closed-form 2x2 blocks); entries negligible against both diagonals are
    zeroed instead of rotated after the first sweeps."""
    D = C.shape[0]
    A, V = 0.5 * (C + C.T), np.eye(D)
    n = D + (D & 1)
    sweeps = 0
    for sweep in range(max_sweeps):
        off = ((A - np.diag(np.diag(A))) ** 2).sum()
        if off == 0.0 or off <= 1e-30 * (np.diag(A) ** 2).sum():
            break
        sweeps += 1
        for r in range(n - 1):
            rots = []
            for p, q in _round_robin_pairs(D, r):
                apq, app, aqq = A[p, q], A[p, p], A[q, q]
                g = 100.0 * abs(apq)
                if sweep > 3 and abs(app) + g == abs(app) and abs(aqq) + g == abs(aqq):
                    rots.append((p, q, 1.0, 0.0, 0.0, app, aqq, apq))
                elif apq != 0.0:
                    tau = (aqq - app) / (2.0 * apq)
                    t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1.0 + tau * tau))
                    c = 1.0 / np.sqrt(1.0 + t * t)
                    rots.append((p, q, c, t * c, t, app, aqq, apq))
            for p, q, c, s, *_ in rots:
                if s != 0.0:
                    a, b = A[:, p].copy(), A[:, q].copy()
                    A[:, p], A[:, q] = c * a - s * b, s * a + c * b
                    a, b = V[:, p].copy(), V[:, q].copy()
                    V[:, p], V[:, q] = c * a - s * b, s * a + c * b
            for p, q, c, s, *_ in rots:
                if s != 0.0:
                    a, b = A[p, :].copy(), A[q, :].copy()
                    A[p, :], A[q, :] = c * a - s * b, s * a + c * b
            for p, q, c, s, t, app, aqq, apq in rots:
                A[p, p], A[q, q] = app - t * apq, aqq + t * apq
                A[p, q] = A[q, p] = 0.0
    return np.diag(A).copy(), V, sweeps


def test_round_robin_jacobi_matches_lapack():
    from oracle import pca_oracle
    for D, seed in ((32, 5), (8, 7), (64, 8), (31, 9), (2, 1), (1, 0)):
        x = pca_oracle.synth_latents(seed, 800, D)
        C = np.atleast_2d(np.cov(x.T))
        seen = set()
        n = D + (D & 1)
        for r in range(n - 1):                     # every pair exactly once per sweep
            for pq in _round_robin_pairs(D, r):
                assert pq not in seen
                seen.add(pq)
        assert len(seen) == D * (D - 1) // 2
        lam, V, sweeps = _jacobi_round_robin(C)
        w = np.linalg.eigvalsh(C)
        assert np.abs(np.sort(lam) - w).max() <= 1e-12 * max(abs(w).max(), 1e-300)
        assert np.abs(C @ V - V * lam).max() <= 1e-12 * max(abs(w).max(), 1e-300)
        assert np.abs(V.T @ V - np.eye(D)).max() <= 1e-13
        assert sweeps <= 10
    # degenerate inputs: zero matrix (converged at once), rank one
    lam, V, sweeps = _jacobi_round_robin(np.zeros((32, 32)))
    assert sweeps == 0 and not lam.any()
    v = np.arange(32.0)
    lam, V, _ = _jacobi_round_robin(np.outer(v, v))
    assert abs(lam.max() - (v ** 2).sum()) <= 1e-9 * (v ** 2).sum()
