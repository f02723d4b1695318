"""SURVEY section 8(f) 'next' rows, to the same bar as the hot path: the syllable
preprocessing driver (N2: process_sylls on the batched GPU get_spec) and the MMD^2
estimators on latent means (N4), each checked against outputs of the unmodified reference
(tests/golden/process_sylls.npz, mmd_cases.npz; oracle/make_golden.py)."""
import importlib
import os

import numpy as np
import pytest

from oracle import make_golden, mmd_oracle
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
PKG = "autoencoded-vocal-analysis_b200"


def _read_written(path):
    if path.endswith(".npz"):
        return dict(np.load(path))
    import h5py
    with h5py.File(path, "r") as f:
        return {k: np.array(f[k]) for k in f}


def test_process_sylls_matches_reference(tmp_path):
    pp = importlib.import_module(PKG + ".preprocessing.preprocess")
    pre = importlib.import_module(PKG + ".preprocessing.utils")
    g = load_golden("process_sylls")
    adir, sdir = make_golden.write_process_corpus(str(tmp_path))
    p = dict(make_golden.PROCESS_P)
    p['get_spec'] = pre.get_spec
    written = pp.process_sylls(adir, sdir, str(tmp_path / "out"), p, shuffle=True, verbose=False)
    ref_files = sorted({k.split(":")[0] for k in g.files if ":" in k})
    assert [os.path.basename(w)[:14] for w in written] == [f[:14] for f in ref_files]
    for w, rf in zip(written, ref_files):
        d = _read_written(w)
        np.testing.assert_array_equal(d['onsets'], g[rf + ":onsets"])
        np.testing.assert_array_equal(d['offsets'], g[rf + ":offsets"])
        names = [os.path.basename(i.decode()) for i in d['audio_filenames']]
        assert names == [i.decode() for i in g[rf + ":audio_filenames"]]
        assert d['specs'].dtype == np.float64 and d['specs'].shape == g[rf + ":specs"].shape
        assert np.abs(d['specs'] - g[rf + ":specs"]).max() <= 1e-5      # north_star: spectrograms 1e-5
    # the short segment (< nperseg samples) is an all-zero, still "valid" spectrogram
    allspecs = np.concatenate([_read_written(w)['specs'] for w in written])
    assert any(not s.any() for s in allspecs)
    # plugin contract: a user-supplied p['get_spec'] is called per syllable
    calls = []

    def plugin(t1, t2, audio, p, fs, target_freqs=None):
        calls.append((t1, t2))
        return np.full((p['num_freq_bins'], p['num_time_bins']), 0.5), len(calls) % 2 == 0
    p2 = dict(p)
    p2['get_spec'] = plugin
    specs, valid = pp.get_syll_specs([0.1, 0.3, 0.5], [0.2, 0.4, 0.6], os.path.join(adir, "a.wav"), p2)
    assert len(calls) == 3 and valid == [1] and len(specs) == 1


def test_mmd_matches_reference():
    mmd = importlib.import_module(PKG + ".plotting.mmd_plots")
    g = load_golden("mmd_cases")
    latent, condition = make_golden.mmd_latent()
    # bandwidth: host RNG in the reference's order + numpy-ordered device sums -> bit-exact
    assert mmd.estimate_median_sigma(latent) == float(g["sigma"])
    assert mmd.estimate_median_sigma(latent, n=500, seed=7) == float(g["sigma_n500_seed7"])
    sigma = float(g["sigma"])
    conds = np.unique(condition)
    groups = [np.argwhere(condition == c).flatten() for c in conds]
    for i in range(2):
        for j in range(i + 1, 3):
            got = mmd._estimate_mmd2(latent, groups[i], groups[j], sigma=sigma)
            assert abs(got - g["mmd2"][i, j]) <= 1e-10 * abs(g["mmd2"][i, j])
            got = mmd._estimate_mmd2_linear_time(latent, groups[i], groups[j], sigma=sigma)
            assert abs(got - g["mmd2_linear"][i, j]) <= 1e-9 * max(abs(g["mmd2_linear"][i, j]), 1e-3)
            got = mmd._estimate_mmd2(latent, groups[i].copy(), groups[j].copy(), sigma=sigma, max_n=64,
                                     seed=5)
            assert abs(got - g["mmd2_max64_seed5"][i, j]) <= 1e-10 * abs(g["mmd2_max64_seed5"][i, j])
    # all pairs of conditions in one pass over the Gram matrix
    m, allc = mmd.mmd2_matrix(latent, condition, sigma=sigma)
    assert list(allc) == list(conds)
    assert np.abs(m - g["mmd2"]).max() <= 1e-10
    m2, _ = mmd.mmd2_matrix(latent, condition)          # sigma from the median heuristic
    assert np.abs(m2 - g["mmd2"]).max() <= 1e-10
    lin, _ = mmd.mmd2_matrix(latent, condition, alg='linear', sigma=sigma)
    assert np.abs(lin - g["mmd2_linear"]).max() <= 1e-9
    # larger, ragged problem vs the numpy oracle (tile boundaries inside and between conditions)
    rng = np.random.default_rng(3)
    big = rng.standard_normal((1000, 32)) * 0.7
    cond = rng.integers(0, 5, size=1000)
    mb, cb = mmd.mmd2_matrix(big, cond, sigma=3.0)
    for i in range(4):
        for j in range(i + 1, 5):
            want = mmd_oracle.estimate_mmd2(big, np.argwhere(cond == cb[i]).flatten(),
                                            np.argwhere(cond == cb[j]).flatten(), 3.0)
            assert abs(mb[i, j] - want) <= 1e-9 * max(abs(want), 1e-6)
