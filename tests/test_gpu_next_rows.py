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


# ---------------------------------------------------------------- N1: DataContainer + PCA
@pytest.mark.parametrize("name", ["big", "small", "z8", "z64"])
def test_latent_pca_matches_sklearn_golden(name):
    """csrc/pca.cu against scikit-learn's PCA as the reference constructs it
    (ava/data/data_container.py:543; tests/golden/pca_cases.npz).  fp64: 1e-9."""
    from oracle import pca_oracle
    dcm = importlib.import_module(PKG + ".data.data_container")
    g = load_golden("pca_cases")
    seed, n, d = (int(v) for v in g[name + ":shape"])
    x = pca_oracle.synth_latents(seed, n, d)
    pca = dcm.LatentPCA(n_components=2)
    emb = pca.fit_transform(x)
    ref = g[name + ":embedding"]
    assert emb.dtype == np.float64 and emb.shape == ref.shape
    assert np.abs(emb - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(pca.mean_ - g[name + ":mean"]).max() <= 1e-12
    assert np.abs(pca.components_ - g[name + ":components"]).max() <= 1e-9
    assert np.abs(pca.explained_variance_ / g[name + ":explained_variance"] - 1).max() <= 1e-9
    assert np.abs(pca.explained_variance_ratio_ / g[name + ":explained_variance_ratio"] - 1).max() <= 1e-9
    # bitwise reproducible (fixed-order reductions), and more components only append columns
    np.testing.assert_array_equal(dcm.LatentPCA(2).fit_transform(x), emb)
    full = dcm.LatentPCA(n_components=d)
    emb_full = full.fit_transform(x)
    np.testing.assert_array_equal(emb_full[:, :2], emb)
    o_emb, _ = pca_oracle.pca_fit_transform(x, d)
    lead = min(d, 6)       # trailing directions of the synthetic cloud are nearly degenerate
    assert np.abs(emb_full[:, :lead] - o_emb[:, :lead]).max() <= 1e-7 * np.abs(o_emb).max()
    assert abs(full.explained_variance_ratio_.sum() - 1.0) <= 1e-12


def test_latent_pca_float32_device_input_and_edges():
    """The encoder's fp32 latents go in without a host round trip; ragged sizes around the
    64-row tile; degenerate inputs."""
    import torch
    from oracle import pca_oracle
    dcm = importlib.import_module(PKG + ".data.data_container")
    for n in (2, 3, 63, 64, 65, 129, 1000, 20000):
        x32 = pca_oracle.synth_latents(40 + n, n, 32).astype(np.float32)
        want, fit = pca_oracle.pca_fit_transform(x32.astype(np.float64), 2)
        pca = dcm.LatentPCA(2)
        got = pca.fit_transform_device(torch.from_numpy(x32).cuda())
        assert got.is_cuda and got.dtype == torch.float64
        scale = max(np.abs(want).max(), 1e-30)
        if n > 3:
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-9 * scale, n
        else:   # rank-deficient: compare the variances, directions are not unique
            assert np.abs(pca.explained_variance_ - fit["explained_variance"]).max() <= 1e-9 * scale
    # a large off-centre cloud: the one-pass shifted Gram matrix must not cancel
    rng = np.random.default_rng(9)
    x = rng.standard_normal((300000, 32)) * np.linspace(2.0, 0.1, 32) + 1000.0
    want, fit = pca_oracle.pca_fit_transform(x, 2)
    pca = dcm.LatentPCA(2)
    got = pca.fit_transform(x)
    assert np.abs(got - want).max() <= 1e-8 * np.abs(want).max()
    assert np.abs(pca.mean_ - fit["mean"]).max() <= 1e-10
    # embedding columns are centred and uncorrelated with variances = explained_variance_
    c = np.cov(got.T)
    assert np.abs(got.mean(0)).max() <= 1e-9
    assert abs(c[0, 1]) <= 1e-9 * c[0, 0]
    assert np.abs(np.diag(c) / pca.explained_variance_ - 1).max() <= 1e-9
    # constant input: zero variance, zero embedding
    z = dcm.LatentPCA(2).fit_transform(np.full((100, 32), 3.5))
    assert np.abs(z).max() == 0.0
    with pytest.raises(ValueError):
        dcm.LatentPCA(40).fit_transform(np.zeros((10, 32)))
    with pytest.raises(RuntimeError):
        dcm.LatentPCA(2).fit_transform(np.zeros((1, 32)))      # fewer than 2 rows
    with pytest.raises(RuntimeError):
        dcm.LatentPCA(2).fit_transform(np.zeros((10, 65)))     # more than 64 dimensions


def test_data_container_latent_means_and_pca_match_reference(tmp_path):
    """DataContainer.request('latent_means') / ('latent_mean_pca') against the reference's
    DataContainer run unmodified on the same corpus and checkpoint
    (tests/golden/container_case.npz; ava/data/data_container.py:435-487, 538-551)."""
    from oracle import vae_oracle
    from tests.helpers import rel_err
    dcm = importlib.import_module(PKG + ".data.data_container")
    mu = importlib.import_module(PKG + ".models.utils")
    vae_mod = importlib.import_module(PKG + ".models.vae")
    g = load_golden("container_case")
    try:
        import h5py  # noqa: F401
        ext = ".hdf5"
    except ImportError:
        ext = ".npz"
    nd = len(make_golden.CONTAINER_FILES)
    spec_dirs = [str(tmp_path / ("specs%d" % d)) for d in range(nd)]
    proj_dirs = [str(tmp_path / ("proj%d" % d)) for d in range(nd)]
    for sd, files in zip(spec_dirs, make_golden.container_corpus()):
        os.makedirs(sd)
        for base, specs in files:
            mu.append_field(os.path.join(sd, base + ext), 'specs', specs)
    model = vae_mod.VAE(save_dir=str(tmp_path))
    model.load_flat_state(vae_oracle.make_params(make_golden.CONTAINER_SEED))
    model.save_state("checkpoint_000.tar")
    dc = dcm.DataContainer(spec_dirs=spec_dirs, projection_dirs=proj_dirs,
                           model_filename=str(tmp_path / "checkpoint_000.tar"), verbose=False)
    emb = dc.request('latent_mean_pca')        # makes latent_means first, like the reference
    latent = dc.request('latent_means')        # read back from the projection files
    assert latent.dtype == np.float64 and latent.shape == g["latent_means"].shape
    # forward latents: rtol 1e-4 (north_star); train-mode BN on batches of 64 as the reference
    assert rel_err(latent, g["latent_means"]) <= 1e-4
    # per-file layout
    for d, nf in enumerate(make_golden.CONTAINER_FILES):
        assert sorted(os.listdir(proj_dirs[d])) == ["syllables_%04d%s" % (j, ext) for j in range(nf)]
        for j in range(nf):
            fn = os.path.join(proj_dirs[d], "syllables_%04d%s" % (j, ext))
            assert sorted(mu.stored_fields(fn)) == ["latent_mean_pca", "latent_means"]
            ref = g["proj%d/syllables_%04d.hdf5:latent_means" % (d, j)]
            assert rel_err(mu.read_field(fn, 'latent_means'), ref) <= 1e-4
    # PCA of the reference's own latents: fp64 parity
    pca_ref = dcm.LatentPCA(2).fit_transform(g["latent_means"])
    assert np.abs(pca_ref - g["latent_mean_pca"]).max() <= 1e-9 * np.abs(g["latent_mean_pca"]).max()
    # end to end (fp32 latent differences pass through the eigenvectors)
    assert emb.shape == g["latent_mean_pca"].shape
    assert rel_err(emb, g["latent_mean_pca"]) <= 2e-3
    np.testing.assert_array_equal(dc.request('latent_mean_pca'), emb)
    # a second container finds the stored projections; eval-mode / large-batch variant runs
    dc2 = dcm.DataContainer(spec_dirs=spec_dirs, projection_dirs=proj_dirs, verbose=False)
    assert {'latent_means', 'latent_mean_pca'} <= set(dc2.fields)
    np.testing.assert_array_equal(dc2.request('latent_means'), latent)
    dc2.clear_projections()
    dc3 = dcm.DataContainer(spec_dirs=spec_dirs, projection_dirs=proj_dirs, verbose=False,
                            model_filename=str(tmp_path / "checkpoint_000.tar"),
                            latent_batch_size=1024, latent_eval=True)
    lat_eval = dc3.request('latent_means')
    want = vae_oracle.encode(vae_oracle.make_params(make_golden.CONTAINER_SEED),
                             vae_oracle.make_input(make_golden.CONTAINER_SEED, len(lat_eval)),
                             train=False)[0].numpy()
    assert rel_err(lat_eval, want) <= 1e-4
