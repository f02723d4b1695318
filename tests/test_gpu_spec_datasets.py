"""GPU parity tests of the get_spec front end and the datasets against goldens generated
by the reference's own get_spec / FixedWindowDataset (oracle/make_golden.py) and against the
numpy oracle (oracle/spec_oracle.py).  Tolerance: 1e-5 absolute (north_star)."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import spec_oracle
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
PKG = "autoencoded-vocal-analysis_b200"
TOL = 1e-5


@pytest.fixture(scope="module")
def pre():
    return importlib.import_module(PKG + ".preprocessing.utils")


@pytest.fixture(scope="module")
def win():
    return importlib.import_module(PKG + ".models.window_vae_dataset")


def test_get_spec_matches_reference_golden(pre):
    g = load_golden("spec_cases")
    p = dict(spec_oracle.MOUSE_P)
    fs = p['fs']
    audio = spec_oracle.synth_audio(11, int(0.6 * fs), fs)
    for name in ["mouse_a", "mouse_b", "mouse_full", "mouse_neg", "mouse_end", "mouse_short"]:
        t1, t2 = g[name + "_t"]
        spec, flag = pre.get_spec(t1, t2, audio, p, fs=fs)
        assert flag is True and spec.shape == (128, 128) and spec.dtype == np.float64
        assert np.abs(spec - g[name]).max() <= TOL, name
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    audio2 = spec_oracle.synth_audio(12, int(4.0 * fs), fs)
    for name in ["finch_a", "finch_b", "finch_c", "finch_d"]:
        onset = float(g[name + "_t"][0])
        offset = onset + p['window_length']
        tt = np.linspace(onset, offset, 128)
        spec, _ = pre.get_spec(max(0.0, onset - 0.05), offset + 0.05, audio2, p, fs=fs,
                               target_times=tt)
        assert np.abs(spec - g[name]).max() <= TOL, name
    p3 = dict(spec_oracle.FINCH_P)
    p3.update(mel=False, noverlap=384, max_dur=0.3, time_stretch=True)
    audio3 = (spec_oracle.synth_audio(13, int(2.0 * fs), fs, dtype=np.float64) / 3.0).astype(np.float32)
    spec, _ = pre.get_spec(0.5, 0.7, audio3, p3, fs=fs, remove_dc_offset=False)
    assert np.abs(spec - g["f32_a"]).max() <= TOL


def test_batched_engine_matches_oracle(pre):
    """A ragged batch through SpecEngine (device-resident audio, one launch) against the
    numpy oracle, including a too-short segment and out-of-range target times."""
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    audio = [spec_oracle.synth_audio(20 + i, int((1.0 + 0.5 * i) * fs), fs) for i in range(3)]
    eng = pre.SpecEngine(audio, fs, p)
    rng = np.random.default_rng(3)
    files = rng.integers(0, 3, size=37)
    onsets = rng.uniform(0.0, 0.85, size=37)
    onsets[5] = 0.0
    wl = p['window_length']
    tt = np.linspace(onsets, onsets + wl, 128, axis=-1)
    t1, t2 = np.maximum(0.0, onsets - 0.05), onsets + wl + 0.05
    t2[9] = t1[9] + 0.004          # shorter than nperseg -> zeros
    out, out64 = eng.specs(files, t1, t2, tt, want_float64=True)
    torch.cuda.synchronize()
    out, out64 = out.cpu().numpy(), out64.cpu().numpy()
    for i in range(37):
        ref, _ = spec_oracle.get_spec(t1[i], t2[i], audio[files[i]], p, fs=fs, target_times=tt[i])
        assert np.abs(out64[i] - ref).max() <= 1e-9, i
        assert np.abs(out[i] - ref).max() <= TOL, i
    assert not out[9].any()


def test_device_time_tables_bit_exact(pre):
    """ava_b200_window_time_tables (target times + bracketing built on the device for the
    fixed-window sampler) reproduces the host float64 tables bit for bit, so specs_linspace
    and specs give identical spectrograms -- windows at the file start, running off the file
    end (out-of-range targets -> fill) and too-short segments included."""
    lib = importlib.import_module(PKG + "._lib")
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    audio = [spec_oracle.synth_audio(30 + i, int((2.0 + 0.7 * i) * fs), fs) for i in range(3)]
    eng = pre.SpecEngine(audio, fs, p)
    rng = np.random.default_rng(11)
    wl = p['window_length']
    for n in (1, 2, 37, 512, 1023):
        files = rng.integers(0, 3, size=n)
        onsets = rng.uniform(0.0, 2.1, size=n)          # file 0 is 2.0 s long: some run off the end
        onsets[0] = 0.0
        t1, t2 = np.maximum(0.0, onsets - 0.05), onsets + wl + 0.05
        if n > 9:
            t2[9] = t1[9] + 0.004                       # shorter than nperseg -> zeros
        tt = np.linspace(onsets, onsets + wl, 128, axis=-1).reshape(n, 128)
        a32, a64 = eng.specs(files, t1, t2, tt, want_float64=True)
        b32, b64 = eng.specs_linspace(files, t1, t2, onsets, onsets + wl, want_float64=True)
        assert torch.equal(a32, b32) and torch.equal(a64, b64), n
        # and the tables themselves
        seg_start, seg_len, K, kmax, base = eng._segments(files, t1, t2)
        t_idx, t_frac = pre.bracket(np.maximum(0.0, t1), base, K, tt)
        dev = eng.device
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)       # noqa: E731
        g0, Kd, bd, ts, te = d(np.maximum(0.0, t1)), d(K.astype(np.int32)), d(base), d(onsets), d(onsets + wl)
        ti = torch.empty(n, 128, dtype=torch.int32, device=dev)
        tf = torch.empty(n, 128, dtype=torch.float64, device=dev)
        lib.call("ava_b200_window_time_tables", g0.data_ptr(), Kd.data_ptr(), bd.data_ptr(), kmax,
                 ts.data_ptr(), te.data_ptr(), n, 128, ti.data_ptr(), tf.data_ptr(),
                 torch.cuda.current_stream().cuda_stream)
        np.testing.assert_array_equal(ti.cpu().numpy(), t_idx)
        np.testing.assert_array_equal(tf.cpu().numpy(), t_frac)
        if n == 512:
            assert (t_idx < 0).any() and (t_idx >= 0).any()


def test_fixed_window_dataset_bit_exact_sampling(win, tmp_path):
    from scipy.io import wavfile
    g = load_golden("sampler_cases")
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    adir, rdir = tmp_path / "audio", tmp_path / "rois"
    adir.mkdir()
    rdir.mkdir()
    names = ["b_song", "a_song", "d_song", "c_song", "e_song"]
    rng = np.random.default_rng(5)
    for i, nm in enumerate(names):
        audio = spec_oracle.synth_audio(100 + i, int(3.0 * fs), fs)
        wavfile.write(str(adir / (nm + ".wav")), fs, audio)
        n_roi = 1 + (i % 3)
        starts = np.sort(rng.uniform(0.1, 2.0, size=n_roi))
        rois = np.stack([starts, starts + rng.uniform(0.2, 0.8, size=n_roi)], 1)
        np.savetxt(str(rdir / (nm + ".txt")), rois)
    part = win.get_window_partition([str(adir)], [str(rdir)], split=0.8)
    assert [os.path.basename(s) for s in part['train']['audio']] == list(g["part_train_audio"])
    assert [os.path.basename(s) for s in part['train']['rois']] == list(g["part_train_rois"])
    assert [os.path.basename(s) for s in part['test']['audio']] == list(g["part_test_audio"])
    part1 = win.get_window_partition([str(adir)], [str(rdir)], split=1.0)
    assert [os.path.basename(s) for s in part1['train']['audio']] == list(g["part1_audio"])
    ds = win.FixedWindowDataset(part1['train']['audio'], part1['train']['rois'], p)
    assert np.array_equal(ds.file_weights, g["file_weights"])
    for seed in (0, 1, 7):
        specs, fidx, onsets, offsets = ds.__getitem__(np.arange(12), seed=seed, return_seg_info=True)
        assert np.array_equal(np.array(fidx), g["seed%d_files" % seed])          # bit-exact
        assert np.array_equal(np.array(onsets), g["seed%d_onsets" % seed])       # bit-exact
        assert np.array_equal(np.array(offsets), g["seed%d_offsets" % seed])
        if seed == 0:
            got = torch.stack(specs[:3]).cpu().numpy()
            assert np.abs(got - g["seed0_specs"]).max() <= TOL
    # single index, loader protocol
    one = ds[3]
    assert one.shape == (128, 128) and one.dtype == torch.float32 and one.is_cuda
    loaders = win.get_fixed_window_data_loaders({'train': part1['train'], 'test': part1['train']}, p,
                                                batch_size=100)
    batches = list(loaders['train'])
    assert len(loaders['train'].dataset) == 2048 and len(batches) == 21
    assert batches[0].shape == (100, 128, 128) and batches[-1].shape == (48, 128, 128)
    # silence rejection keeps stream order: accepted windows are a subsequence of the
    # unfiltered candidate stream
    ds2 = win.FixedWindowDataset(part1['train']['audio'], part1['train']['rois'], p,
                                 min_spec_val=0.0)
    _, f2, on2, _ = ds2.__getitem__(np.arange(12), seed=7, return_seg_info=True)
    assert np.array_equal(np.array(on2), g["seed7_onsets"])


def test_warped_dataset_null_warp(win):
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    audio = [spec_oracle.synth_audio(40 + i, int(0.8 * fs), fs) for i in range(3)]
    ds = win.WarpedWindowDataset(["c.wav", "a.wav", "b.wav"], p, warp_type='null', audio=audio, fs=fs)
    specs = ds.__getitem__(np.arange(5), seed=3)
    # oracle: same draws, identity warp, whole-file get_spec with explicit target times
    np.random.seed(3)
    for i in range(5):
        fi = np.random.randint(3)
        start = ds.start_q + np.random.rand() * (ds.stop_q - ds.start_q - ds.window_frac)
        tv = np.linspace(start, start + ds.window_frac, 128) * ds.template_dur
        ref, _ = spec_oracle.get_spec(0.0, ds.template_dur, audio[fi], p, fs=fs, target_times=tv)
        assert np.abs(specs[i].cpu().numpy() - ref).max() <= TOL, i
    np.random.seed(None)


def test_syllable_loaders(tmp_path):
    ds_mod = importlib.import_module(PKG + ".models.vae_dataset")
    g = load_golden("sampler_cases")
    hdir = tmp_path / "h5"
    hdir.mkdir()
    for i in range(13):
        (hdir / ("syllables_%04d.hdf5" % i)).write_bytes(b"")
    (hdir / "notes.txt").write_bytes(b"")
    part = ds_mod.get_syllable_partition([str(hdir)], 0.75)
    assert [os.path.basename(s) for s in part['train']] == list(g["syll_train"])
    assert [os.path.basename(s) for s in part['test']] == list(g["syll_test"])
    # device-resident loader over .npy stand-ins for the HDF5 files (h5py is not installed)
    rng = np.random.default_rng(0)
    files = []
    for i in range(3):
        fn = str(tmp_path / ("sylls_%d.npy" % i))
        np.save(fn, rng.random((5, 128, 128)))
        files.append(fn)
    loaders = ds_mod.get_syllable_data_loaders({'train': files, 'test': files[:1]}, batch_size=4,
                                               shuffle=(False, False))
    tr = loaders['train']
    assert len(tr.dataset) == 15 and len(tr) == 4
    batches = list(tr)
    assert [b.shape[0] for b in batches] == [4, 4, 4, 3] and batches[0].is_cuda
    item = tr.dataset[7]           # file 1, row 2 (i // spf, i % spf)
    assert torch.equal(item, torch.from_numpy(np.load(files[1])[2]).float())
    assert torch.equal(batches[1][3].cpu(), item)
    several = tr.dataset[np.array([0, 14])]
    assert len(several) == 2 and several[1].shape == (128, 128)


def test_round2_goldens_float64_audio_and_within_syll_normalize(pre):
    """Reference goldens (tests/golden/round2_cases.npz): float64 audio goes to the device as
    float64 (the reference's complex128 STFT path; round 1 down-cast it to float32), and
    within_syll_normalize runs on the device (radix-select quantile + numpy's lerp)."""
    g = load_golden("round2_cases")
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    a64 = spec_oracle.synth_audio(21, int(2.0 * fs), fs, dtype=np.float64) / 7.0 + 0.123456789
    for name in ("f64_a", "f64_b"):
        onset = float(g[name + "_t"][0])
        tt = np.linspace(onset, onset + p['window_length'], 128)
        spec, _ = pre.get_spec(max(0.0, onset - 0.05), onset + p['window_length'] + 0.05, a64, p, fs=fs,
                               target_times=tt)
        assert np.abs(spec - g[name]).max() <= 1e-9, name          # fp64 end to end
    eng = pre.SpecEngine([a64], fs, p)
    assert eng.is_f32 == 2 and eng.audio_dev.dtype == torch.float64
    for q in (0.5, 0.87):
        pn = dict(spec_oracle.MOUSE_P)
        pn.update(within_syll_normalize=True, normalize_quantile=q)
        am = spec_oracle.synth_audio(11, int(0.6 * pn['fs']), pn['fs'])
        spec, _ = pre.get_spec(0.100, 0.180, am, pn, fs=pn['fs'])
        assert np.abs(spec - g["norm_q%02d" % int(100 * q)]).max() <= 1e-9, q
        assert spec.max() <= 1.0 and spec.min() == 0.0
    pn = dict(spec_oracle.FINCH_P)
    pn.update(within_syll_normalize=True, normalize_quantile=0.5)
    a2 = spec_oracle.synth_audio(12, int(4.0 * fs), fs)
    tt = np.linspace(1.2345, 1.2345 + pn['window_length'], 128)
    spec, _ = pre.get_spec(1.2345 - 0.05, 1.2345 + pn['window_length'] + 0.05, a2, pn, fs=fs, target_times=tt)
    assert np.abs(spec - g["norm_finch"]).max() <= 1e-9
    # batched engine: same normalisation per spectrogram, float32 output = cast of the float64 one
    eng = pre.SpecEngine([a2], fs, pn)
    onsets = np.array([1.2345, 0.5, 2.75])
    tts = np.linspace(onsets, onsets + pn['window_length'], 128, axis=-1)
    o32, o64 = eng.specs(np.zeros(3, dtype=np.int64), np.maximum(0, onsets - 0.05),
                         onsets + pn['window_length'] + 0.05, tts, want_float64=True)
    assert np.abs(o64[0].cpu().numpy() - g["norm_finch"]).max() <= 1e-9
    assert torch.equal(o32, o64.float())
    for i in (1, 2):
        pr = dict(pn)
        ref, _ = spec_oracle.get_spec(max(0, onsets[i] - 0.05), onsets[i] + pn['window_length'] + 0.05, a2, pr,
                                      fs=fs, target_times=tts[i])
        assert np.abs(o64[i].cpu().numpy() - ref).max() <= 1e-9, i


def test_fixed_window_dataset_silence_rejection_golden(win):
    """min_spec_val with REAL rejections (56 and 43 candidates for 16 accepted windows in the
    reference run): the chunked redraw visits the same candidates in the same order, so the
    accepted (file, onset) pairs are bit-identical to the reference's."""
    g = load_golden("round2_cases")
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    names = ["x_song.wav", "y_song.wav", "z_song.wav"]
    audio = [spec_oracle.silent_half_audio(200 + i, fs) for i in range(3)]
    rois = [np.array([[0.2, 2.8]])] * 3
    ds = win.FixedWindowDataset(names, None, p, audio=audio, fs=fs, rois=rois, min_spec_val=0.3)
    for seed in (0, 3):
        specs, fidx, onsets, offsets = ds.__getitem__(np.arange(16), seed=seed, return_seg_info=True)
        assert np.array_equal(np.array(fidx), g["rej_seed%d_files" % seed])
        assert np.array_equal(np.array(onsets), g["rej_seed%d_onsets" % seed])
        assert all(float(s.max()) >= 0.3 for s in specs)
        if seed == 0:
            got = torch.stack(specs[:2]).cpu().numpy()
            assert np.abs(got - g["rej_seed0_specs"]).max() <= TOL
    # without the threshold the same stream gives different (unfiltered) windows
    ds0 = win.FixedWindowDataset(names, None, p, audio=audio, fs=fs, rois=rois)
    _, _, on0, _ = ds0.__getitem__(np.arange(16), seed=0, return_seg_info=True)
    assert not np.array_equal(np.array(on0), g["rej_seed0_onsets"])


def test_window_datasets_honour_a_foreign_get_spec(win):
    """The reference's plugin point: a user-supplied p['get_spec'] is called per window with the
    reference's arguments (window_vae_dataset.py:233-235, 629-631) instead of the GPU engine;
    windows it flags invalid are redrawn."""
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    audio = [spec_oracle.synth_audio(60 + i, int(2.0 * fs), fs) for i in range(2)]
    calls = []

    def my_get_spec(t1, t2, audio_, p_, fs=32000, target_times=None, **kw):
        calls.append((t1, t2, len(audio_), tuple(sorted(kw))))
        spec, _ = spec_oracle.get_spec(t1, t2, audio_, p_, fs=fs, target_times=target_times)
        return 0.5 * spec, (len(calls) % 3 != 0)          # every third window is "invalid"
    p['get_spec'] = my_get_spec
    ds = win.FixedWindowDataset(["a.wav", "b.wav"], None, p, audio=audio, fs=fs,
                                rois=[np.array([[0.1, 1.9]])] * 2)
    specs, fidx, onsets, _ = ds.__getitem__(np.arange(6), seed=2, return_seg_info=True)
    assert len(specs) == 6 and len(calls) >= 8             # rejected windows were redrawn
    t1, t2, n_audio, kw = calls[0]
    assert n_audio == len(audio[0]) and abs((t2 - t1) - (p['window_length'] + 0.1)) < 1e-12
    ref, _ = spec_oracle.get_spec(max(0.0, onsets[0] - 0.05), onsets[0] + p['window_length'] + 0.05,
                                  audio[fidx[0]], p, fs=fs,
                                  target_times=np.linspace(onsets[0], onsets[0] + p['window_length'], 128))
    assert specs[0].is_cuda and np.abs(specs[0].cpu().numpy() - 0.5 * ref).max() <= 1e-6
    # this package's own get_spec in p['get_spec'] keeps the batched GPU path
    pre = importlib.import_module(PKG + ".preprocessing.utils")
    p2 = dict(spec_oracle.FINCH_P)
    p2['get_spec'] = pre.get_spec
    assert win._foreign_get_spec(p2) is None and win._foreign_get_spec(p) is my_get_spec
    wds = win.WarpedWindowDataset(["a.wav", "b.wav"], p, warp_type='null', audio=audio, fs=fs)
    n0 = len(calls)
    out = wds.__getitem__(np.arange(3), seed=1)
    assert len(out) == 3 and len(calls) == n0 + 3 and calls[-1][3] == ('max_dur',)


def test_write_hdf5_files_and_get_specific_item(win, tmp_path, monkeypatch):
    """write_hdf5_files (window_vae_dataset.py:259-293): file k holds __getitem__(arange(n),
    seed=k) -- float32 'specs' and the byte-string 'audio_filenames' -- through a stand-in for
    h5py.File when h5py is not installed; get_specific_item (:643-670) returns the float64
    spectrogram of a given file / quantile."""
    import sys
    import types
    store = {}

    class _File:
        def __init__(self, fn, mode="r"):
            self.fn = fn
            store.setdefault(fn, {})

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def create_dataset(self, key, data=None):
            store[self.fn][key] = np.array(data)
    try:
        import h5py  # noqa: F401
        have = True
    except ImportError:
        have = False
        fake = types.ModuleType("h5py")
        fake.File = _File
        monkeypatch.setitem(sys.modules, "h5py", fake)
    p = dict(spec_oracle.FINCH_P)
    fs = p['fs']
    audio = [spec_oracle.synth_audio(70 + i, int(2.0 * fs), fs) for i in range(3)]
    names = ["c.wav", "a.wav", "b.wav"]
    ds = win.FixedWindowDataset(names, None, p, audio=audio, fs=fs, rois=[np.array([[0.1, 1.9]])] * 3)
    out_dir = str(tmp_path / "h5")
    ds.write_hdf5_files(out_dir, num_files=2, sylls_per_file=5)
    for k in range(2):
        fn = os.path.join(out_dir, "syllables_%04d.hdf5" % k)
        if have:
            import h5py
            with h5py.File(fn, "r") as f:
                specs, fns = np.array(f['specs']), np.array(f['audio_filenames'])
        else:
            specs, fns = store[fn]['specs'], store[fn]['audio_filenames']
        want, fidx, _, _ = ds.__getitem__(np.arange(5), seed=k, return_seg_info=True)
        assert specs.shape == (5, 128, 128) and specs.dtype == np.float32
        assert np.array_equal(specs, torch.stack(want).cpu().numpy())
        assert [s.decode() for s in fns] == [ds.filenames[i] for i in fidx]
    wds = win.WarpedWindowDataset(names, p, warp_type='null', audio=audio, fs=fs)
    item = wds.get_specific_item("b.wav", 0.4)
    assert item.shape == (128, 128) and item.dtype == np.float64
    fi = wds.audio_filenames.index("b.wav")
    start = wds.start_q + 0.4 * (wds.stop_q - wds.start_q - wds.window_frac)
    tv = np.linspace(start, start + wds.window_frac, 128) * wds.template_dur
    ref, _ = spec_oracle.get_spec(0.0, wds.template_dur, audio[fi], p, fs=fs, target_times=tv)
    assert np.abs(item - ref).max() <= 1e-9
