#!/usr/bin/env python
"""
Secondary measurements for the other BASELINE.json configs (one JSON object per line):

  * train step at batch 64 (configs[1], CUDA-graph replay) and 256
  * get_latent-style inference (configs[4]): encode-only specs/s, eval-mode BN, batch 1024
  * shotgun front end (configs[3]): GPU get_spec windows/s on a synthetic finch-style corpus
    (kernel only, and including the host-side sampling + coordinate tables), and end-to-end
    train samples/s with windows generated on the fly

    python bench_extra.py > profiles/r01_extra.jsonl
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "autoencoded-vocal-analysis_b200"

import numpy as np  # noqa: E402
import torch  # noqa: E402

FINCH_P = {  # examples/finch_window_mwe.py:29-49
    'fs': 32000, 'num_freq_bins': 128, 'num_time_bins': 128, 'nperseg': 512, 'noverlap': 256,
    'max_dur': 1e9, 'window_length': 0.12, 'min_freq': 400, 'max_freq': 10e3, 'spec_min_val': 2.0,
    'spec_max_val': 6.5, 'mel': True, 'time_stretch': False, 'within_syll_normalize': False,
}


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, 1e3 * (time.perf_counter() - t0) / iters


def n1_section(out):
    """PCA of latent means (csrc/pca.cu) and the DataContainer latent-mean
    path on a synthetic corpus of syllable files."""
    import tempfile
    dcm = importlib.import_module(PKG + ".data.data_container")
    mu = importlib.import_module(PKG + ".models.utils")
    vae_mod = importlib.import_module(PKG + ".models.vae")
    from oracle import pca_oracle
    N, D = 4_000_000, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, D, device="cuda", dtype=torch.float64, generator=g) * \
        torch.linspace(3.0, 0.2, D, device="cuda", dtype=torch.float64)
    pca = dcm.LatentPCA(2)
    pca.fit_transform_device(x[:4096])
    torch.cuda.synchronize()
    lib = importlib.import_module(PKG + "._lib")
    ws_bytes = int(lib.lib().ava_b200_pca_ws_bytes(D))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    mean, ev = (torch.empty(D, dtype=torch.float64, device="cuda") for _ in range(2))
    cov, comps = (torch.empty(D, D, dtype=torch.float64, device="cuda") for _ in range(2))
    emb = torch.empty(N, 2, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fit = lambda: lib.call("ava_b200_pca_fit", x.data_ptr(), 0, N, D, mean.data_ptr(), cov.data_ptr(),  # noqa: E731
                           ev.data_ptr(), comps.data_ptr(), ws.data_ptr(), ws_bytes, st)
    tr = lambda: lib.call("ava_b200_pca_transform", x.data_ptr(), 0, N, D, mean.data_ptr(),  # noqa: E731
                          comps.data_ptr(), 2, emb.data_ptr(), st)
    ms_fit, _ = timed(fit, 10)
    ms_tr, _ = timed(tr, 10)
    xs = x[:200000].cpu().numpy()
    t0 = time.perf_counter()
    pca_oracle.pca_fit_transform(xs, 2)
    cpu = time.perf_counter() - t0
    out.append({"what": "LatentPCA fit + transform of %d x %d float64 latent means (device resident)" % (N, D),
                "fit_ms": ms_fit, "transform_ms": ms_tr,
                "fit_GBps": N * D * 8 / (ms_fit * 1e-3) / 1e9,
                "transform_GBps": N * (D + 2) * 8 / (ms_tr * 1e-3) / 1e9,
                "rows_per_s": N / ((ms_fit + ms_tr) * 1e-3),
                "cpu_port_rows_per_s": len(xs) / cpu,
                "cpu_sample": "200000 rows through oracle/pca_oracle.pca_fit_transform (numpy float64, all cores)"})
    del x, emb
    # DataContainer.request('latent_means'): 64 files x 512 syllables, reference semantics
    # (train-mode BN, batches of 64) and the eval-mode / batch-1024 variant
    ext = ".npz"
    try:
        import h5py  # noqa: F401
        ext = ".hdf5"
    except ImportError:
        pass
    nf, spf = 64, 512
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory() as root:
        sd, pd = os.path.join(root, "specs"), os.path.join(root, "proj")
        os.makedirs(sd)
        for j in range(nf):
            mu.append_field(os.path.join(sd, "syllables_%04d%s" % (j, ext)), 'specs',
                            rng.random((spf, 128, 128), dtype=np.float32).astype(np.float64))
        model = vae_mod.VAE(save_dir=root)
        model.save_state("checkpoint_000.tar")
        del model
        for kw in ({}, {"latent_batch_size": 1024, "latent_eval": True}):
            dc = dcm.DataContainer(spec_dirs=[sd], projection_dirs=[pd], verbose=False,
                                   model_filename=os.path.join(root, "checkpoint_000.tar"), **kw)
            dc.clear_projections() if os.path.exists(pd) else None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lat = dc.request('latent_mean_pca')
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out.append({"what": "DataContainer latent_means + latent_mean_pca, %d files x %d syllables "
                                "(float64 spec files read from disk, projections written)" % (nf, spf),
                        "options": kw or "reference semantics (train-mode BN, batch 64)",
                        "seconds": dt, "syllables_per_s": len(lat) / dt})


def shotgun_section(out):
    """BASELINE configs[3] / SURVEY 8(d) config 4: synthetic corpus of 16 files x 600 s @ 32 kHz
    int16, 2 ROIs per file, finch parameters; windows/s of the sampler + GPU get_spec, and the
    train step fed with windows generated on the fly next to the same step on a resident batch
    ("training never waits" <=> the two rates agree)."""
    vae_mod = importlib.import_module(PKG + ".models.vae")
    win_mod = importlib.import_module(PKG + ".models.window_vae_dataset")
    rng = np.random.default_rng(0)
    fs = FINCH_P['fs']
    n_files, dur = 16, 600.0
    audio = [(3000 * rng.standard_normal(int(dur * fs), dtype=np.float32)).astype(np.int16)
             for _ in range(n_files)]
    rois = [np.array([[1.0, 250.0], [300.0, 598.0]]) for _ in range(n_files)]
    names = ["f%02d.wav" % k for k in range(n_files)]
    ds = win_mod.FixedWindowDataset(names, None, dict(FINCH_P), audio=audio, fs=fs, rois=rois)
    for nb in (128, 1024):
        ms, wall = timed(lambda: ds.sample_batch(nb), 20)
        out.append({"what": "shotgun windows (host sampling + device tables + GPU get_spec)", "batch": nb,
                    "ms_per_batch_gpu": ms, "ms_per_batch_wall": wall, "windows_per_s": nb / (wall * 1e-3)})
    model = vae_mod.VAE(device_name='cuda')
    model.train()
    for nb in (128, 1024):
        x = ds.sample_batch(nb)
        ms0, wall0 = timed(lambda: model.train_step(x), 15)
        ms, wall = timed(lambda: model.train_step(ds.sample_batch(nb)), 15)
        out.append({"what": "shotgun train step with on-the-fly GPU get_spec", "batch": nb,
                    "ms_per_step_wall": wall, "samples_per_s": nb / (wall * 1e-3),
                    "resident_batch_ms_per_step": ms0, "resident_batch_samples_per_s": nb / (ms0 * 1e-3),
                    "ratio_vs_resident": ms0 / wall})


def main():
    for key, fn in (("n1", n1_section), ("shotgun", shotgun_section)):
        if key in sys.argv[1:]:
            out = []
            fn(out)
            for o in out:
                print(json.dumps(o))
    if len(sys.argv) > 1:
        return
    vae_mod = importlib.import_module(PKG + ".models.vae")
    win_mod = importlib.import_module(PKG + ".models.window_vae_dataset")
    torch.manual_seed(0)
    out = []
    # ---- train step at batch 1024 per precision mode (bench.py's headline is 'auto' = tf32x3)
    for prec in ("fp32", "tf32x3", "tf32"):
        model = vae_mod.VAE(device_name='cuda', precision=prec, cuda_graphs=False)
        model.train()
        xs = [torch.rand(1024, 128, 128, device="cuda") for _ in range(2)]
        k = [0]

        def step():
            k[0] += 1
            return model.train_step(xs[k[0] & 1])
        ms, wall = timed(step, 10, warmup=3)
        out.append({"what": "train_step", "batch": 1024, "precision": prec, "ms_per_step": ms,
                    "samples_per_s": 1024 / (ms * 1e-3)})
        del model, xs
    # ---- train step at small batches
    for B, graphs in ((64, True), (64, False), (256, True)):
        model = vae_mod.VAE(device_name='cuda', cuda_graphs=graphs)
        model.train()
        x = torch.rand(B, 128, 128, device="cuda")
        ms, wall = timed(lambda: model.train_step(x), 100, warmup=6)
        out.append({"what": "train_step", "batch": B, "cuda_graph": graphs, "us_per_step": 1e3 * ms,
                    "samples_per_s": B / (wall * 1e-3)})
        del model
    # ---- inference: encode only, eval-mode BN (get_latent inner loop)
    model = vae_mod.VAE(device_name='cuda')
    model.eval()
    B = 1024
    xs = [torch.rand(B, 128, 128, device="cuda") for _ in range(2)]
    i = [0]

    def enc():
        bufs = model._buffers_for(B)
        model._cur = bufs
        model._scratch_need = model._scratch_need_for(B)
        model._encode_native(xs[i[0] & 1], bufs, False)
        i[0] += 1
    with torch.no_grad():
        ms, wall = timed(enc, 30)
    out.append({"what": "get_latent encode (eval BN)", "batch": B, "ms_per_batch": ms,
                "specs_per_s": B / (ms * 1e-3),
                "hbm_GBps_layer_granular": 2.36e6 * B / (ms * 1e-3) / 1e9})
    # ---- shotgun: synthetic corpus 16 files x 60 s @ 32 kHz int16, 2 ROIs per file
    rng = np.random.default_rng(0)
    fs = FINCH_P['fs']
    n_files, dur = 16, 60.0
    audio = [(3000 * rng.standard_normal(int(dur * fs))).astype(np.int16) for _ in range(n_files)]
    rois = [np.array([[1.0, 25.0], [30.0, 58.0]]) for _ in range(n_files)]
    names = ["f%02d.wav" % k for k in range(n_files)]
    ds = win_mod.FixedWindowDataset(names, None, dict(FINCH_P), audio=audio, fs=fs, rois=rois)
    for nb in (128, 1024):
        ms, wall = timed(lambda: ds.sample_batch(nb), 20)
        out.append({"what": "shotgun windows (host sampling + tables + GPU get_spec)", "batch": nb,
                    "ms_per_batch_gpu": ms, "ms_per_batch_wall": wall, "windows_per_s": nb / (wall * 1e-3)})
    # kernel only: same windows re-issued (tables already on the device)
    nb = 1024
    np.random.seed(0)
    files, onsets = ds._draw(nb)
    np.random.seed(None)
    eng = ds._engine
    wl = FINCH_P['window_length']
    tt = np.linspace(onsets, onsets + wl, 128, axis=-1)
    outbuf = torch.empty(nb, 128, 128, device="cuda")
    lib = importlib.import_module(PKG + "._lib")
    eng.specs(files, np.maximum(0, onsets - 0.05), onsets + wl + 0.05, tt, out=outbuf)
    torch.cuda.synchronize()
    captured = {}
    orig_call = lib.call

    def spy(name, *a):
        captured["args"] = (name, a)
        return orig_call(name, *a)
    pre = importlib.import_module(PKG + ".preprocessing.utils")
    pre.call = spy
    eng.specs(files, np.maximum(0, onsets - 0.05), onsets + wl + 0.05, tt, out=outbuf)
    pre.call = orig_call
    keep = eng._keepalive
    name, a = captured["args"]
    ms, _ = timed(lambda: orig_call(name, *a), 50)
    seg_bytes = int((0.22 * fs)) * 2
    out.append({"what": "get_spec kernel only (fp64 STFT+log+resample)", "batch": nb, "us_per_launch": 1e3 * ms,
                "windows_per_s": nb / (ms * 1e-3),
                "hbm_GBps_algorithmic": nb * (seg_bytes + 65536) / (ms * 1e-3) / 1e9})
    del keep
    # end to end: windows generated on the fly feeding the train step
    model = vae_mod.VAE(device_name='cuda')
    model.train()
    for nb in (128, 1024):
        ms, wall = timed(lambda: model.train_step(ds.sample_batch(nb)), 15)
        out.append({"what": "shotgun train step with on-the-fly GPU get_spec", "batch": nb,
                    "ms_per_step_wall": wall, "samples_per_s": nb / (wall * 1e-3)})
    # ---- syllable preprocessing driver (all syllables of a file in one launch)
    pp = importlib.import_module(PKG + ".preprocessing.preprocess")
    from oracle import spec_oracle
    ps = dict(FINCH_P)
    ps.update(max_dur=0.2, time_stretch=True)
    fs = ps['fs']
    audio = spec_oracle.synth_audio(5, int(120 * fs), fs)
    rng = np.random.default_rng(0)
    on = np.sort(rng.uniform(0.1, 119.0, size=2000))
    off = on + rng.uniform(0.03, 0.19, size=2000)
    tf = pp._inv_mel(np.linspace(pp._mel(ps['min_freq']), pp._mel(ps['max_freq']), ps['num_freq_bins']))
    pp._syll_specs_batched(on[:64], off[:64], audio, fs, ps, tf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    specs, _ = pp._syll_specs_batched(on, off, audio, fs, ps, tf)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(40):
        spec_oracle.get_spec(on[i], off[i], audio, ps, fs=fs, target_freqs=tf)
    cpu = (time.perf_counter() - t0) / 40
    out.append({"what": "process_sylls: get_syll_specs of one file (2000 syllables, incl. upload of 120 s "
                        "audio and float64 download)", "syllables_per_s": len(specs) / dt,
                "cpu_port_syllables_per_s_1core": 1.0 / cpu,
                "cpu_sample": "40 syllables through oracle/spec_oracle.get_spec (numpy float64)"})
    # ---- MMD^2 matrix between conditions of a corpus of latent means
    mmd = importlib.import_module(PKG + ".plotting.mmd_plots")
    from oracle import mmd_oracle
    N, ncond = 18020, 8            # corpus size of docs/source/data_management.rst:73
    lat = rng.standard_normal((N, 32))
    cond = rng.integers(0, ncond, size=N)
    mmd.mmd2_matrix(lat[:512], cond[:512], sigma=8.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m, _ = mmd.mmd2_matrix(lat, cond, sigma=8.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    i1, i2 = np.argwhere(cond == 0).flatten()[:600], np.argwhere(cond == 1).flatten()[:600]
    t0 = time.perf_counter()
    mmd_oracle.estimate_mmd2(lat, i1, i2, 8.0)
    cpu = time.perf_counter() - t0
    cpu_pairs = (600 * 599 + 600 * 600)
    out.append({"what": "mmd2_matrix: all %d condition pairs of %d latent means (fp64, incl. upload)" % (ncond * (ncond - 1) // 2, N),
                "ms": 1e3 * dt, "kernel_evals_per_s": N * N / dt,
                "cpu_port_kernel_evals_per_s": cpu_pairs / cpu,
                "cpu_sample": "one pair of 600-point conditions through oracle/mmd_oracle.estimate_mmd2 (vectorised numpy; "
                              "the reference itself is a Python double loop)"})
    n1_section(out)
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
