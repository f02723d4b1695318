"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the `get_spec` spectrogram front end
and the shotgun window sampler.

Plain numpy (float64) restatement.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this file.

Parity status: PINNED against the reference's own `get_spec`
(ava/preprocessing/utils.py:18-110, run in the build container with the interp2d
shim of oracle/_ref_import.py over scipy 1.18.1's `stft`) and against the
reference's `FixedWindowDataset.__getitem__` (ava/models/window_vae_dataset.py:
189-256) through tests/golden/spec_*.npz and tests/golden/sampler_*.npz, generated
by oracle/make_golden.py.  The reference ships no tests of its own.

Third-party arithmetic restated here:
  * scipy.signal.stft (scipy unpinned in the reference's setup.py:31; 1.18.1
    installed) -- scipy/signal/_spectral_py.py `_spectral_helper`: periodic Hann
    window, boundary='zeros' (nperseg//2 zeros each side), padded=True (zero pad
    to a whole number of hops), no detrend, one-sided FFT, scale 1/sum(window).
  * scipy.interpolate.interp2d(kind='linear', bounds_error=False,
    fill_value=v): plain bilinear interpolation on the (t, f) grid; any target
    strictly outside [t0, t_last] (column) or [f0, f_last] (row) is `v`.
  * numpy legacy RandomState (np.random.seed/choice/rand): used as is -- the
    sampler *is* numpy's stream, so the restatement calls numpy itself.
"""
import numpy as np

EPSILON = 1e-12  # ava/preprocessing/utils.py:14


def mel(a):
    """ava/preprocessing/utils.py:113-115"""
    return 1127 * np.log(1 + a / 700)


def inv_mel(a):
    """ava/preprocessing/utils.py:118-120"""
    return 700 * (np.exp(a / 1127) - 1)


def stft(x, fs, nperseg, noverlap):
    """scipy.signal.stft(x, fs, nperseg=nperseg, noverlap=noverlap) with scipy's
    defaults, as called at ava/preprocessing/utils.py:76-77.
    Returns (f [nperseg//2+1], t [K], Z [nperseg//2+1, K] complex128)."""
    x = np.asarray(x, dtype=np.float64)
    hop = nperseg - noverlap
    n = np.arange(nperseg)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)   # periodic Hann
    half = nperseg // 2
    ext = np.concatenate([np.zeros(half), x, np.zeros(half)])
    nadd = (-(ext.shape[0] - nperseg) % hop) % nperseg
    ext = np.concatenate([ext, np.zeros(nadd)])
    K = (ext.shape[0] - nperseg) // hop + 1
    frames = np.stack([ext[k * hop:k * hop + nperseg] for k in range(K)])
    Z = np.fft.rfft(frames * win, n=nperseg, axis=-1) / win.sum()
    # scipy computes the frame times on the extended *and* hop-padded signal.
    n_ext = ext.shape[0]
    t = np.arange(nperseg / 2, n_ext - nperseg / 2 + 1, hop) / float(fs)
    t = t - (nperseg / 2) / fs
    f = np.fft.rfftfreq(nperseg, 1 / fs)
    assert len(t) == K, (len(t), K)
    return f, t, Z.T


def bilinear_fill(xg, yg, S, x, y, fill_value):
    """interp2d(xg, yg, S, kind='linear', bounds_error=False, fill_value)(x, y):
    S is [len(yg), len(xg)]; returns [len(y), len(x)]."""
    def _coords(g, q):
        i = np.searchsorted(g, q, side='right') - 1
        i = np.clip(i, 0, len(g) - 2)
        w = (q - g[i]) / (g[i + 1] - g[i])
        bad = (q < g[0]) | (q > g[-1])
        return i, w, bad
    ix, wx, badx = _coords(xg, x)
    iy, wy, bady = _coords(yg, y)
    s00 = S[np.ix_(iy, ix)]
    s01 = S[np.ix_(iy, ix + 1)]
    s10 = S[np.ix_(iy + 1, ix)]
    s11 = S[np.ix_(iy + 1, ix + 1)]
    wx = wx[None, :]
    wy = wy[:, None]
    out = (1 - wy) * ((1 - wx) * s00 + wx * s01) + wy * ((1 - wx) * s10 + wx * s11)
    out[:, badx] = fill_value
    out[bady, :] = fill_value
    return out


def get_spec(t1, t2, audio, p, fs=32000, target_freqs=None, target_times=None,
             fill_value=-1 / EPSILON, max_dur=None, remove_dc_offset=True):
    """ava/preprocessing/utils.py:18-110, line by line."""
    if max_dur is None:
        max_dur = p['max_dur']
    s1, s2 = int(round(t1 * fs)), int(round(t2 * fs))          # :65
    assert s1 < s2
    temp = min(len(audio), s2) - max(0, s1)                     # :69
    if temp < p['nperseg'] or s2 <= 0 or s1 >= len(audio):      # :70-71
        return np.zeros((p['num_freq_bins'], p['num_time_bins'])), True
    temp_audio = audio[max(0, s1):min(len(audio), s2)]          # :73
    if remove_dc_offset:
        temp_audio = temp_audio - np.mean(temp_audio)           # :74-75
    f, t, spec = stft(temp_audio, fs, p['nperseg'], p['noverlap'])
    t = t + max(0, t1)                                          # :78
    spec = np.log(np.abs(spec) + EPSILON)                       # :79
    if target_freqs is None:                                    # :83-90
        if p['mel']:
            target_freqs = np.linspace(mel(p['min_freq']), mel(p['max_freq']),
                                       p['num_freq_bins'])
            target_freqs = inv_mel(target_freqs)
        else:
            target_freqs = np.linspace(p['min_freq'], p['max_freq'],
                                       p['num_freq_bins'])
    if target_times is None:                                    # :92-97
        duration = t2 - t1
        if p['time_stretch']:
            duration = np.sqrt(duration * max_dur)
        shoulder = 0.5 * (max_dur - duration)
        target_times = np.linspace(t1 - shoulder, t2 + shoulder, p['num_time_bins'])
    spec = bilinear_fill(t, f, spec, np.asarray(target_times, dtype=np.float64),
                         np.asarray(target_freqs, dtype=np.float64), fill_value)
    spec -= p['spec_min_val']                                   # :102
    spec /= (p['spec_max_val'] - p['spec_min_val'])             # :103
    spec = np.clip(spec, 0.0, 1.0)                              # :104
    if p['within_syll_normalize']:                              # :106-109
        spec -= np.quantile(spec, p['normalize_quantile'])
        spec[spec < 0.0] = 0.0
        spec /= np.max(spec) + EPSILON
    return spec, True


def sample_windows(n, seed, rois, file_weights, roi_weights, window_length):
    """The random draws of FixedWindowDataset.__getitem__ for `n` accepted items
    when no item is rejected (ava/models/window_vae_dataset.py:215-231): per
    item exactly three doubles from numpy's legacy global stream, in the order
    file, roi, onset.  Returns (file_indices int64 [n], onsets float64 [n])."""
    np.random.seed(seed)
    files = np.zeros(n, dtype=np.int64)
    onsets = np.zeros(n, dtype=np.float64)
    for i in range(n):
        fi = np.random.choice(np.arange(len(file_weights)), p=file_weights)
        ri = np.random.choice(np.arange(len(roi_weights[fi])), p=roi_weights[fi])
        roi = rois[fi][ri]
        onsets[i] = roi[0] + (roi[1] - roi[0] - window_length) * np.random.rand()
        files[i] = fi
    np.random.seed(None)
    return files, onsets


def fixed_window_item(audio, fs, p, file_index, onset, shoulder=0.05):
    """Spectrogram of one sampled window
    (ava/models/window_vae_dataset.py:229-235)."""
    offset = onset + p['window_length']
    target_times = np.linspace(onset, offset, p['num_time_bins'])
    spec, _ = get_spec(max(0.0, onset - shoulder), offset + shoulder,
                       audio[file_index], p, fs=fs, target_times=target_times)
    return spec


# ------------------------------------------------------------------ synthetic
def synth_audio(seed, n_samples, fs, scale=3000.0, dtype=np.int16):
    """Band-limited noise + chirps, int16-scale."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / fs
    x = rng.standard_normal(n_samples)
    # mild colouring (2-tap) so the spectrum is not flat
    x = 0.7 * x + 0.3 * np.roll(x, 1)
    f0 = 0.05 * fs + 0.2 * fs * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
    chirp = np.sin(2 * np.pi * np.cumsum(f0) / fs)
    env = 0.5 + 0.5 * np.sin(2 * np.pi * 7.0 * t + rng.uniform(0, 6.28))
    y = scale * (0.3 * x + env * chirp)
    if np.issubdtype(dtype, np.integer):
        return np.clip(np.round(y), -32768, 32767).astype(dtype)
    return y.astype(dtype)


MOUSE_P = {  # examples/mouse_sylls_mwe.py:56-79
    'max_dur': 0.2, 'min_freq': 30e3, 'max_freq': 110e3, 'num_freq_bins': 128,
    'num_time_bins': 128, 'nperseg': 1024, 'noverlap': 512, 'spec_min_val': 2.0,
    'spec_max_val': 6.0, 'fs': 250000, 'mel': False, 'time_stretch': True,
    'within_syll_normalize': False,
}

FINCH_P = {  # examples/finch_window_mwe.py:29-49
    'fs': 32000, 'num_freq_bins': 128, 'num_time_bins': 128, 'nperseg': 512,
    'noverlap': 256, 'max_dur': 1e9, 'window_length': 0.12, 'min_freq': 400,
    'max_freq': 10e3, 'spec_min_val': 2.0, 'spec_max_val': 6.5, 'mel': True,
    'time_stretch': False, 'within_syll_normalize': False,
}
