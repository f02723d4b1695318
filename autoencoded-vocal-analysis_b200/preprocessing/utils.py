"""
B200-native spectrogram front end: drop-in for ``ava.preprocessing.utils``.

``get_spec`` keeps the reference signature and return contract
(ava/preprocessing/utils.py:18-110): ``(spec float64 [freq_bins, time_bins], True)``.
The STFT, log-magnitude, bilinear resampling and normalise/clip run in one CUDA kernel
(csrc/spec.cu, fp64); the host only does what the reference does in float64 *coordinate*
space -- sample indices, frame-time grid, target times/frequencies and their bracketing
indices -- with the same numpy expressions, so in/out-of-range decisions are bit-identical.

``SpecEngine`` is the batched form used by the datasets: audio stays resident in HBM and a
whole batch of windows is one kernel launch writing ``[n, 128, 128]`` fp32 in place.

``within_syll_normalize`` (the per-spectrogram quantile normalisation) also runs on the device
(``ava_b200_quantile_normalize``), in the single-segment and the batched path alike.

No CPU fallback: without a CUDA device / the native library these functions raise.
"""
import warnings

import numpy as np
import torch

from .. import _lib
from .._lib import call

EPSILON = 1e-12


def _mel(a):
    """https://en.wikipedia.org/wiki/Mel-frequency_cepstrum (utils.py:113-115)"""
    return 1127 * np.log(1 + a / 700)


def _inv_mel(a):
    """https://en.wikipedia.org/wiki/Mel-frequency_cepstrum (utils.py:118-120)"""
    return 700 * (np.exp(a / 1127) - 1)


def _hann_window(nperseg):
    """The analysis window scipy.signal.stft uses by default (periodic Hann), obtained
    from scipy itself when available so the values are bit-identical."""
    try:
        from scipy.signal import get_window
        return np.asarray(get_window('hann', nperseg), dtype=np.float64)
    except ImportError:
        n = np.arange(nperseg)
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)


def target_freqs_for(p):
    """ava/preprocessing/utils.py:83-90"""
    if p['mel']:
        tf = np.linspace(_mel(p['min_freq']), _mel(p['max_freq']), p['num_freq_bins'])
        return _inv_mel(tf)
    return np.linspace(p['min_freq'], p['max_freq'], p['num_freq_bins'])


def num_frames(n_samples, nperseg, hop):
    """Number of STFT frames scipy produces (boundary='zeros', padded=True)."""
    n_samples = np.asarray(n_samples, dtype=np.int64)
    L = n_samples + 2 * (nperseg // 2)
    nadd = (-(L - nperseg) % hop) % nperseg
    return (L + nadd - nperseg) // hop + 1


def bracket(grid0, step_grid, K, targets):
    """For every target, the index i of the grid interval [g_i, g_{i+1}] containing it and
    the weight (T-g_i)/(g_{i+1}-g_i); i = -1 where T < g_0 or T > g_{K-1}
    (scipy interp2d(bounds_error=False) fill rule, ava/preprocessing/utils.py:80-81,99).

    grid0: [n] offset added to the shared frame-time base `step_grid` [Kmax] (the
    reference's ``t += max(0, t1)``); K: [n] frames per row; targets: [n, m]."""
    n, m = targets.shape
    g = step_grid[None, :] + grid0[:, None]                      # [n, Kmax], same adds as numpy
    dt = step_grid[1] - step_grid[0]
    rows = np.arange(n)[:, None]
    Km2 = (K - 2)[:, None]
    i = np.floor((targets - g[:, :1]) / dt).astype(np.int64)
    i = np.clip(i, 0, Km2)
    for _ in range(2):                                           # exact fix-up against the grid
        i = np.where(g[rows, i] > targets, i - 1, i)
        i = np.clip(i, 0, Km2)
        i = np.where((g[rows, i + 1] <= targets) & (i < Km2), i + 1, i)
    lo, hi = g[rows, i], g[rows, i + 1]
    w = (targets - lo) / (hi - lo)
    last = g[rows, (K - 1)[:, None]]
    bad = (targets < g[:, :1]) | (targets > last)
    return np.where(bad, -1, i).astype(np.int32), np.where(bad, 0.0, w)


def bracket_1d(grid, targets):
    """Same for one fixed grid (the frequency axis)."""
    i = np.searchsorted(grid, targets, side='right') - 1
    i = np.clip(i, 0, len(grid) - 2)
    w = (targets - grid[i]) / (grid[i + 1] - grid[i])
    bad = (targets < grid[0]) | (targets > grid[-1])
    return np.where(bad, -1, i).astype(np.int32), np.where(bad, 0.0, w)


class SpecEngine:
    """Device-resident audio + batched `get_spec`.

    Parameters
    ----------
    audio : list of 1-D numpy arrays (int16 or float32/float64), one per file
    fs : sample rate
    p : preprocessing parameter dict (same keys the reference's get_spec reads)
    """

    def __init__(self, audio, fs, p, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ava_b200 SpecEngine requires a CUDA device; no CPU fallback")
        _lib.lib()
        self.device = torch.device(device) if device is not None else \
            torch.device("cuda", torch.cuda.current_device())
        self.fs = fs
        self.p = p
        # sample type on the device: int16 as is; float32 as is; anything else (float64, wider
        # integers) as float64 -- the reference's scipy stft runs in complex128 on all of these
        # and in complex64 only on float32 audio (the kernel computes in fp64 throughout)
        dtypes = {np.asarray(a).dtype for a in audio}
        if dtypes <= {np.dtype(np.int16)}:
            self.is_f32, dt = 0, np.int16
        elif dtypes <= {np.dtype(np.float32)}:
            self.is_f32, dt = 1, np.float32
        else:
            self.is_f32, dt = 2, np.float64
        self.lengths = np.array([len(a) for a in audio], dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)
        flat = np.concatenate([np.asarray(a).astype(dt, copy=False) for a in audio]) if len(audio) \
            else np.zeros(0, dt)
        self.audio_dev = torch.from_numpy(flat).to(self.device)
        self.nperseg = int(p['nperseg'])
        self.noverlap = int(p['noverlap'])
        self.hop = self.nperseg - self.noverlap
        win = _hann_window(self.nperseg)
        self.scale = float(1.0 / win.sum())
        self.window_dev = torch.from_numpy(win).to(self.device)
        self._freq_cache = None
        self._base_cache = None
        self._pin_rings = {}

    def _freq_tables(self, target_freqs):
        f = np.fft.rfftfreq(self.nperseg, 1 / self.fs)
        if target_freqs is None:
            if self._freq_cache is None:
                idx, w = bracket_1d(f, target_freqs_for(self.p))
                self._freq_cache = (torch.from_numpy(idx).to(self.device),
                                    torch.from_numpy(w).to(self.device), idx)
            return self._freq_cache
        idx, w = bracket_1d(f, np.asarray(target_freqs, dtype=np.float64))
        return torch.from_numpy(idx).to(self.device), torch.from_numpy(w).to(self.device), idx

    def _segments(self, file_index, t1, t2):
        """Sample ranges and STFT frame counts of n segments [t1_i, t2_i]
        (ava/preprocessing/utils.py:66-72): (seg_start int64 [n] into the concatenated audio,
        seg_len int32 [n] with 0 for the reference's "too short" branch, K frames [n], kmax,
        base frame times [kmax] of a segment starting at t = 0)."""
        fs = self.fs
        # int(round(t*fs)): Python round == numpy rint (half to even)
        s1 = np.rint(t1 * fs).astype(np.int64)
        s2 = np.rint(t2 * fs).astype(np.int64)
        assert np.all(s1 < s2), "s1 must be < s2"
        flen = self.lengths[file_index]
        lo, hi = np.maximum(0, s1), np.minimum(flen, s2)
        seg_len = hi - lo
        short = (seg_len < self.nperseg) | (s2 <= 0) | (s1 >= flen)     # utils.py:69-71
        seg_len = np.where(short, 0, seg_len)
        K = np.where(short, 3, num_frames(np.maximum(seg_len, self.nperseg), self.nperseg, self.hop))
        kmax = int(K.max()) if len(K) else 3
        base = np.arange(self.nperseg / 2, self.nperseg / 2 + kmax * self.hop, self.hop) / float(fs)
        base = base - (self.nperseg / 2) / fs
        return self.offsets[file_index] + lo, seg_len.astype(np.int32), K, kmax, base

    def _launch(self, seg_start, seg_len_d, n, t_idx_d, t_frac_d, n_t, target_freqs, kmax,
                remove_dc_offset, out, want_float64):
        p, dev = self.p, self.device
        f_idx_dev, f_frac_dev, _ = self._freq_tables(target_freqs)
        n_f = int(p['num_freq_bins']) if target_freqs is None else len(target_freqs)
        if out is None:
            out = torch.empty(n, n_f, n_t, dtype=torch.float32, device=dev)
        out64 = torch.empty(n, n_f, n_t, dtype=torch.float64, device=dev) if want_float64 else None
        normalize = bool(p.get('within_syll_normalize', False))
        if normalize and out64 is None:
            out64 = torch.empty(n, n_f, n_t, dtype=torch.float64, device=dev)
        call("ava_b200_get_spec_batch", self.audio_dev.data_ptr(), int(self.is_f32),
             seg_start.data_ptr(), seg_len_d.data_ptr(), n, self.nperseg, self.noverlap,
             1 if remove_dc_offset else 0, self.window_dev.data_ptr(), self.scale,
             t_idx_d.data_ptr(), t_frac_d.data_ptr(), n_t, f_idx_dev.data_ptr(),
             f_frac_dev.data_ptr(), n_f, kmax + 1, float(p['spec_min_val']), float(p['spec_max_val']),
             out.data_ptr(), out64.data_ptr() if out64 is not None else None,
             torch.cuda.current_stream().cuda_stream)
        if normalize:
            self._normalize(out64, out)
        return (out, out64) if want_float64 else out

    def _normalize(self, out64, out32):
        """within_syll_normalize (ava/preprocessing/utils.py:106-109) on the device, in float64."""
        n = out64.shape[0]
        call("ava_b200_quantile_normalize", out64.data_ptr(), out32.data_ptr() if out32 is not None else None,
             n, int(out64.shape[1] * out64.shape[2]), float(self.p['normalize_quantile']),
             torch.cuda.current_stream().cuda_stream)

    def specs(self, file_index, t1, t2, target_times, target_freqs=None, remove_dc_offset=True,
              out=None, want_float64=False):
        """Spectrograms of n segments [t1_i, t2_i] of files file_index_i with explicit
        target times [n, n_t].  Returns fp32 [n, n_f, n_t] on the device (and the float64
        version if want_float64).  Mirrors ava/preprocessing/utils.py:59-104 per segment."""
        file_index = np.asarray(file_index, dtype=np.int64)
        t1 = np.asarray(t1, dtype=np.float64)
        t2 = np.asarray(t2, dtype=np.float64)
        tt = np.asarray(target_times, dtype=np.float64)
        n, n_t = tt.shape
        seg_start, seg_len, K, kmax, base = self._segments(file_index, t1, t2)
        t_idx, t_frac = bracket(np.maximum(0.0, t1), base, K, tt)
        dev = self.device
        seg_start = torch.from_numpy(seg_start).to(dev)
        seg_len_d = torch.from_numpy(seg_len).to(dev)
        t_idx_d = torch.from_numpy(np.ascontiguousarray(t_idx)).to(dev)
        t_frac_d = torch.from_numpy(np.ascontiguousarray(t_frac)).to(dev)
        # keep the argument tensors alive until the kernel has consumed them
        self._keepalive = (seg_start, seg_len_d, t_idx_d, t_frac_d)
        return self._launch(seg_start, seg_len_d, n, t_idx_d, t_frac_d, n_t, target_freqs, kmax,
                            remove_dc_offset, out, want_float64)

    def _pinned_slot(self, m):
        """A pinned [5, m] int64 staging buffer from a ring of 4 per batch size; a slot is
        reused only after the upload that last read it has completed."""
        if m not in self._pin_rings and len(self._pin_rings) >= 8:
            # many distinct batch sizes (silence-rejection retries): keep the staging rings of
            # the 8 most recent sizes; torch's pinned-memory allocator defers the reuse of a
            # dropped buffer until the copies reading it have completed
            self._pin_rings.pop(next(iter(self._pin_rings)))
        ring = self._pin_rings.setdefault(m, {"bufs": [], "events": [], "next": 0})
        if len(ring["bufs"]) < 4:
            ring["bufs"].append(torch.zeros(5, m, dtype=torch.int64).pin_memory())
            ring["events"].append(None)
            slot = len(ring["bufs"]) - 1
        else:
            slot = ring["next"]
            ring["next"] = (slot + 1) % 4
            if ring["events"][slot] is not None:
                ring["events"][slot].synchronize()
        return ring["bufs"][slot], ring, slot

    def specs_linspace(self, file_index, t1, t2, tstart, tstop, n_t=None, remove_dc_offset=True,
                       out=None, want_float64=False):
        """Same as ``specs`` with ``target_times[i] = np.linspace(tstart_i, tstop_i, n_t)`` (the
        fixed-window sampler, ava/models/window_vae_dataset.py:231-235), the [n, n_t] tables
        being built on the device by ``ava_b200_window_time_tables`` -- the same float64
        operations as ``bracket``, bit for bit -- so that the host only handles O(n) numbers
        per batch: one packed upload of 40 bytes per window."""
        file_index = np.asarray(file_index, dtype=np.int64)
        t1 = np.asarray(t1, dtype=np.float64)
        t2 = np.asarray(t2, dtype=np.float64)
        tstart = np.asarray(tstart, dtype=np.float64)
        tstop = np.asarray(tstop, dtype=np.float64)
        n = len(file_index)
        n_t = int(self.p['num_time_bins']) if n_t is None else int(n_t)
        if n_t < 2 or n == 0 or np.any(tstop == tstart):
            # numpy's linspace switches formula when a step is zero; keep one code path for that
            tt = np.linspace(tstart, tstop, n_t, axis=-1).reshape(n, n_t)
            return self.specs(file_index, t1, t2, tt, remove_dc_offset=remove_dc_offset, out=out,
                              want_float64=want_float64)
        seg_start, seg_len, K, kmax, base = self._segments(file_index, t1, t2)
        dev = self.device
        # one upload: rows [seg_start i64 | grid0 f64 | tstart f64 | tstop f64 | seg_len, K i32]
        m = n + (n & 1)
        pinned, slot_ring, slot = self._pinned_slot(m)
        packed = pinned.numpy()
        packed[0, :n] = seg_start
        packed[1, :n] = np.maximum(0.0, t1).view(np.int64)
        packed[2, :n] = np.ascontiguousarray(tstart).view(np.int64)
        packed[3, :n] = np.ascontiguousarray(tstop).view(np.int64)
        tail = packed[4].view(np.int32)                          # 2m int32 slots
        tail[:n] = seg_len
        tail[m:m + n] = K
        # pinned staging + async copy: the host never waits behind the train step queued on
        # this stream (a pageable upload of more than 64 KB would)
        packed_d = pinned.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        slot_ring["events"][slot] = ev
        if self._base_cache is None or self._base_cache[0] < kmax:
            # frame times do not depend on the batch: element k of a longer table is the same
            # number, so a cached table of at least kmax entries serves every batch
            kcap = max(kmax, 64)
            b = np.arange(self.nperseg / 2, self.nperseg / 2 + kcap * self.hop, self.hop) / float(self.fs)
            b = b - (self.nperseg / 2) / self.fs
            assert np.array_equal(b[:kmax], base)
            self._base_cache = (kcap, torch.from_numpy(b).to(dev))
        base_d = self._base_cache[1]
        tail_d = packed_d[4].view(torch.int32)
        seg_len_d, K_d = tail_d[:n], tail_d[m:m + n]
        t_idx_d = torch.empty(n, n_t, dtype=torch.int32, device=dev)
        t_frac_d = torch.empty(n, n_t, dtype=torch.float64, device=dev)
        call("ava_b200_window_time_tables", packed_d[1].data_ptr(), K_d.data_ptr(), base_d.data_ptr(), kmax,
             packed_d[2].data_ptr(), packed_d[3].data_ptr(), n, n_t, t_idx_d.data_ptr(),
             t_frac_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        self._keepalive = (packed_d, seg_len_d, K_d, t_idx_d, t_frac_d)
        return self._launch(packed_d[0], seg_len_d, n, t_idx_d, t_frac_d, n_t, None, kmax,
                            remove_dc_offset, out, want_float64)


def get_spec(t1, t2, audio, p, fs=32000, target_freqs=None, target_times=None,
             fill_value=-1 / EPSILON, max_dur=None, remove_dc_offset=True):
    """
    Norm, scale, threshold, stretch, and resize a Short Time Fourier Transform.

    Same parameters and return value as the reference (ava/preprocessing/utils.py:18-110):
    ``(spec, flag)`` with ``spec`` a float64 ``[num_freq_bins, num_time_bins]`` numpy array.
    This single-segment form uploads the segment and downloads the result on every call; the
    datasets use ``SpecEngine`` instead, which keeps audio and spectrograms on the device.
    """
    if max_dur is None:
        max_dur = p['max_dur']
    if t2 - t1 > max_dur + 1e-4:
        message = "Found segment longer than max_dur: " + str(t2 - t1) + \
            "s, max_dur = " + str(max_dur) + "s"
        warnings.warn(message)
    s1, s2 = int(round(t1 * fs)), int(round(t2 * fs))
    assert s1 < s2, "s1: " + str(s1) + " s2: " + str(s2) + " t1: " + str(t1) + \
        " t2: " + str(t2)
    temp = min(len(audio), s2) - max(0, s1)
    if temp < p['nperseg'] or s2 <= 0 or s1 >= len(audio):
        return np.zeros((p['num_freq_bins'], p['num_time_bins'])), True
    audio = np.asarray(audio)
    if target_times is None:
        duration = t2 - t1
        if p['time_stretch']:
            duration = np.sqrt(duration * max_dur)  # stretched duration
        shoulder = 0.5 * (max_dur - duration)
        target_times = np.linspace(t1 - shoulder, t2 + shoulder, p['num_time_bins'])
    target_times = np.asarray(target_times, dtype=np.float64)
    # upload only the segment (indices shifted so that the segment starts at sample 0)
    lo, hi = max(0, s1), min(len(audio), s2)
    seg = audio[lo:hi]
    eng = SpecEngine([seg], fs, p)
    fv = None
    if fill_value != -1 / EPSILON:
        # non-default fill value: out-of-range rows / columns take its normalised, clipped value
        fv = float(np.clip((fill_value - p['spec_min_val']) / (p['spec_max_val'] - p['spec_min_val']),
                           0.0, 1.0))
    spec32, spec64 = eng._specs_shifted(t1, t2, lo, hi - lo, target_times, target_freqs,
                                        remove_dc_offset, fill=fv)
    return spec64[0].cpu().numpy(), True


def _specs_shifted(self, t1, t2, seg_lo, seg_len, target_times, target_freqs, remove_dc_offset, fill=None):
    """One segment whose samples [seg_lo, seg_lo+seg_len) of the original file were uploaded
    as samples [0, seg_len): same tables as `specs`, indices relative to the upload."""
    p, fs = self.p, self.fs
    n_t = len(target_times)
    n_f = int(p['num_freq_bins']) if target_freqs is None else len(target_freqs)
    K = np.array([num_frames(seg_len, self.nperseg, self.hop)], dtype=np.int64)
    kmax = int(K[0])
    base = np.arange(self.nperseg / 2, self.nperseg / 2 + kmax * self.hop, self.hop) / float(fs)
    base = base - (self.nperseg / 2) / fs
    t_idx, t_frac = bracket(np.array([max(0.0, t1)]), base, K, target_times[None, :])
    f_idx_dev, f_frac_dev, f_idx = self._freq_tables(target_freqs)
    self._last_bad_t = t_idx[0] < 0
    self._last_bad_f = f_idx < 0
    dev = self.device
    seg_start = torch.zeros(1, dtype=torch.int64, device=dev)
    seg_len_d = torch.full((1,), int(seg_len), dtype=torch.int32, device=dev)
    t_idx_d = torch.from_numpy(np.ascontiguousarray(t_idx)).to(dev)
    t_frac_d = torch.from_numpy(np.ascontiguousarray(t_frac)).to(dev)
    out = torch.empty(1, n_f, n_t, dtype=torch.float32, device=dev)
    out64 = torch.empty(1, n_f, n_t, dtype=torch.float64, device=dev)
    call("ava_b200_get_spec_batch", self.audio_dev.data_ptr(), int(self.is_f32),
         seg_start.data_ptr(), seg_len_d.data_ptr(), 1, self.nperseg, self.noverlap,
         1 if remove_dc_offset else 0, self.window_dev.data_ptr(), self.scale, t_idx_d.data_ptr(),
         t_frac_d.data_ptr(), n_t, f_idx_dev.data_ptr(), f_frac_dev.data_ptr(), n_f, kmax + 1,
         float(p['spec_min_val']), float(p['spec_max_val']), out.data_ptr(), out64.data_ptr(),
         torch.cuda.current_stream().cuda_stream)
    if fill is not None:
        bad_t = torch.from_numpy(self._last_bad_t).to(dev)
        bad_f = torch.from_numpy(self._last_bad_f).to(dev)
        for o in (out, out64):
            o[0][:, bad_t] = fill
            o[0][bad_f, :] = fill
    if p.get('within_syll_normalize', False):
        self._normalize(out64, out)
    torch.cuda.current_stream().synchronize()
    return out, out64


SpecEngine._specs_shifted = _specs_shifted


if __name__ == '__main__':
    pass
