"""
Methods for feeding randomly sampled spectrogram data to the shotgun VAE: drop-in for
``ava.models.window_vae_dataset``.

Sampling stays on the host with numpy's legacy global RNG, drawing exactly the doubles the
reference draws in the same order (file, ROI, onset per candidate window;
ava/models/window_vae_dataset.py:215-231), so window indices/onsets are bit-exact under a
seed.  The spectrograms of a whole batch of windows are then produced by ONE launch of the
GPU ``get_spec`` kernel from audio resident in HBM (preprocessing/utils.py::SpecEngine).

``get_fixed_window_data_loaders`` returns ``WindowBatchLoader`` objects instead of
fork-based torch DataLoaders: iterable, with ``.dataset`` and ``len()``, yielding
``[b,128,128]`` fp32 CUDA tensors.
"""
import os
import warnings

import numpy as np
import torch
from torch.utils.data import Dataset

from ..preprocessing.utils import SpecEngine
from ..preprocessing.utils import get_spec as _builtin_get_spec
from .utils import _get_wavs_from_dir, _require_h5py, numpy_to_tensor  # noqa: F401

DEFAULT_WARP_PARAMS = {
    'n_knots': 0,  # number of pieces minus one in the piecwise linear warp
    'warp_reg_scale': 1e-2,  # penalizes distance of warp to identity line
    'smoothness_reg_scale': 1e-1,  # penalizes L2 norm of warp second derivatives
    'l2_reg_scale': 1e-7,  # penalizes L2 norm of warping template
}
"""Default time-warping parameters sent to affinewarp"""

EPSILON = 1e-9


def _foreign_get_spec(p):
    """The user's spectrogram function if ``p['get_spec']`` (the reference's plugin point,
    ava/models/window_vae_dataset.py:233-235,629-631) is anything other than this package's or
    the reference's own ``get_spec``; None if the batched GPU engine computes the same thing."""
    fn = p.get('get_spec', None) if isinstance(p, dict) else None
    if fn is None or fn is _builtin_get_spec:
        return None
    if getattr(fn, "__module__", "") == "ava.preprocessing.utils" and getattr(fn, "__name__", "") == "get_spec":
        return None
    return fn


def _read_wav(fn):
    from scipy.io import wavfile
    from scipy.io.wavfile import WavFileWarning
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", category=WavFileWarning)
        return wavfile.read(fn)


def get_window_partition(audio_dirs, roi_dirs, split=0.8, shuffle=True,
                         exclude_empty_roi_files=True):
    """Get a train/test split for fixed-duration shotgun VAE
    (ava/models/window_vae_dataset.py:40-99)."""
    assert(split > 0.0 and split <= 1.0)
    audio_filenames, roi_filenames = [], []
    for audio_dir, roi_dir in zip(audio_dirs, roi_dirs):
        temp_wavs = _get_wavs_from_dir(audio_dir)
        temp_rois = [os.path.join(roi_dir, os.path.split(i)[-1][:-4] + '.txt')
                     for i in temp_wavs]
        if exclude_empty_roi_files:
            for i in reversed(range(len(temp_wavs))):
                segs = np.loadtxt(temp_rois[i])
                if len(segs) == 0:
                    del temp_wavs[i]
                    del temp_rois[i]
        audio_filenames += temp_wavs
        roi_filenames += temp_rois
    # Reproducibly shuffle.
    audio_filenames = np.array(audio_filenames)
    roi_filenames = np.array(roi_filenames)
    perm = np.argsort(audio_filenames)
    audio_filenames, roi_filenames = audio_filenames[perm], roi_filenames[perm]
    if shuffle:
        np.random.seed(42)
        perm = np.random.permutation(len(audio_filenames))
        audio_filenames = audio_filenames[perm]
        roi_filenames = roi_filenames[perm]
        np.random.seed(None)
    i = int(round(split * len(audio_filenames)))
    return {
        'train': {'audio': audio_filenames[:i], 'rois': roi_filenames[:i]},
        'test': {'audio': audio_filenames[i:], 'rois': roi_filenames[i:]},
    }


class WindowBatchLoader:
    """Iterable of device batches of freshly sampled windows (one epoch =
    ``ceil(len(dataset) / batch_size)`` batches, like DataLoader over the reference's
    arbitrary-length dataset)."""

    def __init__(self, dataset, batch_size=64):
        self.dataset = dataset
        self.batch_size = batch_size

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        for start in range(0, n, self.batch_size):
            b = min(self.batch_size, n - start)
            yield self.dataset.sample_batch(b)


def get_fixed_window_data_loaders(partition, p, batch_size=64, shuffle=(True, False),
                                  num_workers=4, min_spec_val=None):
    """Get loaders for training and testing: fixed-duration shotgun VAE
    (ava/models/window_vae_dataset.py:102-135).  `shuffle`/`num_workers` are accepted for
    compatibility (every item is a fresh random window; no worker processes)."""
    train_dataset = FixedWindowDataset(partition['train']['audio'], partition['train']['rois'], p,
                                       transform=numpy_to_tensor, min_spec_val=min_spec_val)
    train_dataloader = WindowBatchLoader(train_dataset, batch_size=batch_size)
    # The reference tests `not partition['test']` on a dict (always truthy) and then
    # crashes on an empty split; be tolerant instead.
    if partition.get('test') is None or len(partition['test']['audio']) == 0:
        return {'train': train_dataloader, 'test': None}
    test_dataset = FixedWindowDataset(partition['test']['audio'], partition['test']['rois'], p,
                                      transform=numpy_to_tensor, min_spec_val=min_spec_val)
    test_dataloader = WindowBatchLoader(test_dataset, batch_size=batch_size)
    return {'train': train_dataloader, 'test': test_dataloader}


class FixedWindowDataset(Dataset):
    """Chunks of animal vocalization (ava/models/window_vae_dataset.py:139-256)."""

    def __init__(self, audio_filenames, roi_filenames, p, transform=None, dataset_length=2048,
                 min_spec_val=None, audio=None, fs=None, rois=None, device=None):
        """`audio`/`fs`/`rois` (addition): pass in-memory data instead of reading the files
        (used by benchmarks on synthetic corpora)."""
        # NOTE (quirk F11, kept): audio filenames are sorted, ROI filenames are not.
        self.filenames = np.array(sorted(audio_filenames))
        if audio is None:
            self.audio = [_read_wav(fn)[1] for fn in self.filenames]
            self.fs = _read_wav(audio_filenames[0])[0]
        else:
            self.audio, self.fs = list(audio), fs
        self.roi_filenames = roi_filenames
        self.dataset_length = dataset_length
        self.min_spec_val = min_spec_val
        self.p = p
        self.rois = [np.loadtxt(i, ndmin=2) for i in roi_filenames] if rois is None else \
            [np.asarray(r, dtype=np.float64).reshape(-1, 2) for r in rois]
        self.file_weights = np.array([np.sum(np.diff(i)) for i in self.rois])
        self.file_weights /= np.sum(self.file_weights)
        self.roi_weights = []
        for i in range(len(self.rois)):
            temp = np.diff(self.rois[i]).flatten()
            self.roi_weights.append(temp / np.sum(temp))
        self.transform = transform
        # numpy's choice(arange(n), p=w) == searchsorted(cdf, one uniform double, 'right')
        self._file_cdf = self._cdf(self.file_weights)
        self._roi_cdf = [self._cdf(w) for w in self.roi_weights]
        self._device = device
        self._engine_obj = None   # created on first use (needs a CUDA device)

    @property
    def _engine(self):
        if self._engine_obj is None:
            self._engine_obj = SpecEngine(self.audio, self.fs, self.p, device=self._device)
        return self._engine_obj

    @staticmethod
    def _cdf(w):
        cdf = np.cumsum(np.asarray(w, dtype=np.float64))
        cdf /= cdf[-1]
        return cdf

    def __len__(self):
        """NOTE: length is arbitrary"""
        return self.dataset_length

    def _draw(self, m):
        """m candidate windows from numpy's global legacy stream: per candidate exactly
        three doubles in the order (file, roi, onset) -- window_vae_dataset.py:219-228."""
        u = np.random.random_sample(3 * m)
        files = np.searchsorted(self._file_cdf, u[0::3], side='right')
        onsets = np.empty(m, dtype=np.float64)
        wl = self.p['window_length']
        for f in np.unique(files):
            sel = np.nonzero(files == f)[0]
            ri = np.searchsorted(self._roi_cdf[f], u[1::3][sel], side='right')
            roi = self.rois[f][ri]
            onsets[sel] = roi[:, 0] + (roi[:, 1] - roi[:, 0] - wl) * u[2::3][sel]
        return files.astype(np.int64), onsets

    def _specs(self, files, onsets, shoulder):
        wl = self.p['window_length']
        offsets = onsets + wl
        plugin = _foreign_get_spec(self.p)
        if plugin is not None:
            # a user-supplied spectrogram function: call it exactly as the reference does, one
            # window at a time on the host (window_vae_dataset.py:229-235); windows it flags as
            # invalid come back as NaN rows and are dropped by `sample`
            specs = []
            for f, on, off in zip(files, onsets, offsets):
                tt = np.linspace(on, off, self.p['num_time_bins'])
                spec, flag = plugin(max(0.0, on - shoulder), off + shoulder, self.audio[f], self.p,
                                    fs=self.fs, target_times=tt)
                spec = np.asarray(spec, dtype=np.float32)
                specs.append(spec if flag else np.full_like(spec, np.nan))
            dev = self._device if self._device is not None else torch.device("cuda", torch.cuda.current_device())
            return torch.from_numpy(np.stack(specs)).to(dev)
        # target_times = np.linspace(onset, offset, num_time_bins), built on the device
        return self._engine.specs_linspace(files, np.maximum(0.0, onsets - shoulder), offsets + shoulder,
                                           onsets, offsets)

    def sample(self, n, seed=None, shoulder=0.05):
        """n accepted windows: (specs [n,128,128] fp32 on the device, file_indices, onsets,
        offsets).  The reference's retry loop (window_vae_dataset.py:216-249) examines
        candidate windows in stream order -- three doubles each -- and keeps the ones that
        pass the silence test; the k-th kept candidate is item k.  Here candidates are drawn
        and evaluated a chunk at a time, which visits the same candidates in the same order
        (it may draw past the n-th acceptance, which is unobservable: the stream is reseeded
        from entropy afterwards, as in the reference)."""
        np.random.seed(seed)
        got_specs, got_f, got_on = [], [], []
        need = n
        while need > 0:
            files, onsets = self._draw(need)
            specs = self._specs(files, onsets, shoulder)
            keep = None
            if _foreign_get_spec(self.p) is not None:
                keep = ~torch.isnan(specs[:, 0, 0])           # the plugin's `flag` (:236-237)
            if self.min_spec_val is not None:
                loud = specs.amax(dim=(1, 2)) >= self.min_spec_val
                keep = loud if keep is None else (keep & loud)
            if keep is not None:
                keep_h = keep.cpu().numpy()
                if not keep_h.all():
                    specs, files, onsets = specs[keep], files[keep_h], onsets[keep_h]
            got_specs.append(specs)
            got_f.append(files)
            got_on.append(onsets)
            need -= len(files)
        np.random.seed(None)
        specs = torch.cat(got_specs) if len(got_specs) > 1 else got_specs[0]
        files = np.concatenate(got_f)
        onsets = np.concatenate(got_on)
        return specs, files, onsets, onsets + self.p['window_length']

    def sample_batch(self, n):
        """One training batch of fresh windows (seed=None, as DataLoader items are)."""
        return self.sample(n, seed=None)[0]

    def __getitem__(self, index, seed=None, shoulder=0.05, return_seg_info=False):
        """Get spectrograms (window_vae_dataset.py:189-256).  Items are fp32 device tensors."""
        single_index = False
        try:
            n = len(list(iter(index)))
        except TypeError:
            n, single_index = 1, True
        specs, files, onsets, offsets = self.sample(n, seed=seed, shoulder=shoulder)
        specs = list(specs.unbind(0))
        file_indices = [int(f) for f in files]
        onsets, offsets = [float(o) for o in onsets], [float(o) for o in offsets]
        if return_seg_info:
            if single_index:
                return specs[0], file_indices[0], onsets[0], offsets[0]
            return specs, file_indices, onsets, offsets
        if single_index:
            return specs[0]
        return specs

    def write_hdf5_files(self, save_dir, num_files=500, sylls_per_file=100):
        """Write hdf5 files containing spectrograms of random audio chunks
        (window_vae_dataset.py:259-293)."""
        h5py = _require_h5py()
        if not os.path.exists(save_dir):
            os.mkdir(save_dir)
        for write_file_num in range(num_files):
            specs, file_indices, _, _ = self.__getitem__(np.arange(sylls_per_file),
                                                         seed=write_file_num, return_seg_info=True)
            specs = np.array([spec.detach().cpu().numpy() for spec in specs])
            filenames = np.array([self.filenames[i] for i in file_indices])
            fn = "syllables_" + str(write_file_num).zfill(4) + '.hdf5'
            fn = os.path.join(save_dir, fn)
            with h5py.File(fn, "w") as f:
                f.create_dataset('specs', data=specs)
                f.create_dataset('audio_filenames', data=filenames.astype('S'))


def get_warped_window_data_loaders(audio_dirs, p, batch_size=64, num_workers=4, load_warp=False,
                                   warp_fn=None, warp_params={}, warp_type='spectrogram'):
    """Get loaders for the warped shotgun VAE (window_vae_dataset.py:297-355).  Fitting a
    warp needs the third-party `affinewarp` package (out of scope); the
    'null' warp and pre-fitted knots loaded from `warp_fn` are supported."""
    assert type(p) == type({})
    assert warp_type in ['amplitude', 'spectrogram', 'null']
    audio_fns = []
    for audio_dir in audio_dirs:
        audio_fns += _get_wavs_from_dir(audio_dir)
    dataset = WarpedWindowDataset(audio_fns, p, transform=numpy_to_tensor, load_warp=load_warp,
                                  warp_fn=warp_fn, warp_params=warp_params, warp_type=warp_type)
    dataloader = WindowBatchLoader(dataset, batch_size=batch_size)
    return {'train': dataloader, 'test': dataloader}


class WarpedWindowDataset(Dataset):
    """Time-warped chunks of animal vocalization (window_vae_dataset.py:359-701): sampling
    and inverse-warp path; every item is one whole-file `get_spec` with warped target times,
    batched on the GPU."""

    def __init__(self, audio_filenames, p, transform=None, dataset_length=2048, load_warp=False,
                 save_warp=True, start_q=-0.1, stop_q=1.1, warp_fn=None, warp_params={},
                 warp_type='spectrogram', audio=None, fs=None, template_dur=None, device=None):
        assert type(p) == type({})
        assert warp_type in ['amplitude', 'spectrogram', 'null']
        self.audio_filenames = sorted(audio_filenames)
        if audio is None:
            self.audio = [_read_wav(fn)[1] for fn in self.audio_filenames]
            self.fs = _read_wav(self.audio_filenames[0])[0]
        else:
            self.audio, self.fs = list(audio), fs
        self.dataset_length = dataset_length
        self.p = p
        self.transform = transform
        self.start_q = start_q
        self.stop_q = stop_q
        self.warp_fn = warp_fn
        self.warp_params = {**DEFAULT_WARP_PARAMS, **warp_params}
        if warp_type == 'null':
            knots = np.zeros((len(self.audio), 2))
            knots[:, 1] = 1.0
            self.x_knots = knots
            self.y_knots = np.copy(knots)
            if template_dur is None:
                # template duration as the reference derives it: (#STFT frames) * dt of the
                # shortest file (models/utils.py:337-418)
                hop = p['nperseg'] - p['noverlap']
                n = min(len(a) for a in self.audio)
                frames = -(-n // hop) + 1
                template_dur = frames * (hop / self.fs)
            self.template_dur = template_dur
        elif load_warp and warp_fn is not None:
            data = np.load(warp_fn, allow_pickle=True).item()
            self.x_knots = data['x_knots']
            self.y_knots = data['y_knots']
            self.template_dur = data['template_dur']
            temp_fns = list(data['audio_filenames'])
            assert len(temp_fns) >= len(self.audio_filenames)
            if len(temp_fns) != len(self.audio_filenames):
                perm = np.array([temp_fns.index(fn) for fn in self.audio_filenames], dtype='int')
                self.x_knots = self.x_knots[perm]
                self.y_knots = self.y_knots[perm]
            else:
                assert np.array_equal(temp_fns, self.audio_filenames), \
                    "Input filenames do not match saved filenames!"
        else:
            raise NotImplementedError(
                "fitting a time warp requires the third-party `affinewarp` package (not part of "
                "the hot path); use warp_type='null' or load_warp=True with a saved warp_fn")
        self.window_frac = self.p['window_length'] / self.template_dur
        self._device = device
        self._engine_obj = None

    @property
    def _engine(self):
        if self._engine_obj is None:
            self._engine_obj = SpecEngine(self.audio, self.fs, self.p, device=self._device)
        return self._engine_obj

    def __len__(self):
        """NOTE: length is arbitrary."""
        return self.dataset_length

    def _get_unwarped_times(self, y_vals, index):
        """Template (warped) quantile times -> empirical quantile times
        (window_vae_dataset.py:461-477): ``scipy.interpolate.interp1d(y_knots, x_knots,
        fill_value='extrapolate')``, i.e. piecewise-linear interpolation through the knots with
        linear extrapolation, in the arithmetic of the installed SciPy (1.18:
        ``_interpolate.py::interp1d._call_linear`` -- interval from ``searchsorted`` (left)
        clipped to [1, n-1], value = ((t-lo)/(hi-lo))*x_hi + ((hi-t)/(hi-lo))*x_lo), so the
        target times are bit-identical to the reference's."""
        x_knots, y_knots = self.x_knots[index], self.y_knots[index]
        hi = np.clip(np.searchsorted(y_knots, y_vals), 1, len(y_knots) - 1)
        lo = hi - 1
        y_lo, y_hi = y_knots[lo], y_knots[hi]
        return ((y_vals - y_lo) / (y_hi - y_lo)) * x_knots[hi] + ((y_hi - y_vals) / (y_hi - y_lo)) * x_knots[lo]

    def _draw(self, n, seed=None):
        """(file index, start quantile) of n items from numpy's global legacy stream, in the
        reference's order -- ``randint`` then ``rand`` per item (window_vae_dataset.py:613-617)
        -- and the [n, num_time_bins] un-warped target times (:618-624).  Only the two draws
        per item are a Python loop; the linspace / inverse-warp arithmetic runs on whole arrays
        with the same float64 operations per element."""
        np.random.seed(seed)
        files = np.empty(n, dtype=np.int64)
        u = np.empty(n, dtype=np.float64)
        n_files = len(self.audio)
        for i in range(n):
            files[i] = np.random.randint(n_files)
            u[i] = np.random.rand()
        np.random.seed(None)
        start_t = self.start_q + u * (self.stop_q - self.start_q - self.window_frac)
        stop_t = start_t + self.window_frac
        n_t = self.p['num_time_bins']
        t_vals = np.linspace(start_t, stop_t, n_t, axis=-1).reshape(n, n_t)
        tts = np.empty((n, n_t), dtype=np.float64)
        for f in np.unique(files):
            sel = np.nonzero(files == f)[0]
            tts[sel] = self._get_unwarped_times(t_vals[sel], f) * self.template_dur
        return files, tts

    def sample(self, n, seed=None):
        files, tts = self._draw(n, seed)
        plugin = _foreign_get_spec(self.p)
        if plugin is not None:
            # user-supplied spectrogram function, called as the reference does (:629-631)
            specs = [np.asarray(plugin(0.0, self.template_dur, self.audio[f], self.p, fs=self.fs,
                                       max_dur=None, target_times=tt)[0], dtype=np.float32)
                     for f, tt in zip(files, tts)]
            dev = self._device if self._device is not None else torch.device("cuda", torch.cuda.current_device())
            return torch.from_numpy(np.stack(specs)).to(dev), files
        t1 = np.zeros(n)
        t2 = np.full(n, self.template_dur)
        return self._engine.specs(files, t1, t2, tts), files

    def sample_batch(self, n):
        return self.sample(n)[0]

    def __getitem__(self, index, seed=None):
        single_index = False
        try:
            n = len(list(iter(index)))
        except TypeError:
            n, single_index = 1, True
        specs = list(self.sample(n, seed=seed)[0].unbind(0))
        return specs[0] if single_index else specs

    def get_specific_item(self, query_filename, quantile):
        """A specific window of birdsong as a numpy array (window_vae_dataset.py:643-670)."""
        file_index = self.audio_filenames.index(query_filename)
        start_t = self.start_q + quantile * (self.stop_q - self.start_q - self.window_frac)
        stop_t = start_t + self.window_frac
        t_vals = np.linspace(start_t, stop_t, self.p['num_time_bins'])
        target_ts = self._get_unwarped_times(t_vals, file_index) * self.template_dur
        _, s64 = self._engine.specs([file_index], [0.0], [self.template_dur], target_ts[None, :],
                                    want_float64=True)
        return s64[0].cpu().numpy()


if __name__ == '__main__':
    pass
