"""
Compute and save syllable spectrograms -- the step in front of the syllable VAE
(mirror of ava/preprocessing/preprocess.py:23-150, 313-373 for the non-interactive part).

Same functions, arguments, file naming and on-disk fields as the reference
(`syllables_%04d.hdf5` with `specs` float64 [n,128,128], `onsets`, `offsets`,
`audio_filenames`), but all syllables of an audio file go through ONE batched launch of the
GPU `get_spec` kernel (`SpecEngine.specs`) instead of one scipy STFT + interpolation per
syllable.  The interactive `tune_syll_preprocessing_params` (matplotlib + input()) is out of
scope.

If ``p['get_spec']`` is a user-supplied function (the reference's plugin point,
preprocess.py:145-146) other than this package's / the reference's own ``get_spec``, it is
honoured and called per syllable, exactly as the reference does.
"""
import os
import warnings

import numpy as np
from scipy.io import wavfile
from scipy.io.wavfile import WavFileWarning

from .utils import EPSILON, SpecEngine, _inv_mel, _mel, get_spec

__all__ = ["process_sylls", "get_syll_specs", "get_audio_seg_filenames", "get_audio_filenames",
           "read_onsets_offsets_from_file", "is_audio_file"]


def _is_builtin_get_spec(fn):
    """True for this package's get_spec and for the reference's (same arithmetic)."""
    if fn is None or fn is get_spec:
        return True
    return getattr(fn, "__module__", "") == "ava.preprocessing.utils" and \
        getattr(fn, "__name__", "") == "get_spec"


def _write_batch(save_filename, syll_data, n, audio_dir):
    """One file of `n` syllables with the reference's four fields (preprocess.py:82-94).
    HDF5 when h5py is importable; otherwise the `.npz` stand-in the loaders of this package
    read (same keys)."""
    fields = {
        'onsets': np.array(syll_data['onsets'][:n]),
        'offsets': np.array(syll_data['offsets'][:n]),
        'specs': np.stack(syll_data['specs'][:n]),
        'audio_filenames': np.array([os.path.join(audio_dir, i) for i in
                                     syll_data['audio_filenames'][:n]]).astype('S'),
    }
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(save_filename, "w") as f:
            for key in ('onsets', 'offsets', 'specs', 'audio_filenames'):
                f.create_dataset(key, data=fields[key])
        return save_filename
    alt = save_filename[:-5] + ".npz" if save_filename.endswith(".hdf5") else save_filename + ".npz"
    np.savez(alt, **fields)
    return alt


def process_sylls(audio_dir, segment_dir, save_dir, p, shuffle=True, verbose=True):
    """
    Extract syllables from `audio_dir` and save to `save_dir`
    (ava/preprocessing/preprocess.py:23-105; same parameters).
    """
    if verbose:
        print("Processing audio files in", audio_dir)
    if not os.path.exists(save_dir):
        os.makedirs(save_dir)
    audio_filenames, seg_filenames = get_audio_seg_filenames(audio_dir, segment_dir, p)
    if shuffle:
        np.random.seed(42)
        perm = np.random.permutation(len(audio_filenames))
        np.random.seed(None)
        audio_filenames = np.array(audio_filenames)[perm]
        seg_filenames = np.array(seg_filenames)[perm]
    write_file_num = 0
    syll_data = {'specs': [], 'onsets': [], 'offsets': [], 'audio_filenames': []}
    sylls_per_file = p['sylls_per_file']
    written = []
    for audio_filename, seg_filename in zip(audio_filenames, seg_filenames):
        onsets, offsets = read_onsets_offsets_from_file(seg_filename, p)
        specs, good_sylls = get_syll_specs(onsets, offsets, audio_filename, p)
        onsets = [onsets[i] for i in good_sylls]
        offsets = [offsets[i] for i in good_sylls]
        syll_data['specs'] += specs
        syll_data['onsets'] += onsets
        syll_data['offsets'] += offsets
        syll_data['audio_filenames'] += len(onsets) * [os.path.split(audio_filename)[-1]]
        while len(syll_data['onsets']) >= sylls_per_file:
            save_filename = os.path.join(save_dir, "syllables_" + str(write_file_num).zfill(4) + '.hdf5')
            written.append(_write_batch(save_filename, syll_data, sylls_per_file, audio_dir))
            write_file_num += 1
            for key in syll_data:
                syll_data[key] = syll_data[key][sylls_per_file:]
            if p['max_num_syllables'] is not None and \
                    write_file_num * sylls_per_file >= p['max_num_syllables']:
                if verbose:
                    print("\tSaved max_num_syllables (" + str(p['max_num_syllables']) + "). Returning.")
                return written
    if verbose:
        print("\tDone.")
    return written


def get_syll_specs(onsets, offsets, audio_filename, p):
    """
    Return the spectrograms corresponding to `onsets` and `offsets`
    (ava/preprocessing/preprocess.py:108-150): ``(specs, valid_syllables)`` with ``specs`` a
    list of float64 ``[num_freq_bins, num_time_bins]`` arrays.
    """
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", category=WavFileWarning)
        fs, audio = wavfile.read(audio_filename)
    assert p['nperseg'] % 2 == 0 and p['nperseg'] > 2
    if p['mel']:
        target_freqs = np.linspace(_mel(p['min_freq']), _mel(p['max_freq']), p['num_freq_bins'])
        target_freqs = _inv_mel(target_freqs)
    else:
        target_freqs = np.linspace(p['min_freq'], p['max_freq'], p['num_freq_bins'])
    plugin = p.get('get_spec', None)
    if not _is_builtin_get_spec(plugin):
        specs, valid_syllables = [], []
        for i, t1, t2 in zip(range(len(onsets)), onsets, offsets):
            spec, valid = plugin(t1, t2, audio, p, fs, target_freqs=target_freqs)
            if valid:
                valid_syllables.append(i)
                specs.append(spec)
        return specs, valid_syllables
    return _syll_specs_batched(np.asarray(onsets, dtype=np.float64),
                               np.asarray(offsets, dtype=np.float64), audio, fs, p, target_freqs)


def _syll_specs_batched(onsets, offsets, audio, fs, p, target_freqs):
    """All syllables of one file in one launch; per-syllable semantics of get_spec
    (ava/preprocessing/utils.py:59-110): over-long warning, short segments -> zeros (still
    'valid'), time-stretched target times, optional within-syllable normalisation."""
    n = len(onsets)
    if n == 0:
        return [], []
    max_dur = p['max_dur']
    for t1, t2 in zip(onsets, offsets):
        if t2 - t1 > max_dur + 1e-4:
            warnings.warn("Found segment longer than max_dur: " + str(t2 - t1) + "s, max_dur = " +
                          str(max_dur) + "s")
    duration = offsets - onsets
    if p['time_stretch']:
        duration = np.sqrt(duration * max_dur)
    shoulder = 0.5 * (max_dur - duration)
    target_times = np.stack([np.linspace(t1 - s, t2 + s, p['num_time_bins'])
                             for t1, t2, s in zip(onsets, offsets, shoulder)])
    audio = np.asarray(audio)
    if audio.ndim > 1:
        raise ValueError("mono audio expected (the reference slices a 1-D array)")
    eng = SpecEngine([audio], fs, p)
    # (within_syll_normalize, when set, is applied on the device by the engine)
    _, spec64 = eng.specs(np.zeros(n, dtype=np.int64), onsets, offsets, target_times,
                          target_freqs=target_freqs, want_float64=True)
    out = spec64.cpu().numpy()
    return [out[i] for i in range(n)], list(range(n))


def get_audio_seg_filenames(audio_dir, segment_dir, p):
    """Return lists of sorted filenames (preprocess.py:313-326)."""
    temp_filenames = [i for i in sorted(os.listdir(audio_dir)) if is_audio_file(i)]
    audio_filenames = [os.path.join(audio_dir, i) for i in temp_filenames]
    temp_filenames = [i[:-4] + '.txt' for i in temp_filenames]
    seg_filenames = [os.path.join(segment_dir, i) for i in temp_filenames]
    for i in range(len(seg_filenames) - 1, -1, -1):
        if not os.path.exists(seg_filenames[i]):
            del seg_filenames[i]
            del audio_filenames[i]
    return audio_filenames, seg_filenames


def get_audio_filenames(audio_dir):
    """Return a list of sorted audio files (preprocess.py:329-333)."""
    return [os.path.join(audio_dir, i) for i in sorted(os.listdir(audio_dir)) if is_audio_file(i)]


def read_onsets_offsets_from_file(txt_filename, p):
    """Read a text file to collect onsets and offsets (preprocess.py:336-348)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        segs = np.loadtxt(txt_filename)
    assert segs.size % 2 == 0, "Incorrect formatting: " + txt_filename
    segs = segs.reshape(-1, 2)
    return segs[:, 0], segs[:, 1]


def is_audio_file(fn):
    """Return whether the given filename is an audio filename (preprocess.py:366-368)."""
    return len(fn) >= 4 and fn[-4:] == '.wav'
