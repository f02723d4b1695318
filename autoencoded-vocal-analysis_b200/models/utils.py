"""
Loader glue: the hot-path subset of ``ava.models.utils``
(ava/models/utils.py:311-334, 421-461).  The warp hyper-parameter search and the
whole-file ``_get_spec`` helpers of that module are out of scope.
"""
import os

import numpy as np
import torch


def _require_h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:
        raise ImportError("reading/writing AVA's .hdf5 syllable files needs h5py, which is not "
                          "installed in this environment") from e


def read_specs(filename):
    """The ``specs`` array of one syllable file: AVA's HDF5 layout
    (ava/preprocessing/preprocess.py:87-88) or, for environments without h5py, an ``.npy``
    / ``.npz`` file holding the same array."""
    if filename.endswith(".npy"):
        return np.load(filename, mmap_mode="r")
    if filename.endswith(".npz"):
        return np.load(filename)["specs"]
    h5py = _require_h5py()
    with h5py.File(filename, 'r') as f:
        return f['specs'][:]


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


def read_field(filename, key):
    """One dataset of a syllable / projection file as a numpy array (HDF5, or the ``.npz``
    stand-in with the same keys)."""
    if filename.endswith(".npz"):
        with np.load(filename) as f:
            assert key in f.files, "Can't find field '" + key + "' in file '" + filename + "'!"
            return f[key]
    h5py = _require_h5py()
    with h5py.File(filename, 'r') as f:
        assert key in f, "Can't find field '" + key + "' in file '" + filename + "'!"
        return np.array(f[key])


def stored_fields(filename):
    """{dataset name: number of rows} of one file."""
    if filename.endswith(".npz"):
        with np.load(filename) as f:
            return {k: len(f[k]) for k in f.files}
    h5py = _require_h5py()
    with h5py.File(filename, 'r') as f:
        return {k: len(f[k]) for k in f.keys()}


def append_field(filename, key, data):
    """``h5py.File(filename, 'a').create_dataset(key, data=data)``
    (ava/data/data_container.py:482-483, 663-664): add one dataset to a file, creating the
    file if needed; like h5py, refuses to overwrite an existing dataset."""
    data = np.asarray(data)
    if filename.endswith(".npz"):
        fields = {}
        if os.path.exists(filename):
            with np.load(filename) as f:
                fields = {k: f[k] for k in f.files}
        if key in fields:
            raise ValueError("Unable to create dataset (name already exists): " + key)
        fields[key] = data
        np.savez(filename, **fields)
        return
    h5py = _require_h5py()
    with h5py.File(filename, 'a') as f:
        f.create_dataset(key, data=data)


def _get_sylls_per_file(partition):
    """Open a file and see how many syllables it has (ava/models/utils.py:311-334).
    Assumes all files referenced by `partition` hold the same number of syllables."""
    key = 'train' if len(partition['train']) > 0 else 'test'
    assert len(partition[key]) > 0
    filename = partition[key][0]  # Just grab the first file.
    return len(read_specs(filename))


def get_hdf5s_from_dir(dir):
    """Return a sorted list of all hdf5s in a directory (ava/models/utils.py:421-430).
    Without h5py (and only then) the ``.npz`` stand-ins this package writes are listed
    instead."""
    found = [os.path.join(dir, f) for f in sorted(os.listdir(dir)) if _is_hdf5_file(f)]
    if not found and not _have_h5py():
        found = [os.path.join(dir, f) for f in sorted(os.listdir(dir))
                 if len(f) > 4 and f[-4:] == '.npz']
    return found


def _get_wavs_from_dir(dir):
    """Return a sorted list of wave files from a directory."""
    return [os.path.join(dir, f) for f in sorted(os.listdir(dir)) if _is_wav_file(f)]


def _get_txts_from_dir(dir):
    """Return a sorted list of text files from a directory."""
    return [os.path.join(dir, f) for f in sorted(os.listdir(dir)) if _is_txt_file(f)]


def numpy_to_tensor(x):
    """Transform a numpy array into a torch.FloatTensor (ava/models/utils.py:444-446)."""
    return torch.from_numpy(np.asarray(x)).type(torch.FloatTensor)


def _is_hdf5_file(filename):
    """Is the given filename an hdf5 file?"""
    return len(filename) > 5 and filename[-5:] == '.hdf5'


def _is_wav_file(filename):
    """Is the given filename a wave file?"""
    return len(filename) > 4 and filename[-4:] == '.wav'


def _is_txt_file(filename):
    """Is the given filename a text file?"""
    return len(filename) > 4 and filename[-4:] == '.txt'
