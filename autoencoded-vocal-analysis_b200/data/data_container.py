"""
DataContainer: the latent-mean / PCA path of ``ava.data.data_container``.

The reference's ``DataContainer`` (ava/data/data_container.py:126-717) ties audio, segment,
spectrogram, feature and projection directories together and computes fields on request.
Its only calls into the VAE hot path are

* ``_make_latent_means`` (data_container.py:435-487): ``VAE.get_latent`` over every
  syllable file of every ``spec_dir`` -- one ``h5py.File`` open per syllable through 4 forked
  DataLoader workers, one device->host copy per batch -- then one ``latent_means`` dataset per
  file written into the matching ``projection_dir``;
* ``_make_latent_mean_pca_projection`` (data_container.py:538-551): scikit-learn
  ``PCA(n_components=2, copy=False, random_state=42).fit_transform``.

Here every syllable file is read once and uploaded whole (one file resident at a time, so a
corpus larger than HBM streams through), ``get_latent`` runs on the device with a single
download of the latent means, and the PCA is a handful of kernels (csrc/pca.cu).  The
request / make / read / write protocol, field names, file naming and on-disk datasets are the
reference's, so projections written here are read by the reference and vice versa.  The other
field families (segments, MUPET / DeepSqueak / SAP feature tables, UMAP) never touch the VAE
and are outside the hot path: requesting them raises NotImplementedError.

Storage: HDF5 through h5py when it is importable; otherwise the ``.npz`` stand-in files this
package's ``process_sylls`` writes (same keys).
"""
import os

import numpy as np
import torch

from .. import _lib
from ..models.utils import (append_field, get_hdf5s_from_dir, read_field, stored_fields)
from ..models.vae import VAE
from ..models.vae_dataset import get_syllable_data_loaders, get_syllable_partition, shard_files

AUDIO_FIELDS = ['audio']
FILENAME_FIELDS = ['sap_time']
SEGMENT_FIELDS = ['segments', 'segment_audio']
PROJECTION_FIELDS = ['latent_means', 'latent_mean_pca', 'latent_mean_umap']
SPEC_FIELDS = ['specs', 'onsets', 'offsets', 'audio_filenames']
MUPET_FIELDS = ['syllable_number', 'syllable_start_time', 'syllable_end_time',
                'inter-syllable_interval', 'syllable_duration', 'starting_frequency',
                'final_frequency', 'minimum_frequency', 'maximum_frequency',
                'mean_frequency', 'frequency_bandwidth', 'total_syllable_energy',
                'peak_syllable_amplitude', 'cluster']
DEEPSQUEAK_FIELDS = ['id', 'label', 'accepted', 'score', 'begin_time',
                     'end_time', 'call_length', 'principal_frequency', 'low_freq', 'high_freq',
                     'delta_freq', 'frequency_standard_deviation', 'slope', 'sinuosity',
                     'mean_power', 'tonality']
SAP_FIELDS = ['syllable_duration_sap', 'syllable_start', 'mean_amplitude',
              'mean_pitch', 'mean_FM', 'mean_AM2', 'mean_entropy', 'mean_pitch_goodness',
              'mean_mean_freq', 'pitch_variance', 'FM_variance', 'entropy_variance',
              'pitch_goodness_variance', 'mean_freq_variance', 'AM_variance']
ALL_FIELDS = AUDIO_FIELDS + FILENAME_FIELDS + SEGMENT_FIELDS + \
    PROJECTION_FIELDS + SPEC_FIELDS + MUPET_FIELDS + DEEPSQUEAK_FIELDS + \
    SAP_FIELDS
"""All fields that can be requested by a DataContainer object."""

HOT_PATH_FIELDS = ['latent_means', 'latent_mean_pca'] + SPEC_FIELDS
"""The fields this package computes or reads."""


class LatentPCA:
    """``sklearn.decomposition.PCA(n_components)`` restricted to what the reference uses
    (``fit_transform`` on the latent means, data_container.py:543-546), on the device.

    Attributes after ``fit_transform`` (numpy, named as in scikit-learn): ``mean_``,
    ``components_`` [K,D], ``explained_variance_`` [K], ``explained_variance_ratio_`` [K],
    ``n_samples_``.  Component signs follow scikit-learn's ``svd_flip`` (largest-magnitude
    entry of each component positive), so embeddings agree in sign with the reference's."""

    def __init__(self, n_components=2, device=None):
        self.n_components = n_components
        self.device = device

    def _device(self, x):
        if torch.is_tensor(x) and x.is_cuda:
            return x.device
        if self.device is not None:
            return torch.device(self.device)
        if not torch.cuda.is_available():
            raise _lib.AvaB200Error("LatentPCA needs a CUDA device (there is no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())

    def fit_transform_device(self, x):
        """x: [N,D] numpy array or tensor (float64 or float32).  Returns the [N,K] float64
        embedding as a tensor on the device."""
        dev = self._device(x)
        if not torch.is_tensor(x):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        x = x.to(dev).contiguous()
        if x.dim() != 2:
            raise ValueError("expected a [N,D] array, got shape %s" % (tuple(x.shape),))
        N, D = x.shape
        K = self.n_components
        if not (1 <= K <= D):
            raise ValueError("n_components=%r must be between 1 and D=%d" % (K, D))
        ws_bytes = int(_lib.lib().ava_b200_pca_ws_bytes(D))
        if ws_bytes < 0:
            raise _lib.AvaB200Error("LatentPCA supports at most 64 latent dimensions, got %d" % D)
        with torch.cuda.device(dev):
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            mean = torch.empty(D, dtype=torch.float64, device=dev)
            cov = torch.empty(D, D, dtype=torch.float64, device=dev)
            evals = torch.empty(D, dtype=torch.float64, device=dev)
            comps = torch.empty(D, D, dtype=torch.float64, device=dev)
            out = torch.empty(N, K, dtype=torch.float64, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            is_f32 = 1 if x.dtype == torch.float32 else 0
            _lib.call("ava_b200_pca_fit", x.data_ptr(), is_f32, N, D, mean.data_ptr(), cov.data_ptr(),
                      evals.data_ptr(), comps.data_ptr(), ws.data_ptr(), ws_bytes, stream)
            _lib.call("ava_b200_pca_transform", x.data_ptr(), is_f32, N, D, mean.data_ptr(),
                      comps.data_ptr(), K, out.data_ptr(), stream)
        ev = evals.cpu().numpy()
        self.mean_ = mean.cpu().numpy()
        self.covariance_ = cov.cpu().numpy()
        self.components_ = comps.cpu().numpy()[:K].copy()
        self.explained_variance_ = ev[:K].copy()
        total = ev.sum()
        self.explained_variance_ratio_ = ev[:K] / total if total > 0 else np.zeros(K)
        self.n_samples_ = N
        return out

    def fit_transform(self, x):
        """[N,D] -> numpy float64 [N,K] (what ``PCA.fit_transform`` returns)."""
        return self.fit_transform_device(x).cpu().numpy()


class DataContainer():
    """Link directories containing different data sources for easy access: the subset of
    ``ava.data.data_container.DataContainer`` on the VAE hot path (see the module docstring).

    Same constructor arguments and attributes as the reference (data_container.py:232-248).
    Additions (defaults keep the reference's behaviour): ``latent_batch_size`` (the reference
    goes through ``get_syllable_data_loaders``' default of 64) and ``latent_eval`` -- the
    reference never calls ``eval()`` before ``get_latent``, so BatchNorm runs on
    per-batch statistics; ``latent_eval=True`` uses the trained running statistics instead,
    which makes the latent means independent of the batch size.  ``rank`` / ``world_size``
    (one process per GPU, ``torch.distributed`` initialised): every rank computes and writes the
    latent means of its own contiguous block of syllable files (no collective on the data path,
    one barrier before the files are read back); every rank returns the full arrays, rank 0
    writes the PCA projection."""

    def __init__(self, audio_dirs=None, segment_dirs=None, spec_dirs=None,
                 feature_dirs=None, projection_dirs=None, plots_dir='',
                 model_filename=None, template_dir=None, verbose=True,
                 latent_batch_size=64, latent_eval=False, rank=0, world_size=1):
        self.audio_dirs = audio_dirs
        self.segment_dirs = segment_dirs
        self.spec_dirs = spec_dirs
        self.feature_dirs = feature_dirs
        self.projection_dirs = projection_dirs
        self.plots_dir = plots_dir
        self.model_filename = model_filename
        self.template_dir = template_dir
        self.verbose = verbose
        self.latent_batch_size = latent_batch_size
        self.latent_eval = latent_eval
        self.rank, self.world_size = int(rank), int(world_size)
        assert 0 <= self.rank < self.world_size
        self.sylls_per_file = None  # syllables in each file in spec_dirs
        self.fields = self._check_for_fields()
        if self.plots_dir not in [None, ''] and not os.path.exists(self.plots_dir):
            os.makedirs(self.plots_dir)

    def request(self, field):
        """Request some type of data (data_container.py:251-285)."""
        if field not in ALL_FIELDS:
            print(str(field) + " is not a valid field!")
            raise NotImplementedError
        if field not in self.fields:
            if self.verbose:
                print("Making field:", field)
            data = self._make_field(field)
        else:
            if self.verbose:
                print("Reading field:", field)
            data = self._read_field(field)
        if self.verbose:
            print("\tDone with:", field)
        return data

    def clear_projections(self):
        """Remove all projections: deletes the syllable-projection files in
        ``self.projection_dirs`` (data_container.py:288-301)."""
        for proj_dir in self.projection_dirs:
            if not os.path.exists(proj_dir):
                continue
            for fn in get_hdf5s_from_dir(proj_dir):
                os.remove(fn)
        self.fields = self._check_for_fields()

    def _make_field(self, field):
        """Make a field (data_container.py:304-328)."""
        if field == 'latent_means':
            data = self._make_latent_means()
        elif field == 'latent_mean_pca':
            data = self._make_latent_mean_pca_projection()
        else:
            raise NotImplementedError(
                "field %r is outside the VAE hot path this package implements "
                "(available: %s)" % (field, ", ".join(HOT_PATH_FIELDS)))
        self.fields[field] = 1
        if self.verbose:
            print("Making field:", field)
        return data

    def _read_field(self, field):
        """Read a field from the files (data_container.py:331-373)."""
        if field in PROJECTION_FIELDS:
            load_dirs = self.projection_dirs
        elif field in SPEC_FIELDS:
            load_dirs = self.spec_dirs
        else:
            raise NotImplementedError(
                "field %r is outside the VAE hot path this package implements" % (field,))
        to_return = []
        for i in range(len(self.spec_dirs)):
            spec_dir, load_dir = self.spec_dirs[i], load_dirs[i]
            for spec_fn in get_hdf5s_from_dir(spec_dir):
                filename = os.path.join(load_dir, os.path.split(spec_fn)[-1])
                data = read_field(filename, field)
                if field == 'audio_filenames':
                    data = np.array([k.decode('UTF-8') if isinstance(k, bytes) else str(k)
                                     for k in data])
                to_return.append(np.array(data))
        return np.concatenate(to_return)

    def _load_model(self):
        """VAE with the checkpoint's z_dim, as data_container.py:455-459."""
        z_dim = torch.load(self.model_filename, map_location='cpu', weights_only=False)['z_dim']
        model = VAE(z_dim=z_dim)
        model.load_state(self.model_filename)
        if self.latent_eval:
            model.eval()
        return model

    def _make_latent_means(self):
        """Write latent means for the syllables in self.spec_dirs
        (data_container.py:435-487).  Returns the [n_syllables, z_dim] float64 array."""
        self._check_for_dirs(['projection_dirs', 'spec_dirs', 'model_filename'], 'latent_means')
        temp = get_hdf5s_from_dir(self.spec_dirs[0])
        assert len(temp) > 0, "Found no specs in" + self.spec_dirs[0]
        self.sylls_per_file = len(read_field(temp[0], 'specs'))
        spf = self.sylls_per_file
        model = self._load_model()
        all_latent = []
        for i in range(len(self.spec_dirs)):
            spec_dir, proj_dir = self.spec_dirs[i], self.projection_dirs[i]
            if proj_dir != '' and not os.path.exists(proj_dir):
                os.makedirs(proj_dir, exist_ok=True)
            if self.world_size == 1:
                partition = get_syllable_partition([spec_dir], 1, shuffle=False)
            else:
                # this rank's contiguous block of the directory's files; batch-aligned when
                # BatchNorm runs on per-batch statistics (the reference's train-mode quirk)
                every = get_hdf5s_from_dir(spec_dir)
                first, last = shard_files(len(every), self.rank, self.world_size, spf,
                                          1 if self.latent_eval else self.latent_batch_size)
                partition = {'train': every[first:last], 'test': []}
            try:
                assert len(partition['train']) > 0 or self.world_size > 1
                if len(partition['train']) > 0:
                    # streaming: one file resident at a time, same batches as a resident loader
                    loader = get_syllable_data_loaders(partition, batch_size=self.latent_batch_size,
                                                       shuffle=(False, False), streaming=True)['train']
                    latent_means = model.get_latent(loader)
                    spec_fns = partition['train'] if self.world_size > 1 else get_hdf5s_from_dir(spec_dir)
                    assert len(latent_means) // len(spec_fns) == spf
                    for j in range(len(spec_fns)):
                        filename = os.path.join(proj_dir, os.path.split(spec_fns[j])[-1])
                        append_field(filename, 'latent_means', latent_means[j * spf:(j + 1) * spf])
                    if self.world_size == 1:
                        all_latent.append(latent_means)
            except AssertionError:  # No specs in this directory
                pass
            if self.world_size > 1:
                self._barrier()
                parts = [read_field(os.path.join(proj_dir, os.path.split(fn)[-1]), 'latent_means')
                         for fn in get_hdf5s_from_dir(spec_dir)]
                if parts:
                    all_latent.append(np.concatenate(parts))
        return np.concatenate(all_latent)

    def _barrier(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("DataContainer(world_size > 1) needs torch.distributed to be initialised")
        dist.barrier()

    def _make_latent_mean_pca_projection(self):
        """Project latent means to two dimensions with PCA (data_container.py:538-551)."""
        latent_means = self.request('latent_means')
        transform = LatentPCA(n_components=2)
        if self.verbose:
            print("Running PCA...")
        embedding = transform.fit_transform(latent_means)
        self.pca_ = transform
        if self.verbose:
            print("\tDone.")
        if self.rank == 0:
            self._write_projection("latent_mean_pca", embedding)
        if self.world_size > 1:
            self._barrier()
        return embedding

    def _write_projection(self, key, data):
        """Write the given projection to self.projection_dirs (data_container.py:652-665)."""
        sylls_per_file = self.sylls_per_file
        k = 0
        for i in range(len(self.projection_dirs)):
            spec_dir, proj_dir = self.spec_dirs[i], self.projection_dirs[i]
            spec_fns = get_hdf5s_from_dir(spec_dir)
            for j in range(len(spec_fns)):
                filename = os.path.join(proj_dir, os.path.split(spec_fns[j])[-1])
                append_field(filename, key, data[k:k + sylls_per_file])
                k += sylls_per_file

    def _check_for_fields(self):
        """Check to see which fields are saved (data_container.py:668-695)."""
        fields = {}
        if self.spec_dirs is not None:
            for field in SPEC_FIELDS:
                fields[field] = 1
        if self.audio_dirs is not None:
            fields['audio'] = 1
        if self.segment_dirs is not None:
            fields['segments'] = 1
            fields['segment_audio'] = 1
        if self.projection_dirs is not None:
            if os.path.exists(self.projection_dirs[0]):
                fns = get_hdf5s_from_dir(self.projection_dirs[0])
                if len(fns) > 0 and os.path.exists(fns[0]):
                    for key, n in stored_fields(fns[0]).items():
                        if key in ALL_FIELDS:
                            fields[key] = 1
                            self.sylls_per_file = n
        return fields

    def _check_for_dirs(self, dir_names, field):
        """Check that the given directories exist (data_container.py:698-716)."""
        for dir_name in dir_names:
            if dir_name not in ('audio_dirs', 'segment_dirs', 'spec_dirs', 'feature_dirs',
                                'projection_dirs', 'model_filename'):
                raise NotImplementedError
            temp = getattr(self, dir_name)
            assert temp is not None, dir_name + " must be specified before " + \
                field + " is made!"


if __name__ == '__main__':
    pass
