"""
Methods for feeding syllable data to the VAE: drop-in for ``ava.models.vae_dataset``.

``get_syllable_partition`` and ``SyllableDataset`` keep the reference behaviour
(ava/models/vae_dataset.py:21-145): seed-42 file shuffle, ``i // sylls_per_file`` /
``i % sylls_per_file`` indexing, float32 items.  ``get_syllable_data_loaders`` returns,
instead of fork-based torch DataLoaders (4 worker processes re-opening an HDF5 file per
item), ``DeviceSyllableLoader`` objects: the whole split is read once, kept resident in
HBM as fp32, and batches are gathered on the device, so the train step never waits on the
host.  They iterate like a DataLoader (``for batch in loader``), have ``.dataset`` and
``len()``, and yield ``[b,128,128]`` fp32 CUDA tensors (ragged last batch kept,
``drop_last=False`` as in the reference).
"""
import numpy as np
import torch
from torch.utils.data import Dataset

from .utils import _get_sylls_per_file, get_hdf5s_from_dir, numpy_to_tensor, read_specs

EPSILON = 1e-9


def get_syllable_partition(dirs, split, shuffle=True, max_num_files=None):
    """Partition the filenames into a random test/train split
    (ava/models/vae_dataset.py:21-59)."""
    assert(split > 0.0 and split <= 1.0)
    filenames = []
    for dir in dirs:
        filenames += get_hdf5s_from_dir(dir)
    # Reproducibly shuffle.
    filenames = sorted(filenames)
    if shuffle:
        np.random.seed(42)
        np.random.shuffle(filenames)
        np.random.seed(None)
    if max_num_files is not None:
        filenames = filenames[:max_num_files]
    index = int(round(split * len(filenames)))
    return {'train': filenames[:index], 'test': filenames[index:]}


def shard_files(n_files, rank, world_size, sylls_per_file=1, batch_size=1):
    """Contiguous block ``[lo, hi)`` of the (sorted) syllable files for `rank` (SURVEY 8(e):
    ``get_latent`` shards by contiguous index ranges with no collective).  Block boundaries are
    moved to the nearest file boundary that is also a batch boundary
    (``lo * sylls_per_file % batch_size == 0``) when there is one, so that every batch holds the
    same syllables as in the unsharded run -- which matters because the reference leaves
    BatchNorm in train mode during ``get_latent``."""
    assert 0 <= rank < world_size
    q = batch_size // np.gcd(batch_size, sylls_per_file)      # files per batch-aligned period

    def boundary(r):
        b = (n_files * r) // world_size
        if q > 1 and 0 < r < world_size:
            lo_al, hi_al = (b // q) * q, -(-b // q) * q
            cands = [c for c in (lo_al, hi_al) if 0 <= c <= n_files]
            if cands:
                b = min(cands, key=lambda c: (abs(c - b), c))
        return int(b)
    bounds = [boundary(r) for r in range(world_size + 1)]
    bounds[0], bounds[-1] = 0, n_files
    for r in range(1, world_size + 1):                          # keep the blocks ordered
        bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds[rank], bounds[rank + 1]


class SyllableDataset(Dataset):
    """torch.utils.data.Dataset for animal vocalization syllables
    (ava/models/vae_dataset.py:98-145)."""

    def __init__(self, filenames, sylls_per_file, transform=None):
        self.filenames = filenames
        self.sylls_per_file = sylls_per_file
        self.transform = transform

    def __len__(self):
        return len(self.filenames) * self.sylls_per_file

    def __getitem__(self, index):
        result = []
        single_index = False
        try:
            iterator = iter(index)  # noqa: F841
        except TypeError:
            index = [index]
            single_index = True
        for i in index:
            load_filename = self.filenames[i // self.sylls_per_file]
            file_index = i % self.sylls_per_file
            spec = np.asarray(read_specs(load_filename)[file_index])
            if self.transform:
                spec = self.transform(spec)
            result.append(spec)
        if single_index:
            return result[0]
        return result


class DeviceSyllableLoader:
    """All syllables of a split resident on the device; batches are device-side gathers.

    Iteration order: a fresh ``torch.randperm`` per epoch when ``shuffle`` (the reference's
    DataLoader(shuffle=True) also draws a torch permutation), else dataset order.

    Data parallel (``world_size`` > 1): every rank walks the SAME global batches and takes its
    slice of each, so the union over the ranks is exactly the single-process batch.  The
    permutation is therefore drawn from a generator seeded with (``seed``, epoch number) --
    identical on every rank without communication -- instead of each process's global RNG; and
    every rank takes the same number of steps: a rank whose slice of a small ragged last batch
    is empty still yields a ``[0,128,128]`` batch (``VAE.train_step`` then contributes zero
    gradients to the all-reduce), so the collectives of the ranks always match."""

    def __init__(self, dataset, batch_size=64, shuffle=False, device=None, rank=0, world_size=1,
                 streaming=False, seed=0):
        """`streaming` (addition; needs shuffle=False, one rank): do not keep the split
        resident -- read one file at a time and yield the same batches (a batch may straddle
        files), for corpora larger than HBM (`DataContainer` latent means)."""
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.rank, self.world_size = rank, world_size
        self.seed, self._epoch = int(seed), 0
        self.streaming = bool(streaming)
        if self.streaming:
            assert not shuffle and world_size == 1, "streaming loaders walk the files in order"
            self.data = None
            return
        chunks = [np.asarray(read_specs(fn), dtype=np.float32) for fn in dataset.filenames]
        for c in chunks:
            assert len(c) == dataset.sylls_per_file, "files must hold sylls_per_file syllables"
        data = np.concatenate(chunks) if chunks else np.zeros((0, 128, 128), np.float32)
        self.data = torch.from_numpy(data).to(self.device)

    def __len__(self):
        n = len(self.dataset)
        return (n + self.batch_size - 1) // self.batch_size

    def _iter_streaming(self):
        carry = None
        for fn in self.dataset.filenames:
            chunk = torch.from_numpy(np.asarray(read_specs(fn), dtype=np.float32)).to(self.device)
            assert len(chunk) == self.dataset.sylls_per_file, "files must hold sylls_per_file syllables"
            if carry is not None:
                chunk = torch.cat([carry, chunk])
                carry = None
            full = (len(chunk) // self.batch_size) * self.batch_size
            for start in range(0, full, self.batch_size):
                yield chunk[start:start + self.batch_size]
            if full < len(chunk):
                carry = chunk[full:]
        if carry is not None:
            yield carry

    def __iter__(self):
        if self.streaming:
            yield from self._iter_streaming()
            return
        n = len(self.dataset)
        if self.shuffle and self.world_size > 1:
            gen = torch.Generator()
            gen.manual_seed(1000003 * self.seed + self._epoch)
            order = torch.randperm(n, generator=gen).to(self.device)
        elif self.shuffle:
            order = torch.randperm(n).to(self.device)
        else:
            order = None
        self._epoch += 1
        # data-parallel: every rank walks the same global batches and takes its slice
        for start in range(0, n, self.batch_size):
            stop = min(n, start + self.batch_size)
            if self.world_size > 1:
                per = (stop - start + self.world_size - 1) // self.world_size
                lo = min(stop, start + self.rank * per)
                hi = min(stop, lo + per)
            else:
                lo, hi = start, stop
            if hi <= lo:
                if self.world_size == 1:
                    continue
                lo = hi = min(lo, stop)      # empty shard of a small last batch: still a step
            if order is None:
                yield self.data[lo:hi]
            else:
                yield self.data.index_select(0, order[lo:hi])


def get_syllable_data_loaders(partition, batch_size=64, shuffle=(True, False), num_workers=4,
                              device=None, rank=0, world_size=1, streaming=False, seed=0):
    """Return a pair of loaders given a test/train split
    (ava/models/vae_dataset.py:62-94).  `num_workers` is accepted for compatibility and
    ignored (no worker processes are needed)."""
    sylls_per_file = _get_sylls_per_file(partition)
    train_dataset = SyllableDataset(filenames=partition['train'], transform=numpy_to_tensor,
                                    sylls_per_file=sylls_per_file)
    train_dataloader = DeviceSyllableLoader(train_dataset, batch_size=batch_size, shuffle=shuffle[0],
                                            device=device, rank=rank, world_size=world_size,
                                            streaming=streaming, seed=seed)
    if not partition['test']:
        return {'train': train_dataloader, 'test': None}
    test_dataset = SyllableDataset(filenames=partition['test'], transform=numpy_to_tensor,
                                   sylls_per_file=sylls_per_file)
    test_dataloader = DeviceSyllableLoader(test_dataset, batch_size=batch_size, shuffle=shuffle[1],
                                           device=device, rank=rank, world_size=world_size,
                                           streaming=streaming, seed=seed)
    return {'train': train_dataloader, 'test': test_dataloader}


if __name__ == '__main__':
    pass
