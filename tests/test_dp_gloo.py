"""CPU, world_size 2 over gloo: the host-side data-parallel logic (gradient SUM buckets,
per-batch loss-constant correction, BN-buffer averaging, loader sharding, rank-0-only
checkpoint).  The kernels themselves need a GPU and are covered by the -m gpu tests."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "autoencoded-vocal-analysis_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    vae = importlib.import_module(PKG + ".models.vae")
    ds_mod = importlib.import_module(PKG + ".models.vae_dataset")
    torch.manual_seed(100 + rank)                      # different init per rank ...
    model = vae.VAE(save_dir=tmp, device_name='cpu')
    model.enable_data_parallel()                       # ... rank 0's parameters win
    ref = [torch.zeros_like(model._flat_p) for _ in range(world)]
    dist.all_gather(ref, model._flat_p)
    assert torch.equal(ref[0], ref[1])
    # gradient buckets: SUM over ranks, every element reduced exactly once
    early, mid, late = model._grad_buckets()
    covered = torch.zeros(model._n_flat, dtype=torch.int32)
    for lo, hi in early + mid + late:
        covered[lo:hi] += 1
    # each bucket holds exactly the parameters its point of the backward pass has completed
    off = model._off
    assert early == [(off["fc5.weight"], off["bn1.weight"])] and mid == [(off["fc1.weight"], off["fc5.weight"])]
    assert off["fc43.bias"] < off["fc5.weight"] and off["conv7.bias"] < off["fc1.weight"]
    assert bool((covered == 1).all())
    model._flat_g.copy_(torch.arange(model._n_flat, dtype=torch.float32) % 97 + rank)
    works = model._allreduce(early, async_op=True) + model._allreduce(mid, async_op=True) + \
        model._allreduce(late, async_op=True)
    for w in works:
        w.wait()
    want = 2 * (torch.arange(model._n_flat, dtype=torch.float32) % 97) + 1
    assert torch.equal(model._flat_g, want)
    # loss bookkeeping: per-batch constants counted once per GLOBAL batch
    c = model.loss_constant()
    model._loss_sum.fill_(10.0 * (rank + 1) + 3 * c)   # 3 local steps
    model._flat_run.fill_(float(rank))
    total = model._epoch_loss(3)
    assert abs(total - (30.0 + 3 * c)) < 1e-6 * abs(c)
    assert torch.allclose(model._flat_run, torch.full_like(model._flat_run, 0.5))
    # loader sharding: the two ranks together see every global batch exactly once, in order
    files = []
    for i in range(2):
        fn = os.path.join(tmp, "s%d.npy" % i)
        if rank == 0:
            np.save(fn, np.arange(5 * 128 * 128, dtype=np.float64).reshape(5, 128, 128) + 1e6 * i)
        files.append(fn)
    dist.barrier()
    loaders = ds_mod.get_syllable_data_loaders({'train': files, 'test': []}, batch_size=4,
                                               shuffle=(False, False), device='cpu', rank=rank,
                                               world_size=world)
    assert loaders['test'] is None
    mine = [b[:, 0, 0].clone() for b in loaders['train']]
    sizes = [len(b) for b in mine]
    assert sizes == ([2, 2, 1] if rank == 0 else [2, 2, 1])
    gathered = [None, None]
    dist.all_gather_object(gathered, [b.tolist() for b in mine])
    firsts = []
    for step in range(3):
        firsts += gathered[0][step] + gathered[1][step]
    want_first = [float(np.float32(r * 128 * 128 + 1e6 * f)) for f in range(2) for r in range(5)]
    assert firsts == want_first
    # shuffle=True with a 1-sample tail (round-1 advisor finding): both ranks walk the SAME
    # permutation, take the same number of steps (the rank with an empty shard still yields a
    # [0,128,128] batch), and their shards tile each global batch exactly
    sh = ds_mod.get_syllable_data_loaders({'train': files, 'test': []}, batch_size=3,
                                          shuffle=(True, False), device='cpu', rank=rank,
                                          world_size=world, seed=7)['train']
    for epoch in range(2):
        mine = [b[:, 0, 0].tolist() for b in sh]
        assert len(mine) == 4 and [len(b) for b in mine] == ([2, 2, 2, 1] if rank == 0 else [1, 1, 1, 0])
        both = [None, None]
        dist.all_gather_object(both, mine)
        seen = [v for step in range(4) for r in range(2) for v in both[r][step]]
        assert sorted(seen) == sorted(want_first)                 # every syllable exactly once
        if epoch == 0:
            first_epoch = seen
        else:
            assert seen != first_epoch                            # a fresh permutation per epoch
        gen = torch.Generator()
        gen.manual_seed(1000003 * 7 + epoch)
        perm = torch.randperm(10, generator=gen).tolist()
        assert seen == [want_first[i] for i in perm]              # == the single-process batches
    # only rank 0 writes checkpoints
    model.save_state("dp.tar")
    dist.barrier()
    assert os.path.exists(os.path.join(tmp, "dp.tar"))
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


class _FakeModel:
    """Stands in for the VAE in the DataContainer sharding test: 'latent means' that are a
    deterministic function of each spectrogram AND of the batch it arrived in (like BatchNorm
    on per-batch statistics), so misplaced rows or changed batch boundaries are both caught."""

    def get_latent(self, loader):
        out = []
        for batch in loader:
            b = batch.double()
            out.append(torch.stack([b[:, 0, 0], b.mean(dim=(1, 2)), b[:, 5, 7] - b[:, 0, 0].mean()], 1))
        return torch.cat(out).numpy()


def _container_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import functools
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dcm = importlib.import_module(PKG + ".data.data_container")
    mu = importlib.import_module(PKG + ".models.utils")
    dcm.get_syllable_data_loaders = functools.partial(dcm.get_syllable_data_loaders, device='cpu')
    ext = ".npz" if not mu._have_h5py() else ".hdf5"
    spf, batch = 4, 8
    sdirs = [os.path.join(tmp, "s0"), os.path.join(tmp, "s1")]
    pdirs = [os.path.join(tmp, "p0_w%d" % world), os.path.join(tmp, "p1_w%d" % world)]
    if rank == 0 and not os.path.exists(sdirs[0]):
        rng = np.random.default_rng(0)
        for sd, nf in zip(sdirs, (5, 1)):
            os.makedirs(sd)
            for j in range(nf):
                mu.append_field(os.path.join(sd, "syllables_%04d%s" % (j, ext)), 'specs',
                                rng.random((spf, 128, 128)))
    dist.barrier()
    dc = dcm.DataContainer(spec_dirs=sdirs, projection_dirs=pdirs, model_filename="unused.tar",
                           verbose=False, latent_batch_size=batch, rank=rank, world_size=world)
    dc._load_model = lambda: _FakeModel()
    latent = dc.request('latent_means')
    # unsharded answer, directory by directory, batches of `batch` across file boundaries
    want = []
    for sd in sdirs:
        specs = np.concatenate([mu.read_field(fn, 'specs') for fn in mu.get_hdf5s_from_dir(sd)])
        chunks = [torch.from_numpy(specs[i:i + batch].astype(np.float32)) for i in range(0, len(specs), batch)]
        want.append(_FakeModel().get_latent(chunks))
    want = np.concatenate(want)
    np.testing.assert_array_equal(latent, want)            # same rows, same batch boundaries
    # every projection file written exactly once, by its owner
    for sd, pd in zip(sdirs, pdirs):
        assert sorted(os.listdir(pd)) == sorted(os.path.basename(f) for f in mu.get_hdf5s_from_dir(sd))
    assert dc.sylls_per_file == spf and 'latent_means' in dc.fields
    np.testing.assert_array_equal(dc.request('latent_means'), want)    # now read from the files
    dist.barrier()
    dist.destroy_process_group()


def test_data_container_sharded_latent_means_world2(tmp_path):
    """get_latent shards by contiguous blocks of syllable files with no data-path
    collective; each rank writes its own projection files; all ranks return the full array."""
    port = _free_port()
    mp.spawn(_container_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # and a world of one gives the same files through the unsharded code path
    port = _free_port()
    mp.spawn(_container_worker, args=(1, port, str(tmp_path)), nprocs=1, join=True)
