"""Data parallelism on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_dp.py -m gpu`; skipped on a single-GPU box, where bench.py's `dp_check` -- the same
routine, run by the driver before every multi-GPU timing -- is the evidence)."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
PKG = "autoencoded-vocal-analysis_b200"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bench
    from oracle import vae_oracle
    vae_mod = importlib.import_module(PKG + ".models.vae")
    torch.manual_seed(50 + rank)
    model = vae_mod.VAE(save_dir='', device_name='cuda')
    model.enable_data_parallel()
    model.train()
    # (1) summed rank gradients == the gradients of the shards pushed through one model
    res = bench.dp_check(model, vae_mod, dist, world, rank)
    assert res["ok"], res
    # (2) the captured-graph step follows the eager step, for the fused NVLink optimizer kernel
    # (csrc/dp.cu; peer loads/stores and, where the box has NVLS, the multimem variant) and for
    # the bucketed NCCL all-reduce path; and all of them end at the same parameters
    def run(graphs, fused, multimem=True):
        os.environ["AVA_B200_DP_MULTIMEM"] = "1" if multimem else "0"
        m = vae_mod.VAE(save_dir='', device_name='cuda', cuda_graphs=graphs)
        m.load_flat_state(vae_oracle.make_params(3))
        m.enable_data_parallel(fused=fused)
        m.train()
        assert (m._dp_fused is not None) == fused
        out = []
        for step in range(6):
            x = vae_oracle.make_input(100 * rank + step, 16).cuda()
            noise = tuple(t.cuda() for t in vae_oracle.make_noise(100 * rank + step, 16))
            out.append(float(m.train_step(x, noise=noise)))
        if graphs:
            assert m._graphs[16]["graph"] is not None, "the data-parallel step was not captured"
        m.dp_check_status()
        torch.cuda.synchronize()
        # replicas stay identical
        mine = m._flat_p.clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(mine, other), "rank %d's parameters differ from rank 0's" % rank
        return out, mine, m
    runs = {}
    for key in (("nccl", False), ("nccl", True), ("fused", False), ("fused", True), ("fused_p2p", True)):
        runs[key] = run(graphs=key[1], fused=key[0] != "nccl", multimem=key[0] != "fused_p2p")
    os.environ.pop("AVA_B200_DP_MULTIMEM", None)
    ref_loss, ref_p, _ = runs[("nccl", False)]
    for key, (loss, p, m) in runs.items():
        for a, b in zip(loss, ref_loss):
            assert abs(a - b) <= 2e-4 * abs(b), (key, loss, ref_loss)
        # Adam's early steps are sign-like: a near-zero gradient summed in another order can flip an
        # lr-sized update; everything else agrees to rounding
        assert (p - ref_p).abs().max().item() <= 6 * 2.5e-3, key
        # NCCL and the peer-load kernel form the 2-rank sum a+b exactly alike; the NVSwitch's adder
        # (multimem.ld_reduce) may round differently in the last bit, and a 1-ulp change of a
        # near-zero gradient flips more of those lr-sized updates: its bar is wider (still 100x
        # below the per-step parameter change a wrong gradient would cause); the bit-exactness of
        # the multimem data path itself is pinned below with integer-valued gradients
        bar = 3e-5 if (key[0] == "fused" and m._dp_fused["multimem"]) else 2e-6
        mean = (p - ref_p).abs().mean().item()
        if mean > bar and rank == 0:       # say where before failing
            for name, off in sorted(m._off.items(), key=lambda kv: kv[1]):
                n = dict(m.named_parameters())[name].numel()
                d = (p[off:off + n] - ref_p[off:off + n]).abs()
                if d.mean().item() > bar:
                    print("  %s %-14s mean %.3e max %.3e frac>1e-4 %.4f" % (key, name, d.mean().item(), d.max().item(),
                                                                         (d > 1e-4).float().mean().item()))
        assert mean <= bar, (key, mean)
    # the sharded Adam moments gather into a complete optimizer state (what save_state writes)
    _, _, mf = runs[("fused", False)]
    _, _, mn = runs[("nccl", False)]
    mf._gather_moment_shards()
    assert (mf._flat_m - mn._flat_m).abs().max().item() <= 1e-3 * mn._flat_m.abs().max().item()
    assert (mf._flat_v - mn._flat_v).abs().max().item() <= 1e-3 * mn._flat_v.abs().max().item()
    # ONE fused optimizer step against its definition, isolated from training dynamics: local
    # gradients from a real backward pass; expected = Adam (the single-GPU kernel) on the NCCL sum
    # of the ranks' gradients, from the same parameters / moments / step count
    lib = importlib.import_module(PKG + "._lib")
    for variant in ("peers_p2p", "peers_mc"):
        if mf._dp_fused[variant] is None:
            continue
        mf._dp_fused["peers"] = mf._dp_fused[variant]
        x = vae_oracle.make_input(900 + rank, 16).cuda()
        noise = tuple(t.cuda() for t in vae_oracle.make_noise(900 + rank, 16))
        mf._ensure_optimizer_state()
        mf._sync_hyper()
        bufs = mf._forward_native(x, noise, True, want_grad_seed=True)
        mf._backward_native(bufs)
        torch.cuda.synchronize()
        g_sum = mf._flat_g.clone()
        dist.all_reduce(g_sum)
        p0, m0, v0, s0 = mf._flat_p.clone(), mf._flat_m.clone(), mf._flat_v.clone(), mf._step_dev.clone()
        lib.call("ava_b200_adam_step_dev", p0.data_ptr(), g_sum.data_ptr(), m0.data_ptr(), v0.data_ptr(),
                 mf._n_flat, s0.data_ptr(), mf._hyper_dev.data_ptr(), 1.0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
        mf._adam_dp_native()
        torch.cuda.synchronize()
        dist.barrier()
        d = (mf._flat_p - p0).abs()
        n4, W = mf._n_flat // 4, world
        lo, hi = 4 * (n4 * rank // W), 4 * (n4 * (rank + 1) // W)
        dm = ((mf._flat_m[lo:hi] - m0[lo:hi]).abs() / (m0[lo:hi].abs() + 1e-30)).max().item()
        dv = ((mf._flat_v[lo:hi] - v0[lo:hi]).abs() / (v0[lo:hi].abs() + 1e-30)).max().item()
        print("rank %d %s: one-step |p - expected| max %.3e mean %.3e; own-slice moments rel. max %.1e / %.1e"
              % (rank, variant, d.max().item(), d.mean().item(), dm, dv))
        # p moves by <= lr = 1e-3 per step: agreement to a few ulps of the UPDATE; the 2-rank sum is
        # a + b on every path, so peer loads must reproduce the single-GPU kernel bit for bit
        assert d.max().item() <= 2e-6, (variant, d.max().item())
        assert dm <= 1e-5 and dv <= 1e-5, (variant, dm, dv)
        if variant == "peers_p2p" and world == 2:
            assert torch.equal(mf._flat_p, p0), d.max().item()
        mf._gather_moment_shards()
    mf.dp_check_status()
    # the two variants of the fused kernel move the same data: with integer-valued gradients every
    # sum is exact whatever the adder's rounding, so peer loads and multimem must agree bit for bit
    f = mf._dp_fused
    if f["peers_mc"] is not None:
        ends = []
        gen = torch.Generator(device="cuda").manual_seed(11 + rank)
        gint = torch.randint(-64, 65, (mf._n_flat,), device="cuda", generator=gen).float()
        keep_p, keep_m, keep_v = mf._flat_p.clone(), mf._flat_m.clone(), mf._flat_v.clone()
        keep_step = mf._step_dev.clone()
        for peers in (f["peers_p2p"], f["peers_mc"]):
            mf._flat_p.copy_(keep_p)
            mf._flat_m.copy_(keep_m)
            mf._flat_v.copy_(keep_v)
            mf._step_dev.copy_(keep_step)        # (same bias correction for both variants)
            mf._flat_g.copy_(gint)
            f["peers"] = peers
            torch.cuda.synchronize()
            dist.barrier()
            mf._adam_dp_native()
            torch.cuda.synchronize()
            dist.barrier()
            ends.append(mf._flat_p.clone())
        mf.dp_check_status()
        assert torch.equal(ends[0], ends[1]), (ends[0] - ends[1]).abs().max().item()
        assert not torch.equal(ends[0], keep_p)
    print("rank %d: fused data-parallel step verified (multimem available: %s)" % (rank, mf._dp_fused["multimem_available"]))
    # (3) a rank with an empty shard still takes the step (zero gradients into the all-reduce)
    m = vae_mod.VAE(save_dir='', device_name='cuda')
    m.load_flat_state(vae_oracle.make_params(3))
    m.enable_data_parallel()
    m.train()
    x = vae_oracle.make_input(7, 4).cuda()
    noise = tuple(t.cuda() for t in vae_oracle.make_noise(7, 4))
    if rank == 0:
        m.train_step(x, noise=noise)
    else:
        m.train_step(x[:0])
    ref = vae_mod.VAE(save_dir='', device_name='cuda')       # not data parallel: the whole batch on one GPU
    ref.load_flat_state(vae_oracle.make_params(3))
    ref.train()
    ref.train_step(x, noise=noise)
    torch.cuda.synchronize()
    d = (m._flat_p - ref._flat_p).abs().max().item()
    assert d <= 2.5e-3, d        # Adam's first step is sign-like: at most lr-sized flips of ~0 gradients
    assert (m._flat_p - ref._flat_p).abs().mean().item() <= 1e-6
    dist.destroy_process_group()


def test_data_parallel_equivalence_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
