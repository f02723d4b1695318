"""Experiment (round 2): can BatchNorm-backward's two per-channel reductions (sum g, sum g*(x-mean))
of layer l be obtained WITHOUT materialising g, from the weight gradient of the same layer?

    sum_q g[ci,q]          = sum_{co,k} W[co,ci,k] * T_k[co]        T_k = border-trimmed sums of dz
    sum_q g[ci,q]*xn[ci,q] = sum_{co,k} W[co,ci,k] * dW[co,ci,k]    (xn = bn(x), zero padded)

If so, the backward-data kernel can apply the BatchNorm backward + ReLU mask in its epilogue and
write the next dz directly (no bn_relu_bwd_apply pass, no g round trip through HBM).  This script
measures, on the real kernels, how far the algebraic values are from the epilogue-accumulated
ones and from the float64 oracle, and what substituting them does to the final gradients.

Run on the GPU box:  python profiles/probes/exp_algebraic_dstats.py [B] [precision]
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vae_oracle  # noqa: E402
from tests.helpers import gpu_relu_masks, masked_oracle_grads, rel_err  # noqa: E402

PKG = "autoencoded-vocal-analysis_b200"
vae_mod = importlib.import_module(PKG + ".models.vae")
LAYERS = vae_mod._LAYERS


def trimmed_sums(l, dz, h_in):
    """T[co,3,3] = weight gradient of layer l w.r.t. an all-ones single input channel."""
    name, ci, co, s, h = LAYERS[l]
    B = dz.shape[0]
    ones = torch.ones(B, 1, h_in, h_in, dtype=torch.float64, device=dz.device)
    if l < 7:
        w0 = torch.zeros(co, 1, 3, 3, dtype=torch.float64, device=dz.device, requires_grad=True)
        y = F.conv2d(ones, w0, stride=s, padding=1)
    else:
        w0 = torch.zeros(1, co, 3, 3, dtype=torch.float64, device=dz.device, requires_grad=True)
        y = F.conv_transpose2d(ones, w0, stride=s, padding=1, output_padding=s - 1)
    (y * dz.double()).sum().backward()
    return w0.grad.reshape(co, 3, 3)


def run(B, precision, substitute, seed=3):
    P = vae_oracle.make_params(seed)
    model = vae_mod.VAE(save_dir='', device_name='cuda', precision=precision, cuda_graphs=False)
    model.load_flat_state(P)
    model.train()
    x = vae_oracle.make_input(seed, B).cuda()
    ew, ed = vae_oracle.make_noise(seed, B)
    bufs = model._forward_native(x, (ew.cuda(), ed.cuda()), True, want_grad_seed=True)
    rows = []
    orig = model._conv_bwd

    def wrapped(l, bufs, g_out, y, xin, g_in, has_next_bn):
        orig(l, bufs, g_out, y, xin, g_in, has_next_bn)
        name, ci, co, s, h = LAYERS[l]
        ho = vae_mod._out_hw(l)
        dz = g_out.reshape(-1)[:B * co * ho * ho].view(B, co, ho, ho)
        W = dict(model.named_parameters())[name + ".weight"].detach().double()
        dW = model._views_g[name + ".weight"].double()
        gamma = dict(model.named_parameters())["bn%d.weight" % (l + 1)].detach().double()
        beta = dict(model.named_parameters())["bn%d.bias" % (l + 1)].detach().double()
        n = float(B * h * h)
        st = bufs.stats[64 * l:64 * l + 64]
        mean = st[:ci] / n
        var = st[32:32 + ci] / n - mean * mean
        invstd = torch.rsqrt(var + 1e-5)
        T = trimmed_sums(l, dz, h)                      # [co,3,3]
        if l < 7:   # W [co,ci,3,3]
            sum_g = torch.einsum("oikl,okl->i", W, T)
            wdw = (W * dW).sum(dim=(0, 2, 3))
        else:       # W [ci,co,3,3]
            sum_g = torch.einsum("iokl,okl->i", W, T)
            wdw = (W * dW).sum(dim=(1, 2, 3))
        S = (wdw - beta * sum_g) / (gamma * invstd)
        ds = bufs.dstats[64 * l:64 * l + 64]
        k_sum_g, k_S = ds[:ci].clone(), ds[32:32 + ci].clone()
        if g_in is not None:
            g = g_in.reshape(-1)[:B * ci * h * h].view(B, ci, h, h).double()
            sd = g.std(dim=(0, 2, 3))
            mg = g.mean(dim=(0, 2, 3)).abs()
        else:
            sd = torch.full((ci,), float("nan"), device=dz.device, dtype=torch.float64)
            mg = sd
        rows.append((l, name, (sum_g - k_sum_g).abs() / n / sd, (S - k_S).abs() * invstd / n / sd, mg / sd,
                     k_S.abs() * invstd / n / sd))
        if substitute:
            ds[:ci] = sum_g
            ds[32:32 + ci] = S
    model._conv_bwd = wrapped
    model._backward_native(bufs)
    torch.cuda.synchronize()
    masks = gpu_relu_masks(bufs)
    _, g64, _, _, err32 = masked_oracle_grads(P, x.cpu(), ew, ed, 10.0, masks)
    err32.pop("__flips32__", None)
    errs = {k: rel_err(v.cpu().numpy(), g64[k].numpy()) for k, v in model.grad_dict().items()}
    return rows, errs, err32


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    precision = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
    rows, e0, err32 = run(B, precision, False)
    print("B=%d precision=%s" % (B, precision))
    print("layer: |d sum_g|/(N std g)   |d S| invstd/(N std g)   |mean g|/std g   |S| invstd/(N std g)   (max over channels)")
    for l, name, a, b, c, d in rows:
        print("%2d %-7s %.2e  %.2e  %.2e  %.2e" % (l, name, float(a.max()), float(b.max()), float(c.max()), float(d.max())))
    _, e1, _ = run(B, precision, True)
    worst0 = sorted(e0.items(), key=lambda kv: -kv[1])[:8]
    print("gradient error vs masked float64 oracle, kernel dstats:     max %.2e  " % max(e0.values()),
          ", ".join("%s %.1e" % kv for kv in worst0))
    worst1 = sorted(e1.items(), key=lambda kv: -kv[1])[:8]
    print("gradient error vs masked float64 oracle, algebraic dstats:  max %.2e  " % max(e1.values()),
          ", ".join("%s %.1e" % kv for kv in worst1))
    print("oracle's own fp32 error on the same pattern: max %.2e" % max(err32.values()))
    bn = [k for k in e0 if k.startswith("bn")]
    print("bn grads: kernel max %.2e, algebraic max %.2e" % (max(e0[k] for k in bn), max(e1[k] for k in bn)))


if __name__ == "__main__":
    main()
