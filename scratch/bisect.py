import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import vae_oracle as vo
from tests.helpers import load_golden, rel_err
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
g = load_golden("vae_train_b7"); seed, B = 0, 7
x = vo.make_input(seed, B).cuda()
noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())

def run(model):
    rec = {}
    bufs = model._forward_native(x, noise, True, True)
    rec["fwd.stats"] = bufs.stats.clone()
    for l in range(14): rec["act%d" % l] = bufs.act[l].clone()
    g_cur, g_nxt = bufs.g[0], bufs.g[1]
    rec["seed"] = g_cur[:B*16384].clone()
    for l in range(13, 6, -1):
        xin = bufs.act[l-1] if l > 7 else bufs.t8
        model._conv_bwd(l, bufs, g_cur, bufs.act[l], xin, g_nxt, has_next_bn=(l < 13))
        g_cur, g_nxt = g_nxt, g_cur
        rec["gin%d" % l] = g_cur[:xin.numel()].clone()
        rec["dstats%d" % l] = bufs.dstats[64*l:64*l+64].clone()
        name = vae_mod._LAYERS[l][0]
        rec["dw%d" % l] = model._views_g[name + ".weight"].clone()
        rec["db%d" % l] = model._views_g[name + ".bias"].clone()
    torch.cuda.synchronize()
    return rec

recs = []
for rep in range(4):
    model = vae_mod.VAE(device_name='cuda'); model.load_flat_state(vo.make_params(seed)); model.train()
    recs.append(run(model))
for rep in range(1, 4):
    print("--- rep", rep, "vs rep 0")
    for k in recs[0]:
        a, b = recs[0][k].double().cpu().numpy(), recs[rep][k].double().cpu().numpy()
        e = rel_err(a, b)
        if e > 1e-6: print("   %-10s differs %.2e" % (k, e))
