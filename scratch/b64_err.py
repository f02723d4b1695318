"""Whole-model gradient error at B=64 per precision mode against the golden (float64 reference)."""
import importlib, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vae_oracle
from tests.helpers import load_golden, rel_err, l2_err, digest
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
g = load_golden("vae_train_b64")
seed, batch = int(g["seed"]), int(g["batch"])
for precision in ("fp32", "tf32x3"):
    model = vae_mod.VAE(save_dir='', model_precision=float(g["model_precision"]), device_name='cuda', precision=precision)
    model.load_flat_state(vae_oracle.make_params(seed))
    model.train(True)
    x = vae_oracle.make_input(seed, batch).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    bufs = model._forward_native(x, noise, True, want_grad_seed=True)
    model._backward_native(bufs)
    torch.cuda.synchronize()
    print(precision, "loss rel", abs(float(bufs.loss.item()) - float(g["loss"])) / abs(float(g["loss"])))
    rows = []
    for k, v in model.grad_dict().items():
        key = "grad:" + k
        v = v.cpu().numpy()
        if key in g.files:
            ref = g[key]; e, e2 = rel_err(v.reshape(ref.shape), ref), l2_err(v.reshape(ref.shape), ref)
        else:
            ref = g[key + "__digest"]; got = digest(v)
            nrm = abs(got[1] - ref[1]) / max(ref[1], 1e-30)
            e = max(nrm, rel_err(got[2:], ref[2:])); e2 = max(nrm, l2_err(got[2:], ref[2:]))
        tol = max(1e-4, 3 * float(g["err32:" + key]))
        rows.append((e / tol, k, e, e2, tol))
    rows.sort(reverse=True)
    for r in rows[:8]:
        print("   %-16s max %.3e l2 %.3e tol %.3e ratio %.2f" % (r[1], r[2], r[3], r[4], r[0]))
