import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import vae_oracle
from tests.helpers import load_golden, rel_err, l2_err, digest
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
name = sys.argv[1] if len(sys.argv) > 1 else "vae_train_b64"
g = load_golden(name)
seed, batch = int(g["seed"]), int(g["batch"])
res = {}
for prec in sys.argv[2:] or ["fp32", "auto", "tf32"]:
    model = vae_mod.VAE(save_dir='', model_precision=float(g["model_precision"]), device_name='cuda', precision=prec)
    model.load_flat_state(vae_oracle.make_params(seed)); model.train(True)
    x = vae_oracle.make_input(seed, batch).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    bufs = model._forward_native(x, noise, True, want_grad_seed=True)
    model._backward_native(bufs); torch.cuda.synchronize()
    print(prec, "loss relerr %.2e" % (abs(float(bufs.loss.item()) - float(g["loss"])) / abs(float(g["loss"]))))
    for k, v in model.grad_dict().items():
        key = "grad:" + k
        if key in g.files:
            ref = g[key]; got = v.cpu().numpy().reshape(ref.shape)
            e = rel_err(got, ref)
        else:
            ref = g[key + "__digest"]; got = digest(v.cpu().numpy())
            e = max(abs(got[1] - ref[1]) / max(ref[1], 1e-30), rel_err(got[2:], ref[2:]))
        res.setdefault(k, {})[prec] = e
        res[k]["ref32"] = float(g["err32:" + key])
for k, d in res.items():
    if k.endswith("weight") and ("conv" in k or "fc1" in k or "fc8" in k):
        print("%-16s" % k, "  ".join("%s %.2e" % (p, e) for p, e in d.items()))
