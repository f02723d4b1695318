import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import vae_oracle as vo
from tests.helpers import load_golden, rel_err, l2_err
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
for name in ("vae_train_b7", "vae_train_b64"):
    g = load_golden(name); seed, B = int(g["seed"]), int(g["batch"])
    model = vae_mod.VAE(device_name='cuda'); model.load_flat_state(vo.make_params(seed)); model.train()
    x = vo.make_input(seed, B).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    ref = None
    for rep in range(6):
        P0 = {k: v.clone() for k, v in model.state_dict().items()}
        bufs = model._forward_native(x, noise, True, True)
        if rep % 2 == 1:   # poison every scratch buffer the backward writes
            for t in bufs.g + [bufs.dt8, bufs.dt7, bufs.dt6, bufs.dt5, bufs.gz, bufs.gheads, bufs.dh3, bufs.dh2, bufs.dh1, bufs.da6]:
                if t is bufs.g[0]: continue   # holds dL/dx_rec
                t.fill_(float('nan'))
            model._flat_g.fill_(float('nan'))
            if model._scratch is not None: model._scratch.fill_(255)
        model._backward_native(bufs); torch.cuda.synchronize()
        grads = {k: v.clone() for k, v in model.grad_dict().items()}
        model.load_state_dict(P0)   # undo running-stat updates
        if ref is None: ref = grads
        worst = max((rel_err(grads[k].cpu().numpy(), ref[k].cpu().numpy()), k) for k in grads)
        gw = rel_err(grads["conv1.weight"].cpu().numpy(), g["grad:conv1.weight"])
        nan = [k for k in grads if not torch.isfinite(grads[k]).all()]
        print(name, "rep", rep, "worst dev from rep0: %.2e (%s)" % worst, " conv1.weight vs golden %.2e" % gw, "nan:", nan[:3])
