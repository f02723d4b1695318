import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import vae_oracle as vo
from tests.helpers import load_golden, rel_err
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
g = load_golden("vae_train_b7")
seed, batch = 0, 7
P = vo.make_params(seed)
model = vae_mod.VAE(device_name='cuda'); model.load_flat_state(P); model.train()
x = vo.make_input(seed, batch)
ew, ed = torch.from_numpy(g["eps_w"]), torch.from_numpy(g["eps_d"])
bufs = model._forward_native(x.cuda(), (ew.cuda(), ed.cuda()), True, True)
torch.cuda.synchronize()
P64 = {k:(v.double() if v.is_floating_point() else v) for k,v in P.items()}
a64, a32 = {}, {}
vo.forward(P64, x.double(), ew.double(), ed.double(), 10.0, True, None, a64)
vo.forward(P, x, ew, ed, 10.0, True, None, a32)
names = [n for n,_,_,_ in vo.ENC_CONVS] + [n for n,_,_,_ in vo.DEC_CONVTS]
for l, n in enumerate(names):
    ref = a64[n].detach().numpy()
    print("%-7s ours %.2e   torch-fp32 %.2e" % (n, rel_err(bufs.act[l].cpu().numpy(), ref), rel_err(a32[n].detach().numpy(), ref)))
print("ReLU sign flips vs fp64 (ours / torch-fp32):")
for l, n in enumerate(names[:-1]):
    ref = a64[n].detach().numpy() > 0
    print("%-7s %d / %d of %d" % (n, ((bufs.act[l].cpu().numpy() > 0) != ref).sum(), ((a32[n].detach().numpy() > 0) != ref).sum(), ref.size))
import torch.nn.functional as F
