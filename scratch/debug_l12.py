import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import torch.nn.functional as F
from oracle import vae_oracle as vo
from tests.helpers import load_golden, rel_err
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
g = load_golden("vae_train_b7")
seed, batch = 0, 7
P = vo.make_params(seed)
model = vae_mod.VAE(device_name='cuda'); model.load_flat_state(P); model.train()
x = vo.make_input(seed, batch).cuda()
noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
bufs = model._forward_native(x, noise, True, True)
bufs.alloc_backward(32)
g0, g1 = bufs.g[0], bufs.g[1]
seed_g = g0[:batch*16384].clone()
model._conv_bwd(13, bufs, g0, bufs.act[13], bufs.act[12], g1, has_next_bn=False)
h13 = g1[:bufs.act[12].numel()].clone().view_as(bufs.act[12])     # grad wrt bn14 output
model._conv_bwd(12, bufs, g1, bufs.act[12], bufs.act[11], g0, has_next_bn=True)
h12 = g0[:bufs.act[11].numel()].clone().view_as(bufs.act[11])
torch.cuda.synchronize()
# fp64 evaluation of layer 12's backward-data from OUR inputs
y = bufs.act[12].cpu().double(); h = h13.cpu().double()
N = y.shape[0]*y.shape[2]*y.shape[3]
mu = y.mean(dim=(0,2,3), keepdim=True); var = ((y-mu)**2).mean(dim=(0,2,3), keepdim=True); invstd = 1/torch.sqrt(var+1e-5)
gam = P["bn14.weight"].double().view(1,-1,1,1)
c1 = h.mean(dim=(0,2,3), keepdim=True); S = (h*(y-mu)).sum(dim=(0,2,3), keepdim=True)
dz = (gam*invstd)*(h-c1) - gam*invstd**3*S/N*(y-mu)
dz = dz*(y>0)
W = P["convt6.weight"].double()    # [ci=8, co=8,3,3]
gin_ref = F.conv2d(dz, W, None, stride=2, padding=1)   # dgrad of convT s2 = conv s2 with same weight layout (out=ci)
print("layer12 g_in vs fp64-on-our-inputs: rel err %.2e" % rel_err(h12.cpu().numpy(), gin_ref.numpy()))
# compare accumulated stats with fp64 ones from our tensors
st = bufs.stats.cpu().numpy().reshape(14,64); ds = bufs.dstats.cpu().numpy().reshape(14,64)
print("stats[13] sum rel err %.2e  sumsq rel err %.2e" % (rel_err(st[13,:8], y.sum(dim=(0,2,3)).numpy()), rel_err(st[13,32:40], (y*y).sum(dim=(0,2,3)).numpy())))
print("dstats[13] dbeta rel err %.2e  S rel err %.2e" % (rel_err(ds[13,:8], h.sum(dim=(0,2,3)).numpy()), rel_err(ds[13,32:40], S.flatten().numpy())))
print("S", S.flatten().numpy(), "\nours", ds[13,32:40], "\nsum|terms|", (h*(y-mu)).abs().sum(dim=(0,2,3)).numpy())
print("dbeta", h.sum(dim=(0,2,3)).numpy(), "\nours ", ds[13,:8], "\nsum|h|", h.abs().sum(dim=(0,2,3)).numpy())
print("mean y", mu.flatten().numpy(), "ours", st[13,:8]/N)
