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
g_cur, g_nxt = bufs.g[0], bufs.g[1]
ours_h = {}
ours_h[14] = g_cur[:batch*16384].clone()
for l in range(13, 6, -1):
    xin = bufs.act[l-1] if l > 7 else bufs.t8
    model._conv_bwd(l, bufs, g_cur, bufs.act[l], xin, g_nxt, has_next_bn=(l < 13))
    g_cur, g_nxt = g_nxt, g_cur
    ours_h[l] = g_cur[:xin.numel()].clone()
torch.cuda.synchronize()
# fp64 oracle with retained BN outputs
P64 = {k:(v.double() if v.is_floating_point() else v) for k,v in P.items()}
Q = {k:(v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v) for k,v in P64.items()}
xd = x.cpu().double()
mu,u,d = vo.encode(Q, xd, True)
z = vo.rsample(mu,u,d,noise[0].cpu().double(), noise[1].cpu().double())
h = F.relu(F.linear(z, Q["fc5.weight"], Q["fc5.bias"])); h = F.relu(F.linear(h, Q["fc6.weight"], Q["fc6.bias"])); h = F.relu(F.linear(h, Q["fc7.weight"], Q["fc7.bias"]))
t8 = F.relu(F.linear(h, Q["fc8.weight"], Q["fc8.bias"])); t8.retain_grad()
hh = t8.reshape(-1,32,16,16)
bnout = {}
for i,(name,_,_,s) in enumerate(vo.DEC_CONVTS):
    n = vo.batch_norm(hh, Q, i+8, True); n.retain_grad(); bnout[i+7] = n
    hh = F.conv_transpose2d(n, Q[name+".weight"], Q[name+".bias"], stride=s, padding=1, output_padding=s-1)
    if name != "convt7": hh = F.relu(hh)
xr = hh.reshape(-1,16384)
loss = 0.5*10.0*((xd.reshape(batch,-1)-xr)**2).sum() + 0.5*(z*z).sum() - vo.entropy(u,d).sum()
loss.backward()
for l in range(13,6,-1):
    ref = bnout[l].grad.numpy().ravel()
    print("layer", l, "g_in rel err %.2e" % rel_err(ours_h[l].cpu().numpy(), ref), " max|g| %.3e" % np.abs(ref).max())
# dt8
st = bufs.stats.cpu().numpy().reshape(14,64); ds = bufs.dstats.cpu().numpy().reshape(14,64)
gbn = bnout[7].grad
t8d = t8.detach().reshape(-1,32,16,16)
mean = t8d.mean(dim=(0,2,3))
print("dstats[7] dbeta rel err %.2e" % rel_err(ds[7,:32], gbn.sum(dim=(0,2,3)).numpy()))
S_ref = (gbn*(t8d-mean.view(1,-1,1,1))).sum(dim=(0,2,3)).numpy()
print("dstats[7] S rel err %.2e" % rel_err(ds[7,32:], S_ref), "S", S_ref[:4], "sum|terms|", (gbn*(t8d-mean.view(1,-1,1,1))).abs().sum(dim=(0,2,3)).numpy()[:4])
print("t8.grad max", t8.grad.abs().max().item(), " bn8out.grad max", gbn.abs().max().item())
# ours dt8 with fp64 math from OUR h and stats
import ctypes
from importlib import import_module
L = import_module("autoencoded-vocal-analysis_b200._lib")
dt8 = torch.empty(batch, 8192, device="cuda")
L.call("ava_b200_bn_relu_bwd_apply", ours_h[7].data_ptr(), bufs.t8.data_ptr(), model._p("bn8.weight"), bufs.stats.data_ptr()+8*64*7, bufs.dstats.data_ptr()+8*64*7, batch, 32, 256, dt8.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
ref_dt8 = (t8.grad * (t8.detach()>0)).numpy()
print("dt8 rel err %.2e" % rel_err(dt8.cpu().numpy(), ref_dt8), "max|dt8| %.3e"%np.abs(ref_dt8).max())
# same with exact (fp64) h and stats on CPU to see sensitivity: perturb h by fp32 rounding only
h32 = gbn.float().double()
c1 = h32.mean(dim=(0,2,3)); S = (h32*(t8d-mean.view(1,-1,1,1))).sum(dim=(0,2,3))
var = ((t8d-mean.view(1,-1,1,1))**2).mean(dim=(0,2,3)); invstd = 1/torch.sqrt(var+1e-5); gam = P64["bn8.weight"]
N = batch*256
dx = (gam*invstd).view(1,-1,1,1)*(h32 - c1.view(1,-1,1,1)) - (gam*invstd**3*S/N).view(1,-1,1,1)*(t8d-mean.view(1,-1,1,1))
dx = (dx*(t8d>0)).reshape(batch,-1).numpy()
print("dt8 from fp32-rounded exact h: rel err %.2e" % rel_err(dx, ref_dt8))
print("db8 from that: rel err %.2e" % rel_err(dx.sum(0), ref_dt8.sum(0)), " ours db8 rel err %.2e" % rel_err(dt8.cpu().numpy().astype(np.float64).sum(0), ref_dt8.sum(0)))
