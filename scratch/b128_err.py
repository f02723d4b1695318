"""Whole-model gradient error at a batch the tcgen05 GEMM path tiles (B % 128 == 0), per
precision mode, against the float64 oracle evaluated on the host."""
import importlib, sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vae_oracle
from tests.helpers import rel_err, l2_err
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
seed = int(os.environ.get("SEED", "21"))
P = vae_oracle.make_params(seed)
x = vae_oracle.make_input(seed, B)
ew, ed = vae_oracle.make_noise(seed, B)
t0 = time.time()
P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}
out64, g64, _ = vae_oracle.loss_and_grads(P64, x.double(), ew.double(), ed.double(), 10.0, True)
out32, g32, _ = vae_oracle.loss_and_grads(P, x, ew, ed, 10.0, True)
print("oracle fp64+fp32 on the host: %.1f s" % (time.time() - t0))
ref32 = {k: rel_err(g32[k].numpy(), g64[k].numpy()) for k in g64}
for precision in sys.argv[2:] or ("fp32", "tf32x3"):
    model = vae_mod.VAE(save_dir='', model_precision=10.0, device_name='cuda', precision=precision)
    model.load_flat_state(P)
    model.train(True)
    bufs = model._forward_native(x.cuda(), (ew.cuda(), ed.cuda()), True, want_grad_seed=True)
    model._backward_native(bufs)
    torch.cuda.synchronize()
    print(precision, "seed", seed, "B", B, "loss rel", abs(float(bufs.loss.item()) - float(out64["loss"])) / abs(float(out64["loss"])))
    rows = []
    for k, v in model.grad_dict().items():
        ref = g64[k].numpy()
        v = v.cpu().numpy().reshape(ref.shape)
        e, e2 = rel_err(v, ref), l2_err(v, ref)
        tol = max(1e-4, 3 * ref32[k])
        rows.append((e / tol, k, e, e2, ref32[k]))
    rows.sort(reverse=True)
    for r in rows[:4]:
        print("   %-16s max %.3e l2 %.3e | torch-fp32-cpu %.3e | ratio to tol %.2f" % (r[1], r[2], r[3], r[4], r[0]))
