"""Per-layer forward / backward-data error vs float64 at a batch large enough that CTAs loop over several tiles."""
import sys, importlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import torch.nn.functional as F
import tests.test_gpu_kernels as T
from tests.helpers import rel_err
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib"); L.lib()
B = int(sys.argv[1]); layers = [int(v) for v in sys.argv[2].split(",")]
for l in layers:
    name, ci, co, s, h, tr = T.LAYERS[l]
    x, w, b, gamma, beta, rm, rv = T.make_layer_inputs(l, B, 100 + l)
    xn, y_ref = T.layer_ref(l, x, w, b, gamma, beta, True, rm, rv)
    dx, dw, db, dg, dbeta, drm, drv = [T.dev(t) for t in (x, w, b, gamma, beta, rm, rv)]
    ho = y_ref.shape[-1]
    gen = torch.Generator().manual_seed(7)
    R = torch.randn(y_ref.shape, generator=gen, dtype=torch.float64).float().double()
    # reference data gradient wrt bn output
    xn_ = xn.detach().requires_grad_(True)
    yy = F.conv_transpose2d(xn_, w, b, stride=s, padding=1, output_padding=s - 1) if tr else F.conv2d(xn_, w, b, stride=s, padding=1)
    (yy * R).sum().backward()
    out = []
    for mode in (0, 2):
        L.call("ava_b200_set_conv_precision", mode)
        stats = torch.zeros(2 * 64, dtype=torch.float64, device="cuda")
        L.call("ava_b200_channel_stats", dx.data_ptr(), B, ci, h * h, stats.data_ptr(), T.stream())
        y = torch.empty(B, co, ho, ho, device="cuda")
        L.call("ava_b200_bnconv_fwd", l, B, dx.data_ptr(), y.data_ptr(), dw.data_ptr(), db.data_ptr(), dg.data_ptr(), dbeta.data_ptr(),
               stats.data_ptr(), drm.data_ptr(), drv.data_ptr(), 1, stats.data_ptr() + 8 * 64, T.stream())
        gin = torch.empty(B, ci, h, h, device="cuda")
        dst = torch.zeros(64, dtype=torch.float64, device="cuda")
        dR = T.dev(R)
        L.call("ava_b200_bnconv_bwd_data", l, B, dR.data_ptr(), dw.data_ptr(), dx.data_ptr(), stats.data_ptr(), gin.data_ptr(), dst.data_ptr(), T.stream())
        torch.cuda.synchronize()
        st = stats.cpu().numpy()
        e_s1 = rel_err(st[64:64 + co], y_ref.sum(dim=(0, 2, 3)).numpy())
        e_s2 = rel_err(st[96:96 + co], (y_ref * y_ref).sum(dim=(0, 2, 3)).numpy())
        mean_x = x.mean(dim=(0, 2, 3), keepdim=True)
        got = dst.cpu().numpy()
        scale = max(np.abs(xn_.grad.numpy()).sum() / ci, 1e-30)
        e_db = np.abs(got[:ci] - xn_.grad.sum(dim=(0, 2, 3)).numpy()).max() / scale
        e_dg = np.abs(got[32:32 + ci] - (xn_.grad * (x - mean_x)).sum(dim=(0, 2, 3)).numpy()).max() / scale
        e_f = rel_err(y.cpu().numpy(), y_ref.numpy())
        e_b = rel_err(gin.cpu().numpy(), xn_.grad.numpy())
        # worst image index
        d = (y.cpu().double() - y_ref).abs().amax(dim=(1, 2, 3))
        out.append("m%d fwd %.1e bwdd %.1e sum %.1e sumsq %.1e dbeta %.1e dgam %.1e" % (mode, e_f, e_b, e_s1, e_s2, e_db, e_dg))
    L.call("ava_b200_set_conv_precision", 0)
    print(l, name, " | ".join(out))
