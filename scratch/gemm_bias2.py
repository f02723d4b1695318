"""Error statistics of linear fwd / bwd_data / bwd_weight at the fc1 and fc8 shapes (batch M),
per inner-product arithmetic, against float64."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
st = lambda: torch.cuda.current_stream().cuda_stream
def stats(name, got, ref):
    e = got.astype(np.float64) - ref
    s = np.abs(ref).max()
    rms = np.sqrt((e ** 2).mean())
    print("   %-10s max %.2e rms %.2e mean %.2e mean/rms %+.3f" % (name, np.abs(e).max() / s, rms / s, e.mean() / s, e.mean() / max(rms, 1e-300)))
for (N, K) in ((1024, 8192), (8192, 1024), (256, 1024)):
    g = torch.Generator().manual_seed(N)
    x = torch.relu(torch.randn(M, K, generator=g) * 0.5)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    dy = torch.randn(M, N, generator=g)
    y64 = torch.relu(x.double() @ w.double().T + b.double())
    mask = (y64 > 0).double()
    dym = dy.double() * mask
    gx64 = (dym @ w.double()).numpy()
    gw64 = (dym.T @ x.double()).numpy()
    xd, wd, bd, dyd, ymask = x.cuda(), w.cuda(), b.cuda(), dy.cuda(), y64.float().cuda()
    ws_bytes = L.lib().ava_b200_linear_ws_bytes(M, N, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    for prec in (0, 2):
        y = torch.empty(M, N, device="cuda"); gx = torch.empty(M, K, device="cuda")
        gw = torch.empty(N, K, device="cuda"); gb = torch.empty(N, device="cuda")
        L.call("ava_b200_linear_fwd", xd.data_ptr(), K, wd.data_ptr(), bd.data_ptr(), y.data_ptr(), N, M, N, K, 1, 1,
               0, 0, 0, 0, prec, ws.data_ptr(), ws_bytes, st())
        L.call("ava_b200_linear_bwd_weight", dyd.data_ptr(), N, ymask.data_ptr(), xd.data_ptr(), K, gw.data_ptr(),
               gb.data_ptr(), M, N, K, 1, 0, 0, 0, 0, prec, ws.data_ptr(), ws_bytes, st())
        L.call("ava_b200_linear_bwd_data", dyd.data_ptr(), N, ymask.data_ptr(), wd.data_ptr(), gx.data_ptr(), K,
               M, N, K, 1, 0, 0, 0, 0, prec, ws.data_ptr(), ws_bytes, st())
        torch.cuda.synchronize()
        print("M=%d N=%d K=%d precision %d MMA3=%s" % (M, N, K, prec, os.environ.get("AVA_B200_GEMM_MMA3", "0")))
        stats("fwd", y.cpu().numpy(), y64.numpy())
        stats("bwd_data", gx.cpu().numpy(), gx64)
        stats("bwd_weight", gw.cpu().numpy(), gw64)
