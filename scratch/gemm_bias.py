"""Error statistics (max / rms / mean = coherent bias, relative to max|ref|) of the dense-layer
forward Y = X W^T at the fc1 shape per inner-product arithmetic, against float64."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, K = 1024, 8192
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(M, K, generator=g) * 0.5)
w = torch.randn(N, K, generator=g) / K ** 0.5
ref = (x.double() @ w.double().T).numpy()
xd, wd = x.cuda(), w.cuda()
y = torch.empty(M, N, device="cuda")
ws_bytes = L.lib().ava_b200_linear_ws_bytes(M, N, K)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
for prec in (0, 1, 2):
    L.call("ava_b200_linear_fwd", xd.data_ptr(), K, wd.data_ptr(), None, y.data_ptr(), N, M, N, K, 0, 1,
           0, 0, 0, 0, prec, ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e = y.cpu().numpy().astype(np.float64) - ref
    s = np.abs(ref).max()
    print("M=%d precision %d (MMA3=%s): max %.2e rms %.2e mean %.2e | mean/rms %.2f" % (
        M, prec, os.environ.get("AVA_B200_GEMM_MMA3", "0"), np.abs(e).max() / s,
        np.sqrt((e ** 2).mean()) / s, e.mean() / s, e.mean() / np.sqrt((e ** 2).mean())))
