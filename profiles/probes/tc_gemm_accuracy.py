"""Probe (round 2): accuracy of the dense-layer kernels at the benchmarked batch (1024) on
post-ReLU-like (non-negative) activations, per precision mode: max-norm relative error vs
float64 and the mean SIGNED relative deviation of the large outputs (a systematic shrink is the
signature of truncating accumulation in the tensor core over a long K chain)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib")
L.lib()
st = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(1024, 1024, 8192), (1024, 8192, 1024), (1024, 256, 1024), (1024, 1024, 256),
                  (128, 1024, 8192)]:
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.relu(torch.randn(M, K, generator=g)).float()
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).float()
    b = (torch.randn(N, generator=g) * 0.1).float()
    ref = (x.double() @ w.double().t() + b.double())
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    ws_bytes = max(int(L.lib().ava_b200_linear_ws_bytes(M, N, K)), 1 << 20)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    for prec in (0, 2, 1):
        y = torch.empty(M, N, device="cuda")
        L.call("ava_b200_linear_fwd", xd.data_ptr(), K, wd.data_ptr(), bd.data_ptr(), y.data_ptr(), N, M, N, K,
               0, 1, 0, 0, 0, 0, prec, ws.data_ptr(), ws_bytes, st)
        torch.cuda.synchronize()
        got = y.double().cpu()
        err = float((got - ref).abs().max() / ref.abs().max())
        big = ref.abs() > 0.2 * ref.abs().max()
        shrink = float(((got[big] - ref[big]) / ref[big]).mean())
        print("fwd M=%d N=%d K=%d precision=%d: max-norm rel err %.2e, mean signed rel dev of large outputs %+.2e"
              % (M, N, K, prec, err, shrink))
