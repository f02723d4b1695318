"""Run selected fused-layer kernels once at a given batch (for ncu)."""
import sys, importlib
sys.path.insert(0, '/root/repo')
import torch
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib"); L.lib()
LAYERS = [(1, 8, 1, 128, 0), (8, 8, 2, 128, 0), (8, 16, 1, 64, 0), (16, 16, 2, 64, 0),
          (16, 24, 1, 32, 0), (24, 24, 2, 32, 0), (24, 32, 1, 16, 0), (32, 24, 1, 16, 1),
          (24, 24, 2, 16, 1), (24, 16, 1, 32, 1), (16, 16, 2, 32, 1), (16, 8, 1, 64, 1),
          (8, 8, 2, 64, 1), (8, 1, 1, 128, 1)]
B = int(sys.argv[1]); jobs = sys.argv[2:]   # e.g. fwd:13 bwdd:13 bwdw:12
s = torch.cuda.current_stream().cuda_stream
for job in jobs:
    if job == "fc1":
        M, N, K = B, 1024, 8192
        x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.01; bb = torch.zeros(N, device="cuda")
        y = torch.empty(M, N, device="cuda")
        wsb = L.lib().ava_b200_linear_ws_bytes(M, N, K); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        for rep in range(2):
            L.call("ava_b200_linear_fwd", x.data_ptr(), K, w.data_ptr(), bb.data_ptr(), y.data_ptr(), N, M, N, K, 1, 1, 0, 0, 0, 0, 2, ws.data_ptr(), wsb, s)
        torch.cuda.synchronize(); continue
    kind, l = job.split(":"); l = int(l)
    ci, co, st, h, tr = LAYERS[l]
    ho = h if st == 1 else (h*2 if tr else h//2)
    x = torch.rand(B, ci, h, h, device="cuda"); y = torch.rand(B, co, ho, ho, device="cuda") - 0.3
    g = torch.randn(B, co, ho, ho, device="cuda"); gin = torch.empty(B, ci, h, h, device="cuda")
    w = torch.randn(ci*co*9, device="cuda")*0.1; b = torch.zeros(co, device="cuda")
    gam = torch.ones(32, device="cuda"); bet = torch.zeros(32, device="cuda")
    stats = torch.zeros(4*64, dtype=torch.float64, device="cuda")
    L.call("ava_b200_channel_stats", x.data_ptr(), B, ci, h*h, stats.data_ptr(), s)
    L.call("ava_b200_channel_stats", y.data_ptr(), B, co, ho*ho, stats.data_ptr()+8*64, s)
    ws = torch.empty(L.lib().ava_b200_bnconv_bwd_weight_ws(l, B), dtype=torch.uint8, device="cuda")
    dw = torch.empty(ci*co*9, device="cuda"); db = torch.empty(32, device="cuda")
    for rep in range(2):
        if kind == "ew":
            L.call("ava_b200_bn_relu_bwd_apply", g.data_ptr(), y.data_ptr(), gam.data_ptr(), stats.data_ptr()+8*64, stats.data_ptr()+8*192, B, co, ho*ho, 1, g.data_ptr(), s)
        elif kind == "fwd":
            L.call("ava_b200_bnconv_fwd", l, B, x.data_ptr(), y.data_ptr(), w.data_ptr(), b.data_ptr(), gam.data_ptr(), bet.data_ptr(),
                   stats.data_ptr(), gam.data_ptr(), gam.data_ptr(), 1, stats.data_ptr()+8*128, s)
        elif kind == "bwdd":
            L.call("ava_b200_bnconv_bwd_data", l, B, g.data_ptr(), w.data_ptr(), x.data_ptr(), stats.data_ptr(), gin.data_ptr(), stats.data_ptr()+8*128, s)
        else:
            L.call("ava_b200_bnconv_bwd_weight", l, B, g.data_ptr(), x.data_ptr(), gam.data_ptr(), bet.data_ptr(), stats.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), s)
    torch.cuda.synchronize()
print("done")
