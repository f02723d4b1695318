"""Time every fused conv layer call (fwd / bwd-data / bwd-weight) at batch B for each conv precision mode."""
import sys, importlib, json
sys.path.insert(0, '/root/repo')
import torch
L = importlib.import_module("autoencoded-vocal-analysis_b200._lib"); L.lib()
LAYERS = [(1, 8, 1, 128, 0), (8, 8, 2, 128, 0), (8, 16, 1, 64, 0), (16, 16, 2, 64, 0),
          (16, 24, 1, 32, 0), (24, 24, 2, 32, 0), (24, 32, 1, 16, 0), (32, 24, 1, 16, 1),
          (24, 24, 2, 16, 1), (24, 16, 1, 32, 1), (16, 16, 2, 32, 1), (16, 8, 1, 64, 1),
          (8, 8, 2, 64, 1), (8, 1, 1, 128, 1)]
B = int(sys.argv[1]); modes = [int(m) for m in sys.argv[2].split(",")]
kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fwd", "bwdd", "bwdw"]
layers = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else list(range(14))
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for l in layers:
    ci, co, st, h, tr = LAYERS[l]
    ho = h if st == 1 else (h*2 if tr else h//2)
    x = torch.rand(B, ci, h, h, device="cuda"); y = torch.rand(B, co, ho, ho, device="cuda") - 0.3
    g = torch.randn(B, co, ho, ho, device="cuda"); gin = torch.empty(B, ci, h, h, device="cuda")
    w = torch.randn(ci*co*9, device="cuda")*0.1; b = torch.zeros(co, device="cuda")
    gam = torch.ones(32, device="cuda"); bet = torch.zeros(32, device="cuda")
    stats = torch.zeros(4*64, dtype=torch.float64, device="cuda")
    L.call("ava_b200_channel_stats", x.data_ptr(), B, ci, h*h, stats.data_ptr(), s)
    ws = torch.empty(L.lib().ava_b200_bnconv_bwd_weight_ws(l, B), dtype=torch.uint8, device="cuda")
    dw = torch.empty(ci*co*9, device="cuda"); db = torch.empty(32, device="cuda")
    for kind in kinds:
        for mode in modes:
            L.call("ava_b200_set_conv_precision", mode)
            def run():
                if kind == "fwd":
                    L.call("ava_b200_bnconv_fwd", l, B, x.data_ptr(), y.data_ptr(), w.data_ptr(), b.data_ptr(), gam.data_ptr(), bet.data_ptr(),
                           stats.data_ptr(), gam.data_ptr(), gam.data_ptr(), 1, stats.data_ptr()+8*128, s)
                elif kind == "bwdd":
                    L.call("ava_b200_bnconv_bwd_data", l, B, g.data_ptr(), w.data_ptr(), x.data_ptr(), stats.data_ptr(), gin.data_ptr(), stats.data_ptr()+8*128, s)
                else:
                    L.call("ava_b200_bnconv_bwd_weight", l, B, g.data_ptr(), x.data_ptr(), gam.data_ptr(), bet.data_ptr(), stats.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), s)
            run(); run()
            ts = []
            for rep in range(5):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            res["%s[%d] mode%d" % (kind, l, mode)] = ts[2]
    L.call("ava_b200_set_conv_precision", 0)
for kind in kinds:
    for l in layers:
        print("%-5s[%2d]" % (kind, l), "  ".join("m%d %7.1f us" % (m, res["%s[%d] mode%d" % (kind, l, m)]) for m in modes))
tot = {m: sum(v for k, v in res.items() if k.endswith("mode%d" % m)) for m in modes}
print("total", tot)
