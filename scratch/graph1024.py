import sys, importlib, time
sys.path.insert(0, '/root/repo')
import torch
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
xs = [torch.rand(1024, 128, 128, device="cuda") for _ in range(2)]
for graphs in (False, True):
    model = vae_mod.VAE(device_name='cuda', cuda_graphs=graphs); model.train()
    for i in range(6): model.train_step(xs[i & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): model.train_step(xs[i & 1])
    e1.record(); torch.cuda.synchronize()
    print("graphs", graphs, "ms/step %.3f" % (e0.elapsed_time(e1) / 20))
    del model
