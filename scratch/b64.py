import sys, importlib, time
sys.path.insert(0, '/root/repo')
import torch
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
lib = importlib.import_module("autoencoded-vocal-analysis_b200._lib")
for B in (64, 256):
    model = vae_mod.VAE(device_name='cuda'); model.train()
    x = torch.rand(B, 128, 128, device="cuda")
    for _ in range(5): model.train_step(x)
    torch.cuda.synchronize()
    n0 = lib.launch_count(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): model.train_step(x)
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("B=%d eager: %.1f us/step (gpu events)  %.1f us/step (host wall)  launches/step %d  -> %.0f samples/s" % (B, 1e3*e0.elapsed_time(e1)/50, 1e6*(t1-t0)/50, (lib.launch_count()-n0)//50, B*50/(t1-t0)))
