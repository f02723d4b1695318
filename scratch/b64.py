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
# graph vs eager equivalence + timing
from oracle import vae_oracle as vo
for B in (64,):
    ma = vae_mod.VAE(device_name='cuda', cuda_graphs=True); mb = vae_mod.VAE(device_name='cuda', cuda_graphs=False)
    P = vo.make_params(1); ma.load_flat_state(P); mb.load_flat_state(P); ma.train(); mb.train()
    for s_ in range(6):
        x = vo.make_input(s_, B).cuda(); nz = tuple(t.cuda() for t in vo.make_noise(s_, B))
        la = ma.train_step(x, noise=nz).item(); lb = mb.train_step(x, noise=nz).item()
        print("step", s_, la, lb, "graph" if ma._graphs[B]["graph"] is not None else "eager")
    same = all(torch.equal(a, b) for a, b in zip(ma.state_dict().values(), mb.state_dict().values()))
    print("params identical after 6 steps:", same, ma._step_host, mb._step_host)
    x = torch.rand(B,128,128, device="cuda")
    for _ in range(5): ma.train_step(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): ma.train_step(x)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("B=%d graph: %.1f us/step -> %.0f samples/s" % (B, 1e6*(t1-t0)/200, B*200/(t1-t0)))
