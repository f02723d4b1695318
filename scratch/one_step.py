import sys, importlib
sys.path.insert(0, '/root/repo')
import torch
from oracle import vae_oracle as vo
vae_mod = importlib.import_module("autoencoded-vocal-analysis_b200.models.vae")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = vae_mod.VAE(device_name='cuda'); model.load_flat_state(vo.make_params(0)); model.train()
x = vo.make_input(0, B).cuda(); ew, ed = vo.make_noise(0, B)
loss = model.train_step(x, noise=(ew.cuda(), ed.cuda()))
torch.cuda.synchronize()
print("loss", float(loss))
