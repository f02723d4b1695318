"""GPU parity tests of the whole VAE hot path against the golden vectors generated
from the reference's own code (tests/golden/, see oracle/make_golden.py) and against
the CPU oracle."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle
from tests.helpers import check_against_golden, load_golden, rel_err

pytestmark = pytest.mark.gpu
PKG = "autoencoded-vocal-analysis_b200"

FWD_TOL = 1e-4        # forward quantities, fp32 kernels vs float64 reference (rtol 1e-4)


def grad_tol(g, key, floor=2e-3, k=3.0):
    """Whole-model gradient tolerance.  Per-kernel arithmetic is held to rtol 1e-4 in
    tests/test_gpu_kernels.py (well-conditioned single layers vs float64).  End to end the
    gradient of this network is chaotic at the 1e-3 level in ANY fp32 implementation: every
    BatchNorm backward cancels the dominant gradient component (amplifying rounding noise
    ~1000x) and a forward rounding difference of 1e-7 can flip a ReLU mask (measured: one
    flipped unit of 917,504 in convt6 moves the conv-stack gradients by 1e-3; the
    reference's own fp32 CPU path is up to 2.5e-3 away from its float64 path at B=64, see
    err32:* in the goldens).  So: within 2e-3 of the float64 reference, or within 3x the
    reference's own fp32 error on that tensor, whichever is larger."""
    return max(floor, k * float(g["err32:" + key]))


KINK_TOL = 5e-2       # gradients of a run whose ReLU pattern differs from the float64 reference's
MAX_FLIPS = 16        # units (of ~1.9 M per sample) allowed to sit on the other side of zero


def relu_flips(bufs, seed, batch, eps_w, eps_d, prec, train):
    """Number of ReLU units whose on/off state in the GPU forward differs from the float64
    oracle's on the same inputs.  The network's gradient is discontinuous there: a unit whose
    pre-activation is within rounding of zero (|v| ~ 1e-7) lands on either side depending on
    summation order, and ONE such unit moves the BatchNorm-amplified gradients by ~1e-2
    (measured over seeds 21-24 at batch 64, profiles/r01_b64_seeds.txt: whichever of the fp32
    FMA / 3xTF32 paths has a flip is ~5e-3..1e-2 off, the other ~2e-5..2e-4)."""
    P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in vae_oracle.make_params(seed).items()}
    acts = {}
    x = vae_oracle.make_input(seed, batch).double()
    vae_oracle.forward(P64, x, torch.from_numpy(eps_w).double(), torch.from_numpy(eps_d).double(), prec,
                       train, {}, acts)
    names = [n for n, _, _, _ in vae_oracle.ENC_CONVS] + [n for n, _, _, _ in vae_oracle.DEC_CONVTS]
    pairs = [(acts[n], bufs.act[l]) for l, n in enumerate(names) if n != "convt7"]
    pairs += [(acts["fc1"], bufs.h1), (acts["fc2"], bufs.h2), (acts["fc3"], bufs.h3), (acts["fc5"], bufs.t5),
              (acts["fc6"], bufs.t6), (acts["fc7"], bufs.t7), (acts["fc8"], bufs.t8)]
    flips = 0
    for ref, got in pairs:
        flips += int(((ref > 0) != (got.detach().cpu().reshape(ref.shape) > 0)).sum())
    return flips


@pytest.fixture(scope="module")
def vae_mod():
    return importlib.import_module(PKG + ".models.vae")


def build(vae_mod, seed, prec=10.0, **kw):
    model = vae_mod.VAE(save_dir='', model_precision=prec, device_name='cuda', **kw)
    model.load_flat_state(vae_oracle.make_params(seed))
    return model


@pytest.mark.parametrize("name", ["vae_train_b7", "vae_eval_b7", "vae_train_b1",
                                  "vae_train_b64"])
def test_forward_backward_matches_reference_golden(vae_mod, name):
    g = load_golden(name)
    seed, batch, train = int(g["seed"]), int(g["batch"]), bool(g["train"])
    prec = float(g["model_precision"])
    model = build(vae_mod, seed, prec)
    model.train(train)
    x = vae_oracle.make_input(seed, batch).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    bufs = model._forward_native(x, noise, train, want_grad_seed=True)
    torch.cuda.synchronize()
    Z = 32
    loss = float(bufs.loss.item())
    assert abs(loss - float(g["loss"])) <= FWD_TOL * abs(float(g["loss"]))
    heads = bufs.heads.cpu().numpy()
    assert rel_err(heads[:, :Z], g["mu"]) <= FWD_TOL
    assert rel_err(heads[:, Z:2 * Z], g["u"][:, :, 0]) <= FWD_TOL
    assert rel_err(bufs.d.cpu().numpy(), g["d"]) <= FWD_TOL
    assert rel_err(bufs.z.cpu().numpy(), g["z"]) <= FWD_TOL
    check_against_golden(g, "", "x_rec", bufs.act[13].cpu().numpy().reshape(batch, 128, 128),
                         FWD_TOL)
    if train:
        flips = relu_flips(bufs, seed, batch, g["eps_w"], g["eps_d"], prec, train)
        assert flips <= MAX_FLIPS, "%d ReLU units differ from the float64 forward" % flips
        model._backward_native(bufs)
        torch.cuda.synchronize()
        grads = model.grad_dict()
        for k, v in grads.items():
            # same ReLU pattern as the reference: the fp32 bar; a unit on the other side of the
            # kink: the gradient legitimately differs by the kink's size (see relu_flips)
            tol = grad_tol(g, "grad:" + k) if flips == 0 else max(KINK_TOL, grad_tol(g, "grad:" + k))
            check_against_golden(g, "grad:", k, v.cpu().numpy(), tol)
        print("%s: %d ReLU flips vs float64" % (name, flips))
    sd = model.state_dict()
    for k in g.files:
        if not k.startswith("buf:"):
            continue
        kk = k[4:]
        if kk.endswith("num_batches_tracked"):
            assert int(sd[kk]) == int(g[k]), kk
        else:
            assert rel_err(sd[kk].cpu().numpy(), g[k]) <= FWD_TOL, kk


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
def test_precision_modes(vae_mod, precision):
    """precision='fp32' (FMA only) and 'tf32x3' (error-compensated tensor-core products, the
    'auto' default) both hold the fp32 parity bar; 'tf32' is the opt-in reduced-precision mode
    (what torch/cuDNN do by default for the reference's convs on a GPU) with its own
    stated tolerance: forward 1e-2, gradients within 25% in the L2 sense (BatchNorm backward
    amplifies a 5e-4 operand rounding by ~1000x, see grad_tol)."""
    from tests.helpers import l2_err
    g = load_golden("vae_train_b7")
    seed, batch = int(g["seed"]), int(g["batch"])
    model = build(vae_mod, seed, float(g["model_precision"]), precision=precision)
    model.train(True)
    x = vae_oracle.make_input(seed, batch).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    bufs = model._forward_native(x, noise, True, want_grad_seed=True)
    model._backward_native(bufs)
    torch.cuda.synchronize()
    lib = importlib.import_module(PKG + "._lib").lib()
    assert lib.ava_b200_get_conv_precision() == {"fp32": 0, "tf32x3": 2, "tf32": 1}[precision]
    ftol = 1e-2 if precision == "tf32" else FWD_TOL
    assert abs(float(bufs.loss.item()) - float(g["loss"])) <= ftol * abs(float(g["loss"]))
    assert rel_err(bufs.heads.cpu().numpy()[:, :32], g["mu"]) <= ftol
    check_against_golden(g, "", "x_rec", bufs.act[13].cpu().numpy().reshape(batch, 128, 128), ftol)
    for k, v in model.grad_dict().items():
        if precision == "tf32":
            key = "grad:" + k
            if key in g.files:
                assert l2_err(v.cpu().numpy().reshape(g[key].shape), g[key]) <= 0.25, k
        else:
            check_against_golden(g, "grad:", k, v.cpu().numpy(), grad_tol(g, "grad:" + k))


def test_public_api_encode_decode_forward(vae_mod):
    g = load_golden("vae_eval_b7")
    seed, batch = int(g["seed"]), int(g["batch"])
    model = build(vae_mod, seed)
    model.eval()
    x = vae_oracle.make_input(seed, batch).cuda()
    with torch.no_grad():
        mu, u, d = model.encode(x)
        assert mu.shape == (batch, 32) and u.shape == (batch, 32, 1) and d.shape == (batch, 32)
        assert rel_err(mu.cpu().numpy(), g["mu"]) <= FWD_TOL
        assert rel_err(u.cpu().numpy(), g["u"]) <= FWD_TOL
        assert rel_err(d.cpu().numpy(), g["d"]) <= FWD_TOL
        xr = model.decode(torch.from_numpy(g["z"]).float().cuda())
        assert xr.shape == (batch, 16384)
        check_against_golden(g, "", "x_rec", xr.cpu().numpy().reshape(batch, 128, 128), FWD_TOL)
        noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
        loss, z, rec = model.forward(x, return_latent_rec=True, noise=noise)
        assert loss.dim() == 0 and isinstance(z, np.ndarray) and rec.shape == (batch, 128, 128)
        assert abs(loss.item() - float(g["loss"])) <= FWD_TOL * abs(float(g["loss"]))
        assert rel_err(z, g["z"]) <= FWD_TOL
    # eval mode must not touch the running buffers
    sd = model.state_dict()
    P = vae_oracle.make_params(seed)
    for k in ("bn1.running_mean", "bn9.running_var"):
        assert torch.equal(sd[k].cpu(), P[k])
    assert int(sd["bn3.num_batches_tracked"]) == 3


def test_autograd_backward_and_torch_seeded_noise(vae_mod):
    """loss = model.forward(x); loss.backward() (the reference's train_epoch idiom,
    ava/models/vae.py:350-352) must leave the same gradients as the native path, and
    with no injected noise the draws come from torch.randn in the reference's order."""
    g = load_golden("vae_train_b7")
    seed, batch = int(g["seed"]), int(g["batch"])
    model = build(vae_mod, seed)
    model.train()
    x = vae_oracle.make_input(seed, batch).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    model.optimizer.zero_grad()
    loss = model.forward(x, noise=noise)
    assert loss.requires_grad
    loss.backward()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        check_against_golden(g, "grad:", k, p.grad.cpu().numpy(), grad_tol(g, "grad:" + k))
    # torch-seeded draws: eps_W [B,1] first, then eps_D [B,32]
    torch.manual_seed(123)
    ew = torch.randn(batch, 1, device="cuda")
    ed = torch.randn(batch, 32, device="cuda")
    model2 = build(vae_mod, seed)
    model2.train()
    with torch.no_grad():
        l_inj = model2.forward(x, noise=(ew, ed)).item()
    model3 = build(vae_mod, seed)
    model3.train()
    torch.manual_seed(123)
    with torch.no_grad():
        l_drawn = model3.forward(x).item()
    # (equal up to the summation order of the statistics atomics)
    assert abs(l_inj - l_drawn) <= 1e-6 * abs(l_inj)


def test_train_steps_match_reference_adam_trajectory(vae_mod):
    g = load_golden("adam_b5_s3")
    seed, batch, steps = int(g["seed"]), int(g["batch"]), int(g["steps"])
    model = build(vae_mod, seed)
    model.train()
    for s in range(steps):
        x = vae_oracle.make_input(seed + s, batch).cuda()
        ew, ed = vae_oracle.make_noise(seed + s, batch)
        loss = model.train_step(x, noise=(ew.cuda(), ed.cuda()))
        # Adam's first steps are sign-like (m/sqrt(v) ~ +-1), which amplifies fp32 noise in
        # tiny gradients; the reference's own fp32 run is 2e-4 off the float64 one at step 3.
        tol = max(1e-4, 3 * abs(float(g["err32:losses"])))
        assert abs(loss.item() - g["losses"][s]) <= tol * abs(g["losses"][s]), s
    assert model._step_host == steps
    sd = model.state_dict()
    for k in ("conv1.weight", "fc1.weight", "fc43.bias", "convt7.weight", "bn5.weight"):
        tol = max(1e-4, 3 * float(g["err32:param:" + k]))
        check_against_golden(g, "param:", k, sd[k].cpu().numpy(), tol)
    assert int(sd["bn1.num_batches_tracked"]) == 3 + steps


class _ListLoader:
    """Minimal stand-in for a DataLoader: iterable of CPU batches with a .dataset."""

    def __init__(self, data, batch_size):
        self.dataset = data
        self.batch_size = batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield self.dataset[i:i + self.batch_size]


def test_epoch_loops_get_latent_and_checkpoint(vae_mod, tmp_path):
    seed = 6
    data = vae_oracle.make_input(seed, 20)          # 20 specs, batch 8 -> ragged last batch of 4
    loader = _ListLoader(data, 8)
    model = vae_mod.VAE(save_dir=str(tmp_path), device_name='cuda')
    model.load_flat_state(vae_oracle.make_params(seed))
    torch.manual_seed(0)
    l0 = model.train_epoch(loader)
    assert model.epoch == 1 and np.isfinite(l0)
    lt = model.test_epoch(loader)
    assert np.isfinite(lt)
    model.save_state("checkpoint_001.tar")
    ck = torch.load(os.path.join(str(tmp_path), "checkpoint_001.tar"), map_location="cpu")
    # reference checkpoint layout: 40 layers + 6 entries (ava/models/vae.py:433-446)
    assert len(ck) == 46 and ck['z_dim'] == 32 and ck['epoch'] == 1
    assert set(ck['fc1'].keys()) == {'weight', 'bias'}
    assert set(ck['bn1'].keys()) == {'weight', 'bias', 'running_mean', 'running_var',
                                     'num_batches_tracked'}
    ost = ck['optimizer_state']
    assert len(ost['state']) == 80 and len(ost['param_groups'][0]['params']) == 80
    assert float(ost['state'][0]['step']) == 3.0
    assert ost['state'][14]['exp_avg'].shape == (1,)          # bn1.weight is parameter #14
    # resume == uninterrupted
    model2 = vae_mod.VAE(save_dir=str(tmp_path), device_name='cuda')
    model2.load_state(os.path.join(str(tmp_path), "checkpoint_001.tar"))
    assert model2.epoch == 1 and model2._step_host == 3
    # parameters, BN buffers and Adam moments are restored exactly
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(model._flat_m, model2._flat_m) and torch.equal(model._flat_v, model2._flat_v)
    assert float(model2._step_dev.item()) == 3.0
    x = vae_oracle.make_input(99, 8).cuda()
    noise = tuple(t.cuda() for t in vae_oracle.make_noise(99, 8))
    model.train()
    model2.train()
    la = model.train_step(x, noise=noise).item()
    lb = model2.train_step(x, noise=noise).item()
    assert abs(la - lb) <= 1e-6 * abs(la)
    # (run-to-run bitwise equality is not guaranteed: cross-CTA statistics use fp64 atomics,
    # and Adam's early, sign-like steps turn a sign flip of a near-zero gradient into a
    # 2*lr difference of that weight -- so compare in the L2 sense)
    from tests.helpers import l2_err
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert l2_err(a.double().cpu().numpy(), b.double().cpu().numpy()) <= 1e-2, k
    # get_latent: float64 [N, z], loader order, train-mode BN as in the reference (F8)
    lat = model.get_latent(loader)
    assert lat.shape == (20, 32) and lat.dtype == np.float64
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    model.train()
    with torch.no_grad():
        mu, _, _ = vae_oracle.encode(P, data[:8], train=True)
    lat2 = model.get_latent(_ListLoader(data[:8], 8))
    assert rel_err(lat2, mu.numpy()) <= 1e-3   # running stats moved between the two calls only


def test_full_size_batch_properties(vae_mod):
    """BASELINE.json's full per-GPU batch (1024), where no CPU oracle finishes in seconds:
    size-independent properties.  In eval mode (running statistics) samples are independent,
    so (i) the 1024-batch forward equals the same samples pushed through in chunks of 64 --
    different tile counts, CTA schedules and persistent-loop trip counts of every kernel --
    and (ii) the loss is additive over the chunks up to the per-batch constants (quirk F6:
    added once per batch); (iii) a train step at 1024 is reproducible run to run."""
    torch.manual_seed(5)
    model = build(vae_mod, 2)
    model.eval()
    B, chunk = 1024, 64
    x = torch.rand(B, 128, 128, device="cuda")
    ew, ed = torch.randn(B, 1, device="cuda"), torch.randn(B, 32, device="cuda")
    with torch.no_grad():
        loss_all, z_all, rec_all = model.forward(x, return_latent_rec=True, noise=(ew, ed))
        loss_sum, zs, recs = 0.0, [], []
        for i in range(0, B, chunk):
            sl = slice(i, i + chunk)
            l, z, r = model.forward(x[sl], return_latent_rec=True, noise=(ew[sl], ed[sl]))
            loss_sum += float(l.item())
            zs.append(z)
            recs.append(r)
    z_c, rec_c = np.concatenate(zs), np.concatenate(recs)
    assert rel_err(z_c, z_all) <= 1e-5
    assert rel_err(rec_c, rec_all) <= 1e-5
    n_chunks = B // chunk
    want = loss_sum - (n_chunks - 1) * model.loss_constant()
    assert abs(float(loss_all.item()) - want) <= 1e-5 * abs(want)
    # (iii) two models, same state, same batch and noise: same loss and same updated parameters
    a, b = build(vae_mod, 2), build(vae_mod, 2)
    a.train()
    b.train()
    la = float(a.train_step(x, noise=(ew, ed)).item())
    lb = float(b.train_step(x, noise=(ew, ed)).item())
    assert abs(la - lb) <= 1e-6 * abs(la)
    for (k, p), (_, q) in zip(a.state_dict().items(), b.state_dict().items()):
        if p.is_floating_point():
            d = (p.double() - q.double()).abs().max().item()
            assert d <= 2.5e-3, k      # Adam's first step is sign-like: lr-sized flips of ~0 gradients only
