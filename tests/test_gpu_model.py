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


GRAD_TOL = 1e-4       # parameter gradients vs the float64 oracle ON THE SAME ReLU PATTERN (rtol 1e-4)
MAX_FLIPS = 16        # units (of ~1.9 M per sample) allowed to sit on the other side of zero, or
FLIP_FACTOR = 4       # this many times the float32 oracle's own count, whichever is larger


def assert_grads_match(model, bufs, P, x, eps_w, eps_d, prec, tag, floor=GRAD_TOL):
    """Whole-model gradient parity, north_star's rtol 1e-4, with no allowance for kinks.

    The gradient of a ReLU network is discontinuous where a pre-activation crosses zero, and a
    unit within rounding (1e-7) of zero can land on either side depending on summation order.
    Instead of widening the tolerance for such runs, the float64 oracle is evaluated with the
    ReLU on/off pattern FORCED to the one the GPU forward produced (oracle/vae_oracle._relu):
    that is the exact gradient of the function the GPU differentiated.  Every parameter
    gradient must then be within max(1e-4, 3 x the float32 oracle's own error on the same
    pattern) in the max-norm.  The number of units whose state differs from the float64
    oracle's own forward is asserted separately: at most 16, or 4x the number the reference's
    own float32 arithmetic (the float32 oracle) flips on the same inputs -- at batch 1024 there
    are 1.9e9 units, a few hundred of which sit within float32 rounding of zero (the forward
    outputs themselves are compared elsewhere at rtol 1e-4)."""
    from tests.helpers import gpu_relu_masks, masked_oracle_grads, relu_flips
    masks = gpu_relu_masks(bufs)
    out64, g64, _, acts64, err32 = masked_oracle_grads(P, x, eps_w, eps_d, prec, masks)
    flips = relu_flips(masks, acts64)
    flips32 = err32.pop("__flips32__")
    rows = []
    for k, v in model.grad_dict().items():
        err = rel_err(v.detach().cpu().numpy(), g64[k].numpy())
        rows.append((k, err, max(floor, 3.0 * err32[k]), err32[k]))
    worst = sorted(rows, key=lambda r: -r[1] / r[2])[:6]
    report = "%s: %d ReLU flips vs float64 (float32 oracle: %d); worst gradients (err / tol / oracle-fp32 err): %s" % (
        tag, flips, flips32, ", ".join("%s %.1e/%.1e/%.1e" % r for r in worst))
    print(report)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "grad_parity.txt"), "a") as f:
            f.write(report + "\n")
    except OSError:
        pass
    bad = [r for r in rows if not r[1] <= r[2]]
    assert not bad, report
    assert flips <= max(MAX_FLIPS, FLIP_FACTOR * flips32), report
    return out64, flips


@pytest.fixture(scope="module")
def vae_mod():
    return importlib.import_module(PKG + ".models.vae")


def build(vae_mod, seed, prec=10.0, **kw):
    model = vae_mod.VAE(save_dir='', model_precision=prec, device_name='cuda', **kw)
    model.load_flat_state(vae_oracle.make_params(seed))
    return model


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32x3b", "tf32x3c", "tf32x3d", "tf32x3e"])
@pytest.mark.parametrize("name", ["vae_train_b7", "vae_eval_b7", "vae_train_b1",
                                  "vae_train_b64"])
def test_forward_backward_matches_reference_golden(vae_mod, name, precision):
    """Forward quantities against the goldens written by the reference's own code (rtol 1e-4);
    parameter gradients against the float64 oracle on the GPU's ReLU pattern (rtol 1e-4, see
    assert_grads_match) -- and, when that pattern equals the reference's (no unit on the other
    side of a kink), against the reference's golden gradients at the same tolerance."""
    g = load_golden(name)
    seed, batch, train = int(g["seed"]), int(g["batch"]), bool(g["train"])
    prec = float(g["model_precision"])
    model = build(vae_mod, seed, prec, precision=precision)
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
        model._backward_native(bufs)
        torch.cuda.synchronize()
        _, flips = assert_grads_match(model, bufs, vae_oracle.make_params(seed), x.cpu(),
                                      torch.from_numpy(g["eps_w"]), torch.from_numpy(g["eps_d"]), prec,
                                      "%s[%s]" % (name, precision))
        if flips == 0:
            for k, v in model.grad_dict().items():
                check_against_golden(g, "grad:", k, v.cpu().numpy(), GRAD_TOL, kink_factor=1.0)
    sd = model.state_dict()
    for k in g.files:
        if not k.startswith("buf:"):
            continue
        kk = k[4:]
        if kk.endswith("num_batches_tracked"):
            assert int(sd[kk]) == int(g[k]), kk
        else:
            assert rel_err(sd[kk].cpu().numpy(), g[k]) <= FWD_TOL, kk


@pytest.mark.parametrize("name", ["vae_train_b7", "vae_train_b64"])
def test_fused_mlp_path_matches_reference_golden(vae_mod, name, monkeypatch):
    """The opt-in fused small-layer path (AVA_B200_FUSED_MLP=1: csrc/mlp.cu forward / backward chains
    + ava_b200_linear_bwd_weight_multi) through the same golden / float64-oracle bars as the
    default per-layer path."""
    monkeypatch.setattr(vae_mod, "_FUSED_MLP", True)
    test_forward_backward_matches_reference_golden(vae_mod, name, "tf32x3b")


@pytest.mark.parametrize("graphs", [False, True])
def test_forked_stream_changes_scheduling_only(vae_mod, monkeypatch, graphs):
    """The dense layers' weight / bias gradients run on a second stream beside the conv backward
    (VAE._side_fork / _side_join).  Same kernels, same inputs: five train steps with and without the fork
    must end on the same parameters and losses -- a missing dependency (a gradient read before it is
    written, a scratch buffer shared across the streams, a join that comes after the optimizer step)
    shows up as a difference far above the 1e-6 that the order of the fp64 statistics atomics allows."""
    seed, B = 5, 24
    x = vae_oracle.make_input(seed, B).cuda()
    ends = []
    for side in (True, False):
        monkeypatch.setattr(vae_mod, "_SIDE_STREAM", side)
        model = build(vae_mod, seed, cuda_graphs=graphs)
        model.train()
        gen = torch.Generator(device="cuda").manual_seed(77)
        losses = []
        for step in range(5):
            ew = torch.randn(B, 1, device="cuda", generator=gen)
            ed = torch.randn(B, 32, device="cuda", generator=gen)
            losses.append(model.train_step(x, noise=(ew, ed)).clone())
        torch.cuda.synchronize()
        ends.append((model._flat_p.clone(), torch.stack(losses), model._flat_m.clone()))
    (p1, l1, m1), (p0, l0, m0) = ends
    assert float((l1 - l0).abs().max() / l0.abs().max()) <= 1e-6
    assert float((m1 - m0).abs().max() / m0.abs().max()) <= 1e-4
    # Adam's first steps are sign-like (|update| = lr whatever the gradient's size), so a parameter
    # whose gradient is ~0 may move differently: bound the mean, not the max.  (A dependency bug on
    # fc1.weight or fc8.weight -- half of all parameters each -- would move the mean by ~lr/2 = 5e-4.)
    assert float((p1 - p0).abs().mean()) <= 1e-6


def test_reduced_precision_mode_tf32(vae_mod):
    """'tf32' is the opt-in reduced-precision mode (what torch/cuDNN do by default for the
    reference's convs on a GPU) with its own stated tolerance: forward 1e-2,
    gradients within 25% in the L2 sense of the float64 oracle on the same ReLU pattern
    (BatchNorm backward amplifies the 5e-4 operand rounding).  'fp32' and 'tf32x3' hold the
    fp32 bar and are covered by test_forward_backward_matches_reference_golden."""
    from tests.helpers import gpu_relu_masks, l2_err, masked_oracle_grads
    g = load_golden("vae_train_b7")
    seed, batch = int(g["seed"]), int(g["batch"])
    prec = float(g["model_precision"])
    model = build(vae_mod, seed, prec, precision="tf32")
    model.train(True)
    x = vae_oracle.make_input(seed, batch).cuda()
    noise = (torch.from_numpy(g["eps_w"]).cuda(), torch.from_numpy(g["eps_d"]).cuda())
    bufs = model._forward_native(x, noise, True, want_grad_seed=True)
    model._backward_native(bufs)
    torch.cuda.synchronize()
    lib = importlib.import_module(PKG + "._lib").lib()
    assert lib.ava_b200_get_conv_precision() == 1
    ftol = 1e-2
    assert abs(float(bufs.loss.item()) - float(g["loss"])) <= ftol * abs(float(g["loss"]))
    assert rel_err(bufs.heads.cpu().numpy()[:, :32], g["mu"]) <= ftol
    check_against_golden(g, "", "x_rec", bufs.act[13].cpu().numpy().reshape(batch, 128, 128), ftol)
    _, g64, _, _, _ = masked_oracle_grads(vae_oracle.make_params(seed), x.cpu(), torch.from_numpy(g["eps_w"]),
                                          torch.from_numpy(g["eps_d"]), prec, gpu_relu_masks(bufs))
    for k, v in model.grad_dict().items():
        assert l2_err(v.cpu().numpy(), g64[k].numpy()) <= 0.25, k


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
        # .grad is the native backward's flat gradient buffer (scaled by grad_loss = 1)
        assert torch.equal(p.grad, model.grad_dict()[k]), k
    assert_grads_match(model, model._cur, vae_oracle.make_params(seed), x.cpu(), torch.from_numpy(g["eps_w"]),
                       torch.from_numpy(g["eps_d"]), float(g["model_precision"]), "autograd_b7")
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


class _DS:
    """Dataset stand-in with the reference SyllableDataset's indexing contract
    (ava/models/vae_dataset.py:121-145): an int gives one item, an iterable a LIST of items."""

    def __init__(self, t):
        self.t = t

    def __len__(self):
        return len(self.t)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return self.t[idx]
        if np.ndim(idx) == 0:
            return self.t[int(idx)]
        return [self.t[int(i)] for i in idx]


class _ListLoader:
    """Minimal stand-in for a DataLoader: iterable of CPU batches with a .dataset."""

    def __init__(self, data, batch_size):
        self.dataset = data if isinstance(data, _DS) else _DS(data)
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
    l0 = model.train_epoch(loader)       # (values: test_train_epoch_and_test_epoch_losses_match_oracle)
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


def _replay_noise(seed, sizes):
    """The (eps_W, eps_D) pairs the eager path draws for batches of these sizes after
    torch.manual_seed(seed): torch.randn [b,1] then [b,32] per batch (VAE._draw_noise, the
    reference's rsample order)."""
    torch.manual_seed(seed)
    return [(torch.randn(b, 1, device="cuda").cpu(), torch.randn(b, 32, device="cuda").cpu()) for b in sizes]


def test_train_epoch_and_test_epoch_losses_match_oracle(vae_mod, tmp_path):
    """train_epoch / test_epoch (ava/models/vae.py:330-385) return the reference's numbers: the
    epoch's summed per-batch losses over len(dataset), with a ragged last batch, Adam between
    the batches, and eval-mode BatchNorm (running buffers as updated by the training epoch) in
    test_epoch.  Oracle: the float64 restatement stepped through the same batches and noise."""
    seed = 6
    data = vae_oracle.make_input(seed, 20)          # batches of 8, 8, 4
    loader = _ListLoader(data, 8)
    model = vae_mod.VAE(save_dir=str(tmp_path), device_name='cuda', cuda_graphs=False)
    model.load_flat_state(vae_oracle.make_params(seed))
    torch.manual_seed(11)
    l_train = model.train_epoch(loader)
    torch.manual_seed(12)
    l_test = model.test_epoch(loader)
    assert model.epoch == 1
    P = {k: (v.double() if v.is_floating_point() else v) for k, v in vae_oracle.make_params(seed).items()}
    keys = [k for k, _ in vae_oracle.param_order()]
    st = {"step": 0, "m": {k: torch.zeros_like(P[k]) for k in keys},
          "v": {k: torch.zeros_like(P[k]) for k in keys}}
    want = 0.0
    for i, (ew, ed) in enumerate(_replay_noise(11, (8, 8, 4))):
        want += float(vae_oracle.train_step_cpu(P, st, data[8 * i:8 * i + 8].double(), ew.double(), ed.double()))
    want /= 20
    # Adam's first steps are sign-like, which amplifies fp32 noise in tiny gradients: the
    # reference's own fp32 run is 2e-4 off its float64 run after 3 steps (err32:losses in
    # tests/golden/adam_b5_s3.npz); same bar as test_train_steps_match_reference_adam_trajectory
    tol = max(1e-4, 3 * abs(float(load_golden("adam_b5_s3")["err32:losses"])))
    assert abs(l_train - want) <= tol * abs(want), (l_train, want)
    want_t = 0.0
    with torch.no_grad():
        for i, (ew, ed) in enumerate(_replay_noise(12, (8, 8, 4))):
            want_t += float(vae_oracle.forward(P, data[8 * i:8 * i + 8].double(), ew.double(), ed.double(),
                                               10.0, train=False)["loss"])
    want_t /= 20
    assert abs(l_test - want_t) <= tol * abs(want_t), (l_test, want_t)


def test_train_loop_runs_reference_schedule(vae_mod, tmp_path, capsys):
    """train_loop (ava/models/vae.py:388-430): epochs range(self.epoch, self.epoch+epochs), test
    every test_freq, checkpoint_{epoch:03d}.tar when epoch % save_freq == 0 and epoch > 0,
    visualize every vis_freq; the loss dictionary is keyed by epoch.  Four batch sizes pass
    through the model per epoch (8, ragged 4, test 8 / ragged 2, visualize 5) with the CUDA-graph
    path on: the result must equal the eager path's (graphs whose buffers were evicted would
    read freed memory -- round-1 advisor finding)."""
    seed = 4
    train = vae_oracle.make_input(seed, 20)
    test = vae_oracle.make_input(seed + 1, 10)
    results = {}
    for graphs in (True, False):
        d = tmp_path / ("g%d" % graphs)
        model = vae_mod.VAE(save_dir=str(d), device_name='cuda', cuda_graphs=graphs)
        model.load_flat_state(vae_oracle.make_params(seed))
        loaders = {'train': _ListLoader(train, 8), 'test': _ListLoader(test, 8)}
        torch.manual_seed(5)
        np.random.seed(5)
        model.train_loop(loaders, epochs=5, test_freq=2, save_freq=2, vis_freq=1)
        assert model.epoch == 5
        assert sorted(model.loss['train']) == [0, 1, 2, 3, 4]
        assert sorted(model.loss['test']) == [0, 2, 4]
        assert sorted(os.listdir(str(d))) == ['checkpoint_002.tar', 'checkpoint_004.tar'] or \
            sorted(f for f in os.listdir(str(d)) if f.endswith('.tar')) == ['checkpoint_002.tar', 'checkpoint_004.tar']
        ck = torch.load(os.path.join(str(d), 'checkpoint_004.tar'), map_location='cpu')
        assert ck['epoch'] == 5 and sorted(ck['loss']['train']) == [0, 1, 2, 3, 4]
        results[graphs] = (dict(model.loss['train']), dict(model.loss['test']),
                           {k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
    out = capsys.readouterr().out
    assert "Training: epochs 0 to 4" in out and "Training set: 20" in out and "Test set: 10" in out
    assert "Epoch: 4 Average loss:" in out and "Test loss:" in out
    # graph replay and eager launches run the same kernels on the same data: same trajectory
    # (up to the fp64-atomic ordering of the BN statistics and Adam's sign-like early steps)
    from tests.helpers import l2_err
    for e in range(5):
        a, b = results[True][0][e], results[False][0][e]
        assert abs(a - b) <= 2e-4 * abs(b), (e, a, b)
    for e in (0, 2, 4):
        a, b = results[True][1][e], results[False][1][e]
        assert abs(a - b) <= 2e-4 * abs(b), (e, a, b)
    for k, v in results[True][2].items():
        if v.is_floating_point():
            assert l2_err(v.double().numpy(), results[False][2][k].double().numpy()) <= 2e-2, k


def test_visualize_returns_specs_and_reconstructions(vae_mod, tmp_path):
    """visualize (ava/models/vae.py:475-516): np.random.choice of num_specs indices without
    replacement, dataset indexed with the index ARRAY, forward(..., return_latent_rec=True),
    returns (specs, rec_specs) as [num_specs,128,128] arrays (the pdf needs matplotlib)."""
    seed = 4
    data = vae_oracle.make_input(seed, 12)
    loader = _ListLoader(data, 4)
    model = vae_mod.VAE(save_dir=str(tmp_path), device_name='cuda')
    model.load_flat_state(vae_oracle.make_params(seed))
    model.eval()
    np.random.seed(3)
    specs, rec = model.visualize(loader, num_specs=5)
    np.random.seed(3)
    idx = np.random.choice(np.arange(12), size=5, replace=False)
    assert specs.shape == (5, 128, 128) and rec.shape == (5, 128, 128)
    assert np.array_equal(specs, data[idx].numpy())
    # the reconstruction is the decoder mean of a SAMPLE z: same encoder outputs as encode()
    with torch.no_grad():
        mu, _, _ = model.encode(data[idx].cuda())
        rec_mu = model.decode(mu).cpu().numpy().reshape(5, 128, 128)
    assert np.isfinite(rec).all() and rel_err(rec, rec_mu) < 0.5
    with pytest.raises(AssertionError):
        model.visualize(loader, num_specs=13)


def test_full_batch_1024_train_mode_matches_float64_oracle(vae_mod):
    """The benchmarked configuration itself (batch 1024, train-mode BatchNorm, precision
    'auto') against the float64 oracle: loss, mu, z, x_rec and the BN running buffers at rtol
    1e-4; every parameter gradient at rtol 1e-4 on the GPU's ReLU pattern.  At this size the
    persistent tile loops make many trips, the grids span several waves and the cross-CTA fp64
    statistics atomics see 1000+ contributions."""
    seed, B, prec = 9, 1024, 10.0
    P = vae_oracle.make_params(seed)
    model = build(vae_mod, seed, prec)
    model.train()
    x = vae_oracle.make_input(seed, B)
    ew, ed = vae_oracle.make_noise(seed, B)
    bufs = model._forward_native(x.cuda(), (ew.cuda(), ed.cuda()), True, want_grad_seed=True)
    model._backward_native(bufs)
    torch.cuda.synchronize()
    out64, flips = assert_grads_match(model, bufs, P, x, ew, ed, prec, "train_b1024[auto]")
    if flips == 0:
        # identical ReLU pattern: the forced oracle IS the reference forward
        assert abs(float(bufs.loss.item()) - float(out64["loss"])) <= FWD_TOL * abs(float(out64["loss"]))
    P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}
    nb = {}
    with torch.no_grad():
        ref = vae_oracle.forward(P64, x.double(), ew.double(), ed.double(), prec, True, nb)
    assert abs(float(bufs.loss.item()) - float(ref["loss"])) <= FWD_TOL * abs(float(ref["loss"]))
    heads = bufs.heads.cpu().numpy()
    assert rel_err(heads[:, :32], ref["mu"].numpy()) <= FWD_TOL
    assert rel_err(heads[:, 32:64], ref["u"][:, :, 0].numpy()) <= FWD_TOL
    assert rel_err(bufs.d.cpu().numpy(), ref["d"].numpy()) <= FWD_TOL
    assert rel_err(bufs.z.cpu().numpy(), ref["z"].numpy()) <= FWD_TOL
    assert rel_err(bufs.act[13].cpu().numpy().reshape(B, -1), ref["x_rec"].numpy()) <= FWD_TOL
    sd = model.state_dict()
    for k, v in nb.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        else:
            assert rel_err(sd[k].cpu().numpy(), v.numpy()) <= FWD_TOL, k


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


def test_optimizer_hyperparameters_follow_param_groups_and_state_survives_reflatten(vae_mod, tmp_path):
    """Round-1 advisor findings.  (1) Adam's lr / betas / eps come from optimizer.param_groups
    (what load_state restores from a checkpoint, ava/models/vae.py:470, and what LR schedulers
    change), also under CUDA-graph replay: with lr set to 0 a replayed step must leave the
    parameters untouched.  (2) Re-flattening the model after training (.cuda() / .to()) keeps the
    Adam step count and moments instead of re-adopting step 0."""
    seed = 8
    model = build(vae_mod, seed, cuda_graphs=True)
    model.train()
    x = vae_oracle.make_input(seed, 8).cuda()
    noise = tuple(t.cuda() for t in vae_oracle.make_noise(seed, 8))
    for _ in range(4):                                  # steps 1, 2 eager, 3 captured, 4 replayed
        model.train_step(x, noise=noise)
    assert model._graphs[8]["graph"] is not None
    before = model._flat_p.clone()
    model.optimizer.param_groups[0]['lr'] = 0.0
    model.train_step(x, noise=noise)                    # replay of the same graph
    torch.cuda.synchronize()
    assert torch.equal(model._flat_p, before)
    model.optimizer.param_groups[0]['lr'] = 5e-4
    model.train_step(x, noise=noise)
    torch.cuda.synchronize()
    step = (model._flat_p - before).abs().max().item()
    assert 0 < step <= 5.1e-4 * 10                      # an Adam step of the new size (|m|/sqrt(v) is O(1))
    assert model._step_host == 6 and float(model._step_dev.item()) == 6.0
    # (2) re-flatten: step count, moments and the optimizer's aliases survive
    m_before = model._flat_m.clone()
    model.cuda()
    assert model._step_host == 6 and float(model._step_dev.item()) == 6.0
    assert torch.equal(model._flat_m, m_before)
    st0 = model.optimizer.state[next(iter(model.parameters()))]
    assert st0['exp_avg'].data_ptr() == model._views_m[next(iter(dict(model.named_parameters())))].data_ptr()
    model.train_step(x, noise=noise)
    assert model._step_host == 7 and float(model._step_dev.item()) == 7.0
    # load_state restores the checkpoint's lr into param_groups and the native step uses it
    model.save_state(str(tmp_path / "lr.tar"))
    other = vae_mod.VAE(save_dir='', device_name='cuda', lr=1e-3)
    other.load_state(str(tmp_path / "lr.tar"))
    assert other.optimizer.param_groups[0]['lr'] == 5e-4 and other._step_host == 7
    other.train()
    other.train_step(x, noise=noise)
    assert other._hyper_host[0] == 5e-4
    with pytest.raises(NotImplementedError):
        other.optimizer.param_groups[0]['weight_decay'] = 0.1
        other.train_step(x, noise=noise)
