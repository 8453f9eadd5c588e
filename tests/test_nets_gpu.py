"""Network- and step-level parity on the GPU against the CPU oracle (same seeded weights and inputs).

Tolerances.  fp32 mode: north_star's relative L2 <= 1e-4 on outputs and losses; gradients are compared against an
fp64 evaluation of the oracle, globally (concatenated gradient vector) and per tensor, with the fp32 oracle's own
deviation from fp64 as the yardstick for ill-conditioned tensors (the cascaded UNets amplify rounding: the fp32
oracle itself is off by up to 3e-1 on individual PReLU slopes; per tensor the bound is max(5e-3, 20 x the fp32
oracle's own error), normalised by max(|g_p|, 10 % of the largest tensor norm); the global bound is 2e-3 (4e-3 for the
perceptual pass, whose 128 patches per image multiply the number of kinked units) because of activation-kink
flips, see check_grads_fp32).  bf16 mode: the 1e-2 bound holds per kernel
(tests/test_kernels_gpu.py) and for the discriminator's outputs, but NOT at network level for ANY bf16
implementation of this randomly-initialised cascade: torch's own autocast-bf16 run of the oracle deviates from fp32
by 1.2e-2 per UNet / 1.1e-1 over 6 UNets on outputs and 1.4e-1 / 8.2e-1 on gradients.  Network-level bf16 checks
are therefore calibrated in-test: error <= 1.5 x the torch-autocast deviation on the same graph + 1e-2.

The two-optimizer step is chaotic in its second half (Adam's first update is lr*sign(g), so rounding-level
differences flip signs of near-zero gradients; the fp32 and fp64 oracles already disagree by 1e-3 on d_loss), so
each optimizer pass is checked from identical states: the generator pass at step 0, the discriminator pass at
step 0 (G not yet updated), and the Adam kernel on its own (tests/test_kernels_gpu.py).
"""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

from mpgan import GAN, CasNetGenerator, Discriminator, PatchDiscriminator, _lib  # noqa: E402
from oracle.gan import GANOracle, lightning_step, sample_patch_origins, synthetic_batch  # noqa: E402
from oracle.nets import CasNetGenerator as OGen, Discriminator as ODis, PatchDiscriminator as OPatch  # noqa: E402

DEV = "cuda"


def to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def grads_of(net, prefix=""):
    return {prefix + n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def global_rel(a, b):
    num = sum(float((a[k].double().cpu() - b[k].double().cpu()).norm()) ** 2 for k in b)
    den = sum(float(b[k].double().norm()) ** 2 for k in b)
    return (num / den) ** 0.5


def check_grads_fp32(mine, ref32, ref64, what, tol=2e-3):
    """tol: a single LeakyReLU / PReLU unit whose pre-activation is within fp32 rounding of 0 takes the other branch
    of the kink in one of the two implementations and alone contributes ~|g|/sqrt(numel) = 7e-4 to the relative L2 of
    every gradient upstream of it (measured: tools/gpu_debug_dlayers.py -- the same kernels reproduce the oracle to
    1e-7 when fed the oracle's own tensors).  The 1e-4 fp32 bound is enforced where it is well defined: outputs,
    losses, and every kernel on identical inputs (tests/test_kernels_gpu.py).  For the same reason the per-tensor
    normalisation has a floor of 10 % of the largest tensor norm: one flip moves EVERY upstream tensor by an absolute
    ~7e-4 * scale, which is far above 5e-3 of a tensor whose own norm is 1e-2 * scale (which flips occur depends on
    the summation order of the BatchNorm statistics, i.e. on kernel scheduling details, not on correctness)."""
    scale = max(float(v.double().norm()) for v in ref64.values())
    contrib = sorted(((float((mine[k].double().cpu() - ref64[k].double()).norm()) / scale, k) for k in ref64),
                     reverse=True)[:4]
    glob = global_rel(mine, ref64)
    gtol = max(tol, 5 * global_rel(ref32, ref64))
    assert glob <= gtol, f"{what}: global gradient rel-L2 {glob:.3e} > {gtol:.3e}; top contributors {contrib}"
    worst = (0.0, None)
    for k, t in ref64.items():
        t = t.double()
        norm = max(float(t.norm()), 1e-1 * scale)
        e = float((mine[k].double().cpu() - t).norm()) / norm
        allow = max(2.5 * tol, 20 * float((ref32[k].double() - t).norm()) / norm)
        if e / allow > worst[0]:
            worst = (e / allow, f"{k}: err {e:.3e} allowed {allow:.3e} |g|/scale {float(t.norm()) / scale:.2e}")
    assert worst[0] <= 1.0, f"{what}: per-tensor gradient mismatch {worst[1]}"
    return glob


def check_grads_bf16(mine, ref32, autocast, what):
    yard = global_rel(autocast, ref32)
    glob = global_rel(mine, ref32)
    assert glob <= 1.5 * yard + 1e-2, f"{what}: global gradient rel-L2 {glob:.3e} vs torch-autocast {yard:.3e}"
    return glob, yard


def run_oracle(net, x, dy, mode):
    """mode: 'fp32' | 'fp64' | 'autocast' -> (output, grads, dx, net)"""
    net = copy.deepcopy(net)
    for p in net.parameters():
        p.grad = None
    if mode == "fp64":
        net = net.double()
        x, dy = x.double(), dy.double()
    x = x.clone().requires_grad_(True)
    if mode == "autocast":
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y = net(x)
        y = y.float()
    else:
        y = net(x)
    y.backward(dy)
    return y.detach(), grads_of(net), x.grad.detach(), net


def test_device_is_b200_and_library_loaded():
    assert _lib.load().mpgan_device_ok() == 1
    assert torch.cuda.get_device_capability()[0] == 10
    assert "libmpgan_sm100.so" in open("/proc/self/maps").read()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dims,size,batch,nblocks", [(2, 64, 3, 6), (2, 64, 2, 1), (3, 16, 2, 6)])
def test_generator_forward_backward(precision, dims, size, batch, nblocks):
    shape = (1,) + (size,) * dims
    torch.manual_seed(0)
    ref = OGen(shape, nblocks, dims)
    mine = CasNetGenerator(shape, n_unet_blocks=nblocks, precision=precision)
    mine.load_state_dict(ref.state_dict())
    x = synthetic_batch(batch, dims, size, seed=1)["t1w"]
    dy = synthetic_batch(batch, dims, size, seed=5)["t2w"]
    y32, g32, _, net32 = run_oracle(ref, x, dy, "fp32")
    y = mine(x.to(DEV))
    assert y.shape == x.shape and y.dtype == torch.float32
    y.backward(dy.to(DEV))
    if precision == "fp32":
        _, g64, _, _ = run_oracle(ref, x, dy, "fp64")
        assert rel_l2(y, y32) <= 1e-4
        check_grads_fp32(grads_of(mine), g32, g64, f"generator fp32 {dims}d x{nblocks}")
        btol = 1e-4
    else:
        ya, ga, _, _ = run_oracle(ref, x, dy, "autocast")
        yard = rel_l2(ya, y32)
        assert rel_l2(y, y32) <= 1.5 * yard + 1e-2, (rel_l2(y, y32), yard)
        check_grads_bf16(grads_of(mine), g32, ga, f"generator bf16 {dims}d x{nblocks}")
        btol = 1.5 * yard + 1e-2
    # BN running statistics after one training-mode forward
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), net32.named_buffers()):
        assert n1 == n2
        if b2.dtype.is_floating_point:
            assert rel_l2(b1, b2) <= btol, n1
        else:
            assert int(b1) == int(b2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dims,size,batch", [(2, 64, 3), (2, 71, 2), (3, 24, 2)])
def test_discriminator_forward_backward(precision, dims, size, batch):
    shape = (1,) + (size,) * dims
    torch.manual_seed(0)
    ref = ODis(shape, dims=dims, spatial=size)
    mine = Discriminator(shape, spatial=size, precision=precision)
    mine.load_state_dict(ref.state_dict())
    x = synthetic_batch(batch, dims, size, seed=1)["t1w"]
    dp = torch.linspace(-1, 1, batch).reshape(batch, 1)
    p32, g32, dx32, net32 = run_oracle(ref, x, dp, "fp32")
    xd = x.to(DEV).requires_grad_(True)
    p = mine(xd)
    assert p.shape == (batch, 1)
    p.backward(dp.to(DEV))
    if precision == "fp32":
        _, g64, _, _ = run_oracle(ref, x, dp, "fp64")
        assert rel_l2(p, p32) <= 1e-4
        check_grads_fp32(grads_of(mine), g32, g64, f"discriminator fp32 {dims}d")
        assert rel_l2(xd.grad, dx32) <= 2e-3
        assert rel_l2(mine.model_conv[1].running_var, net32.model_conv[1].running_var) <= 1e-4
    else:
        pa, ga, dxa, _ = run_oracle(ref, x, dp, "autocast")
        assert rel_l2(p, p32) <= 1e-2                       # discriminator probabilities: north_star's bf16 bound
        check_grads_bf16(grads_of(mine), g32, ga, f"discriminator bf16 {dims}d")
        assert rel_l2(xd.grad, dx32) <= 1.5 * rel_l2(dxa, dx32) + 1e-2
        assert rel_l2(mine.model_conv[1].running_var, net32.model_conv[1].running_var) <= 1e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_eval_mode_generator(precision):
    torch.manual_seed(0)
    ref = OGen((1, 64, 64), 6, 2)
    with torch.no_grad():
        for _ in range(2):
            ref(synthetic_batch(2, 2, 64, seed=7)["t1w"])   # move the running stats off their init
    mine = CasNetGenerator((1, 64, 64), precision=precision)
    mine.load_state_dict(ref.state_dict())
    ref.eval(), mine.eval()
    x = synthetic_batch(4, 2, 64, seed=1)["t1w"]
    with torch.no_grad():
        y32 = ref(x)
        y = mine(x.to(DEV))
        if precision == "fp32":
            assert rel_l2(y, y32) <= 1e-4
        else:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                ya = copy.deepcopy(ref)(x).float()
            assert rel_l2(y, y32) <= 1.5 * rel_l2(ya, y32) + 1e-2
    assert int(mine.model[0].model[0].conv.unit0.norm.num_batches_tracked) == 2


def _oracle_pass(ora, batch, opt_idx, **kw):
    nets = (ora.generator, ora.discriminator)
    for i, net in enumerate(nets):
        for p in net.parameters():
            p.requires_grad_(i == opt_idx)
            p.grad = None
    loss = ora.training_step(batch, 0, opt_idx, **kw)
    loss.backward()
    g = grads_of(nets[opt_idx])
    for p in ora.parameters():
        p.requires_grad_(True)
    return float(loss), g


def _autocast_pass(ora, batch, opt_idx, **kw):
    ora = copy.deepcopy(ora)
    nets = (ora.generator, ora.discriminator)
    for i, net in enumerate(nets):
        for p in net.parameters():
            p.requires_grad_(i == opt_idx)
            p.grad = None
    with torch.autocast("cpu", dtype=torch.bfloat16):
        loss = ora.training_step(batch, 0, opt_idx, **kw)
    loss.float().backward()
    return float(loss), grads_of(nets[opt_idx])


def _my_pass(mine, dbatch, opt_idx, **kw):
    nets = (mine.generator, mine.discriminator)
    for i, net in enumerate(nets):
        net.runtime.zero_grad()
        for p in net.parameters():
            p.requires_grad_(i == opt_idx)
    loss = mine.training_step(dbatch, 0, opt_idx, **kw)
    loss.backward()
    g = grads_of(nets[opt_idx])
    for p in mine.parameters():
        p.requires_grad_(True)
    return float(loss), g


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("opt_idx", [0, 1])
def test_training_step_pass_vs_oracle(precision, opt_idx):
    """One optimizer pass of GAN_final.py's training_step from identical states: loss, logged scalars, every
    parameter gradient, BN running buffers."""
    B, S = 2, 64
    torch.manual_seed(0)
    ora = GANOracle("final", dims=2, spatial=S)
    mine = GAN(1, S, S, precision=precision)
    mine.load_state_dict(ora.state_dict())
    batch = synthetic_batch(B, 2, S, seed=1)
    ora64 = copy.deepcopy(ora).double()
    aloss, agrads = _autocast_pass(ora, batch, opt_idx) if precision == "bf16" else (None, None)
    rloss, rgrads = _oracle_pass(ora, batch, opt_idx)
    mloss, mgrads = _my_pass(mine, to_dev(batch), opt_idx)
    name = "generator" if opt_idx == 0 else "discriminator"
    if precision == "fp32":
        _, g64 = _oracle_pass(ora64, {k: v.double() for k, v in batch.items()}, opt_idx)
        assert abs(mloss - rloss) <= 1e-4 * abs(rloss)
        check_grads_fp32(mgrads, rgrads, g64, f"{name} pass fp32")
        btol = 1e-4
    else:
        assert abs(mloss - rloss) <= 1.5 * abs(aloss - rloss) + 1e-2 * abs(rloss)
        check_grads_bf16(mgrads, rgrads, agrads, f"{name} pass bf16")
        btol = 0.2
    keys = ("g_adv_loss", "g_recon_loss", "g_loss") if opt_idx == 0 else ("d_loss",)
    for k in keys:
        assert abs(float(mine.logged[k]) - float(ora.logged[k])) <= (3e-4 if precision == "fp32" else 5e-2), k
    if opt_idx == 0 and precision == "fp32":   # golden: numbers frozen in the authoring container
        fix = torch.load(os.path.join(GOLDEN, "oracle_final_step_2d_64.pt"), weights_only=False)
        assert abs(mloss - float(fix["g_loss"])) <= 3e-4 * float(fix["g_loss"])
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ora.named_buffers()):
        if b2.dtype.is_floating_point:
            assert rel_l2(b1, b2) <= btol, n1
        else:
            assert int(b1) == int(b2), n1


def test_host_fed_step_equals_direct_replay():
    """mpgan.HostFedStep (pinned host batches, H2D on a copy stream into double-buffered staging, D2D into the graph's
    inputs, replay, async D2H of the losses) runs the same steps as copying into the static inputs by hand."""
    from mpgan import HostFedStep
    B, S = 2, 64
    torch.manual_seed(0)
    kw = dict(precision="fp32", g_lr=0.0, d_lr=0.0)
    a, b = GAN(1, S, S, **kw), GAN(1, S, S, **kw)
    b.load_state_dict(a.state_dict())
    hosts = [{k: v.pin_memory() for k, v in synthetic_batch(B, 2, S, seed=s).items()} for s in (1, 2, 3, 4, 5)]
    graph, static, logs = a.capture(to_dev(hosts[0]))
    want = []
    for h in hosts:
        static["t1w"].copy_(h["t1w"]), static["t2w"].copy_(h["t2w"])
        graph.replay()
        want.append(logs.cpu().clone())
    fed = HostFedStep(b, to_dev(hosts[0]))
    got = []
    for h in hosts:
        lh = fed.step(h)
        torch.cuda.synchronize()
        got.append(lh.clone())
    for w, g in zip(want, got):
        assert torch.allclose(g, w, rtol=2e-5, atol=1e-6)
    # back-to-back steps without host synchronisation: the double-buffered staging must not be overwritten early
    c = GAN(1, S, S, **kw)
    c.load_state_dict(a.state_dict())
    fed2 = HostFedStep(c, to_dev(hosts[0]))
    for h in hosts:
        lh = fed2.step(h)
    torch.cuda.synchronize()
    assert torch.allclose(lh, want[-1], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_step_and_graph_replay_equal_the_protocol(precision):
    """fit_batch (reference protocol through autograd), fused_step (static plan) and the captured CUDA graph run the
    same arithmetic.  Learning rates are 0 so the comparison is not at the mercy of Adam's sign flips; the Adam
    kernel still runs (step counter, moments)."""
    B, S = 2, 64
    torch.manual_seed(0)
    kw = dict(precision=precision, g_lr=0.0, d_lr=0.0)
    a, b, c = GAN(1, S, S, **kw), GAN(1, S, S, **kw), GAN(1, S, S, **kw)
    b.load_state_dict(a.state_dict()), c.load_state_dict(a.state_dict())
    batches = [to_dev(synthetic_batch(B, 2, S, seed=s)) for s in (1, 2)]
    probe = {}

    def grab(name, net):
        probe[name] = net.runtime.grad.clone()

    la = [a.fit_batch(bt) for bt in batches]
    lb = [b.fused_step(bt).clone() for bt in batches]
    graph, static, logs = c.capture(batches[0])
    lc = []
    for bt in batches:
        static["t1w"].copy_(bt["t1w"]), static["t2w"].copy_(bt["t2w"])
        graph.replay()
        lc.append(logs.clone())
    torch.cuda.synchronize()
    # bf16: atomically-accumulated BN statistics differ in the last bit run to run; the cascade amplifies that
    tol = 2e-5 if precision == "fp32" else 2e-2
    for i in range(2):
        g_loss, d_loss = float(la[i][0]), float(la[i][1])
        assert abs(float(lb[i][0] + lb[i][1]) - g_loss) <= tol * abs(g_loss)
        assert abs(float(lb[i][2] + lb[i][3]) - d_loss) <= tol * abs(d_loss)
        assert torch.allclose(lb[i], lc[i], rtol=tol, atol=1e-6)
    # BN running buffers (2 G + 3 D train-mode forwards per step) agree across the three drivers
    for (n1, b1), (_, b2), (_, b3) in zip(a.named_buffers(), b.named_buffers(), c.named_buffers()):
        if b1.dtype.is_floating_point:
            assert rel_l2(b2, b1) <= tol and rel_l2(b3, b1) <= tol, n1
        else:
            assert int(b1) == int(b2) == int(b3), n1
    for m in (a, b, c):
        assert int(m.generator.runtime.adam_state[0].view(torch.int32)) == 2
        assert int(m.discriminator.runtime.adam_state[0].view(torch.int32)) == 2
    # gradients the fused path hands to Adam == the protocol's, from identical states
    a2, b2 = GAN(1, S, S, **kw), GAN(1, S, S, **kw)
    a2.load_state_dict(a.state_dict()), b2.load_state_dict(a.state_dict())
    b2.fused_step(batches[1], grad_probe=grab)
    _my_pass(a2, batches[1], 0)
    want, got = a2.generator.runtime.grad, probe["generator"]
    assert want.shape == got.shape
    if precision == "fp32":   # (the bf16 cascade gradient is noise-dominated: see the module docstring)
        assert rel_l2(got, want) <= 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_generator_pass_vs_oracle(precision):
    """BASELINE.json configs[1] at its FULL size (batch 32 of 256x256): the generator pass of the training step --
    generator output, discriminator probabilities and both losses -- against the CPU oracle on the same seeded weights
    and inputs.  fp32 mode: north_star's relative L2 <= 1e-4 throughout.  bf16 mode: <= 1e-2 for the discriminator
    (fed the oracle's generator output); the six cascaded randomly-initialised UNets are checked against the bound
    any bf16 implementation reaches on them (torch's own autocast run deviates by 1.1e-1, module docstring)."""
    B, S = 32, 256
    torch.manual_seed(0)
    ora = GANOracle("final", dims=2, spatial=S)
    mine = GAN(1, S, S, precision=precision)
    mine.load_state_dict(ora.state_dict())
    batch = synthetic_batch(B, 2, S, seed=1)
    with torch.no_grad():
        gen_ref = ora.generator(batch["t1w"])
        p_ref = ora.discriminator(gen_ref)
        adv_ref = float(ora.adversarial_loss(p_ref, torch.ones(B, 1)))
        rec_ref = float(ora.reconstruction_loss(gen_ref, batch["t2w"]))
    d = to_dev(batch)
    with torch.no_grad():
        gen = mine.generator(d["t1w"])
        p_on_ref = mine.discriminator(gen_ref.to(DEV))
    tol_d = 1e-4 if precision == "fp32" else 1e-2
    tol_g = 1e-4 if precision == "fp32" else 1.5e-1
    assert rel_l2(p_on_ref, p_ref) <= tol_d
    assert rel_l2(gen, gen_ref) <= tol_g
    logs = mine.fused_step(d)
    torch.cuda.synchronize()
    assert torch.isfinite(logs).all()
    assert abs(float(logs[0]) - adv_ref) <= (1e-4 if precision == "fp32" else 1e-1) * abs(adv_ref) + 1e-6
    assert abs(float(logs[1]) - rec_ref) <= (1e-4 if precision == "fp32" else 5e-2) * abs(rec_ref)


def test_full_step_tracks_the_oracle():
    """Three full two-optimizer steps with the reference's learning rates (chaotic regime: loose bounds) -- the
    reconstruction loss follows the oracle and the adversarial loss saturates towards the -100 log clamp like the
    reference's own checkpoints record."""
    B, S = 2, 64
    torch.manual_seed(0)
    ora = GANOracle("final", dims=2, spatial=S)
    mine = GAN(1, S, S, precision="fp32")
    mine.load_state_dict(ora.state_dict())
    batch = synthetic_batch(B, 2, S, seed=1)
    opts, _ = ora.configure_optimizers()
    dbatch = to_dev(batch)
    for step in range(3):
        ref = lightning_step(ora, opts, batch, step)
        got = mine.fit_batch(dbatch, step)
        assert all(bool(torch.isfinite(g)) for g in got)
        assert abs(float(mine.logged["g_recon_loss"]) - float(ora.logged["g_recon_loss"])) <= 0.02
        if step == 0:
            assert abs(float(got[0]) - float(ref[0])) <= 1e-4 * float(ref[0])
            assert abs(float(got[1]) - float(ref[1])) <= 3e-2 * float(ref[1])
        else:
            assert float(mine.logged["g_adv_loss"]) > 3.0 and float(ora.logged["g_adv_loss"]) > 3.0


_ACT_SHAPES = [(6, 64, 14, 14)] * 3 + [(6, 128, 12, 12)] * 3 + [(6, 256, 10, 10)] * 3 + [(6, 512, 8, 8)] * 3 + \
              [(6, 32768), (6, 64), (6, 1), (6, 1)]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_patch_discriminator_activations(precision):
    torch.manual_seed(0)
    ref = OPatch((1, 16, 16), dims=2, spatial=16)
    mine = PatchDiscriminator((1, 16, 16), precision=precision)
    mine.load_state_dict(ref.state_dict())
    x = synthetic_batch(6, 2, 16, seed=1)["t1w"]
    w = [torch.randn(s, generator=torch.Generator().manual_seed(k)) for k, s in enumerate(_ACT_SHAPES)]

    def objective(net, xin, cast=lambda t: t):
        v, a = net(xin)
        return v, a, v.float().sum() + sum((a[k].float() * cast(w[k])).sum() for k in range(16))

    def oracle(mode):
        net = copy.deepcopy(ref)
        xin = x.clone().requires_grad_(True)
        if mode == "fp64":
            net, xin = net.double(), x.double().requires_grad_(True)
            v, a = net(xin)
            obj = v.sum() + sum((a[k] * w[k].double()).sum() for k in range(16))
        elif mode == "autocast":
            with torch.autocast("cpu", dtype=torch.bfloat16):
                v, a, obj = objective(net, xin)
        else:
            v, a, obj = objective(net, xin)
        obj.backward()
        return v.detach(), {k: t.detach() for k, t in a.items()}, grads_of(net), xin.grad.detach()

    v32, a32, g32, dx32 = oracle("fp32")
    xd = x.to(DEV).requires_grad_(True)
    v, a, obj = objective(mine, xd, lambda t: t.to(DEV))
    obj.backward()
    assert len(a) == 16
    for k in range(16):
        assert a[k].shape == a32[k].shape, k
    if precision == "fp32":
        _, _, g64, _ = oracle("fp64")
        assert rel_l2(v, v32) <= 1e-4
        for k in range(16):
            assert rel_l2(a[k], a32[k]) <= 1e-4, k
        check_grads_fp32(grads_of(mine), g32, g64, "patch D fp32")
        assert rel_l2(xd.grad, dx32) <= 2e-3
    else:
        va, aa, ga, dxa = oracle("autocast")
        assert rel_l2(v, v32) <= 1e-2
        for k in range(16):
            assert rel_l2(a[k], a32[k]) <= 1.5 * rel_l2(aa[k].float(), a32[k]) + 1e-2, k
        check_grads_bf16(grads_of(mine), g32, ga, "patch D bf16")
        assert rel_l2(xd.grad, dx32) <= 1.5 * rel_l2(dxa, dx32) + 1e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("opt_idx", [0, 1])
def test_perceptual_training_step_pass_vs_oracle(precision, opt_idx):
    """test_runs/GAN.py training_step (patch gather + patch D + perceptual + L1 on patches), one pass."""
    B, S, NS = 2, 32, 6
    torch.manual_seed(0)
    ora = GANOracle("perceptual", dims=2, spatial=S, num_samples=NS)
    mine = GAN(1, S, S, variant="perceptual", num_samples=NS, precision=precision)
    mine.load_state_dict(ora.state_dict())
    batch = synthetic_batch(B, 2, S, seed=1)
    origins = sample_patch_origins(np.random.RandomState(2), B, NS, (S, S), 16)
    ora64 = copy.deepcopy(ora).double()
    aloss, agrads = _autocast_pass(ora, batch, opt_idx, patch_origins=origins) if precision == "bf16" else (None, None)
    rloss, rgrads = _oracle_pass(ora, batch, opt_idx, patch_origins=origins)
    mloss, mgrads = _my_pass(mine, to_dev(batch), opt_idx, patch_origins=origins)
    if precision == "fp32":
        _, g64 = _oracle_pass(ora64, {k: v.double() for k, v in batch.items()}, opt_idx, patch_origins=origins)
        assert abs(mloss - rloss) <= 1e-4 * abs(rloss)
        check_grads_fp32(mgrads, rgrads, g64, f"perceptual pass {opt_idx} fp32", tol=4e-3)
        if opt_idx == 0:
            for k in ("g_perceptual_loss", "g_adv_loss", "g_recon_loss"):
                r = float(ora.logged[k])
                assert abs(float(mine.logged[k]) - r) <= 1e-4 * max(abs(r), 1e-6), k
    else:
        assert abs(mloss - rloss) <= 1.5 * abs(aloss - rloss) + 1e-2 * abs(rloss)
        check_grads_bf16(mgrads, rgrads, agrads, f"perceptual pass {opt_idx} bf16")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_perceptual_step_equals_the_protocol(precision):
    """test_runs/GAN.py step: the static kernel sequence (fused_step, activations in their internal layout, perceptual
    gradients injected in place) and its CUDA graph run the arithmetic of the autograd-driven protocol (fit_batch)
    and of the oracle.  Learning rates 0, like the GAN_final.py twin of this test."""
    B, S, NS = 2, 32, 6
    torch.manual_seed(0)
    kw = dict(variant="perceptual", num_samples=NS, precision=precision, lr=0.0)
    a, b, c = GAN(1, S, S, **kw), GAN(1, S, S, **kw), GAN(1, S, S, **kw)
    b.load_state_dict(a.state_dict()), c.load_state_dict(a.state_dict())
    batch = to_dev(synthetic_batch(B, 2, S, seed=1))
    origins = sample_patch_origins(np.random.RandomState(2), B, NS, (S, S), 16)
    o_dev = torch.as_tensor(np.asarray(origins).reshape(-1, 2), dtype=torch.int32, device=DEV)
    la = a.fit_batch(batch, patch_origins=origins)
    probe = {}
    lb = b.fused_step(dict(batch, origins=o_dev), grad_probe=lambda name, net: probe.__setitem__(name, net.runtime.grad.clone()))
    graph, static, logs = c.capture(dict(batch, origins=o_dev))
    graph.replay()
    torch.cuda.synchronize()
    tol = 2e-5 if precision == "fp32" else 2e-2
    g_loss, d_loss = float(la[0]), float(la[1])
    assert abs(float(lb[0] + lb[1] + lb[2]) - g_loss) <= tol * abs(g_loss)
    assert abs(float(lb[3] + lb[4]) - d_loss) <= tol * abs(d_loss)
    assert abs(float(lb[2]) - float(a.logged["g_perceptual_loss"])) <= tol * abs(float(a.logged["g_perceptual_loss"])) + 1e-9
    assert torch.allclose(lb, logs, rtol=tol, atol=1e-6)
    for (n1, b1), (_, b2), (_, b3) in zip(a.named_buffers(), b.named_buffers(), c.named_buffers()):
        if b1.dtype.is_floating_point:
            assert rel_l2(b2, b1) <= tol and rel_l2(b3, b1) <= tol, n1
    # the gradients handed to Adam == the protocol's (fp32: tight; bf16: the generator cascade is noise dominated)
    a2 = GAN(1, S, S, **kw)
    a2.load_state_dict(a.state_dict())
    _my_pass(a2, batch, 1, patch_origins=origins)
    want, got = a2.discriminator.runtime.grad, probe["discriminator"]
    assert rel_l2(got, want) <= (1e-5 if precision == "fp32" else 3e-2)
    if precision == "fp32":
        a3 = GAN(1, S, S, **kw)
        a3.load_state_dict(a.state_dict())
        _my_pass(a3, batch, 0, patch_origins=origins)
        assert rel_l2(probe["generator"], a3.generator.runtime.grad) <= 1e-5


def test_reference_literal_3d_patch_discriminator_golden():
    """Output of the reference's own test_runs/GAN.py Discriminator (3-D, 16^3), recorded in tests/golden."""
    fix = torch.load(os.path.join(GOLDEN, "ref_patch_discriminator_3d.pt"), weights_only=False)
    torch.manual_seed(fix["seed_weights"])
    mine = PatchDiscriminator((1, 16, 16, 16), precision="fp32")
    x = torch.rand(fix["input_shape"], generator=torch.Generator().manual_seed(fix["seed_input"])) * 2 - 1
    v, acts = mine(x.to(DEV))
    assert rel_l2(v, fix["validity"]) <= 1e-4
    sums = torch.stack([acts[k].double().sum() for k in range(16)]).float().cpu()
    assert rel_l2(sums, fix["act_sums"]) <= 1e-3


def test_reference_literal_3d_generator_golden():
    fix = torch.load(os.path.join(GOLDEN, "ref_generator_3d.pt"), weights_only=False)
    torch.manual_seed(fix["seed_weights"])
    mine = CasNetGenerator((1, 16, 16, 16), precision="fp32")
    assert list(mine.state_dict().keys()) == fix["state_keys"]
    x = torch.rand(fix["input_shape"], generator=torch.Generator().manual_seed(fix["seed_input"])) * 2 - 1
    assert rel_l2(mine(x.to(DEV)), fix["out"]) <= 1e-4
