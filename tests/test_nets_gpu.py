"""Network- and step-level parity on the GPU against the CPU oracle (same seeded weights and inputs).
Tolerances are north_star's: relative L2 <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode."""
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
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def grads_of(net, prefix=""):
    return {prefix + n: p.grad.detach().clone() for n, p in net.named_parameters()}


def check_grads(mine, ref, tol, what):
    """Per-tensor relative L2 for tensors that carry signal; tensors whose reference gradient is numerically
    zero (conv bias in front of a training-mode BatchNorm) are compared on an absolute scale instead."""
    scale = max(float(v.double().norm()) for v in ref.values())
    worst = (0.0, None)
    for k, r in ref.items():
        m = mine[k].float().cpu()
        rn = float(r.double().norm())
        if rn > 1e-3 * scale:
            e = rel_l2(m, r)
        else:
            e = float((m.double() - r.double()).norm()) / scale
        if e > worst[0]:
            worst = (e, k)
    assert worst[0] <= tol, f"{what}: worst gradient mismatch {worst[0]:.3e} at {worst[1]}"
    return worst


def test_device_is_b200_and_library_loaded():
    assert _lib.load().mpgan_device_ok() == 1
    assert torch.cuda.get_device_capability()[0] == 10
    assert "libmpgan_sm100.so" in open("/proc/self/maps").read()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dims,size,batch", [(2, 64, 3), (3, 16, 2)])
def test_generator_forward_backward(precision, dims, size, batch):
    shape = (1,) + (size,) * dims
    torch.manual_seed(0)
    ref = OGen(shape, 6, dims)
    mine = CasNetGenerator(shape, precision=precision)
    mine.load_state_dict(ref.state_dict())
    x = synthetic_batch(batch, dims, size, seed=1)["t1w"]
    y_ref = ref(x)
    dy = synthetic_batch(batch, dims, size, seed=5)["t2w"]
    y_ref.backward(dy)
    y = mine(x.to(DEV))
    assert y.shape == x.shape and y.dtype == torch.float32
    tol = TOL[precision] * (3 if precision == "bf16" else 1)   # 6 cascaded UNets accumulate bf16 rounding
    assert rel_l2(y, y_ref) <= tol
    y.backward(dy.to(DEV))
    check_grads(grads_of(mine), grads_of(ref), 10 * tol, f"generator {precision} {dims}d")
    # BN running statistics after one training-mode forward
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ref.named_buffers()):
        assert n1 == n2
        if b2.dtype.is_floating_point:
            assert rel_l2(b1, b2) <= tol, n1
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
    x = synthetic_batch(batch, dims, size, seed=1)["t1w"].requires_grad_(True)
    p_ref = ref(x)
    dp = torch.linspace(-1, 1, batch).reshape(batch, 1)
    p_ref.backward(dp)
    xd = x.detach().to(DEV).requires_grad_(True)
    p = mine(xd)
    assert p.shape == (batch, 1)
    tol = TOL[precision]
    assert rel_l2(p, p_ref) <= tol
    p.backward(dp.to(DEV))
    check_grads(grads_of(mine), grads_of(ref), 5 * tol, f"discriminator {precision} {dims}d")
    assert rel_l2(xd.grad, x.grad) <= 5 * tol
    assert rel_l2(mine.model_conv[1].running_var, ref.model_conv[1].running_var) <= tol


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
        assert rel_l2(mine(x.to(DEV)), ref(x)) <= 3 * TOL[precision]
    assert int(mine.model[0].model[0].conv.unit0.norm.num_batches_tracked) == 2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_step_protocol_vs_oracle(precision):
    """The reference's own protocol (training_step + backward + optimizer.step per optimizer index) at step 0:
    losses, every parameter gradient, parameters after the step, BN buffers."""
    B, S = 2, 64
    torch.manual_seed(0)
    ora = GANOracle("final", dims=2, spatial=S)
    mine = GAN(1, S, S, precision=precision)
    mine.load_state_dict(ora.state_dict())
    batch = synthetic_batch(B, 2, S, seed=1)
    opts, _ = ora.configure_optimizers()
    ref_grads = {}
    ref_losses = lightning_step(ora, opts, batch, 0, keep_grads=ref_grads)
    # mine: same loop, capturing gradients before the optimizer consumes them
    dbatch = to_dev(batch)
    my_opts, _ = mine.configure_optimizers()
    nets = (mine.generator, mine.discriminator)
    my_grads, my_losses = {}, []
    for opt_idx, opt in enumerate(my_opts):
        for i, net in enumerate(nets):
            for p in net.parameters():
                p.requires_grad_(i == opt_idx)
        loss = mine.training_step(dbatch, 0, opt_idx)
        loss.backward()
        my_grads.update(grads_of(nets[opt_idx], "generator." if opt_idx == 0 else "discriminator."))
        opt.step(), opt.zero_grad()
        for net in nets:
            for p in net.parameters():
                p.requires_grad_(True)
        my_losses.append(float(loss))
    tol = TOL[precision]
    assert abs(my_losses[0] - float(ref_losses[0])) <= 3 * tol * abs(float(ref_losses[0]))
    assert abs(my_losses[1] - float(ref_losses[1])) <= 3 * tol * abs(float(ref_losses[1]))
    for k in ("g_adv_loss", "g_recon_loss", "d_loss"):
        assert abs(float(mine.logged[k]) - float(ora.logged[k])) <= 3 * tol * max(1.0, abs(float(ora.logged[k]))), k
    gtol = 10 * tol if precision == "bf16" else 5 * tol
    check_grads({k: v for k, v in my_grads.items() if k.startswith("generator.")},
                {k: v for k, v in ref_grads.items() if k.startswith("generator.")}, gtol, "G grads")
    check_grads({k: v for k, v in my_grads.items() if k.startswith("discriminator.")},
                {k: v for k, v in ref_grads.items() if k.startswith("discriminator.")}, gtol, "D grads")
    # golden (oracle numbers frozen in the authoring container)
    fix = torch.load(os.path.join(GOLDEN, "oracle_final_step_2d_64.pt"), weights_only=False)
    assert abs(my_losses[0] - float(fix["g_loss"])) <= 3 * tol * float(fix["g_loss"])
    assert abs(my_losses[1] - float(fix["d_loss"])) <= 3 * tol * float(fix["d_loss"])
    # BN buffers after the whole step (G forwarded twice, D three times)
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ora.named_buffers()):
        if b2.dtype.is_floating_point:
            assert rel_l2(b1, b2) <= 3 * tol, n1
        else:
            assert int(b1) == int(b2), n1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_step_equals_protocol_and_graph_replay(precision):
    B, S = 2, 64
    torch.manual_seed(0)
    a = GAN(1, S, S, precision=precision)
    b = GAN(1, S, S, precision=precision)
    b.load_state_dict(a.state_dict())
    c = GAN(1, S, S, precision=precision)
    c.load_state_dict(a.state_dict())
    batch = to_dev(synthetic_batch(B, 2, S, seed=1))
    batch2 = to_dev(synthetic_batch(B, 2, S, seed=2))
    la = [a.fit_batch(batch), a.fit_batch(batch2)]
    lb = [b.fused_step(batch).clone(), b.fused_step(batch2).clone()]
    graph, static, logs = c.capture(batch)
    lc = []
    for bt in (batch, batch2):
        static["t1w"].copy_(bt["t1w"]), static["t2w"].copy_(bt["t2w"])
        graph.replay()
        lc.append(logs.clone())
    torch.cuda.synchronize()
    for i in range(2):
        g_loss, d_loss = float(la[i][0]), float(la[i][1])
        assert abs(float(lb[i][0] + lb[i][1]) - g_loss) <= 2e-3 * abs(g_loss)
        assert abs(float(lb[i][2] + lb[i][3]) - d_loss) <= 2e-3 * abs(d_loss)
        assert torch.allclose(lb[i], lc[i], rtol=2e-3, atol=1e-5)
    # parameters after two steps agree between the three drivers (atomics make wgrad order non-deterministic)
    for (n1, p1), (n2, p2), (n3, p3) in zip(a.named_parameters(), b.named_parameters(), c.named_parameters()):
        assert rel_l2(p2, p1) <= 5e-3, n1
        assert rel_l2(p3, p1) <= 5e-3, n1
    assert int(c.generator.runtime.adam_state[0].view(torch.int32)) == 2


def test_saturated_regime_matches_oracle():
    """After the first optimizer step D saturates: the BCE clamp must reproduce the oracle's loss."""
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
        for r, g in zip(ref, got):
            assert abs(float(g) - float(r)) <= 5e-3 * max(1.0, abs(float(r))), (step, float(g), float(r))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_patch_discriminator_activations(precision):
    torch.manual_seed(0)
    ref = OPatch((1, 16, 16), dims=2, spatial=16)
    mine = PatchDiscriminator((1, 16, 16), precision=precision)
    mine.load_state_dict(ref.state_dict())
    x = synthetic_batch(6, 2, 16, seed=1)["t1w"].requires_grad_(True)
    v_ref, a_ref = ref(x)
    xd = x.detach().to(DEV).requires_grad_(True)
    v, a = mine(xd)
    tol = TOL[precision]
    assert len(a) == 16 and rel_l2(v, v_ref) <= tol
    for k in range(16):
        assert a[k].shape == a_ref[k].shape, k
        assert rel_l2(a[k], a_ref[k]) <= 2 * tol, k
    # perceptual-style objective touching every activation
    w = [torch.randn(a_ref[k].shape, generator=torch.Generator().manual_seed(k)) for k in range(16)]
    (v_ref.sum() + sum((a_ref[k] * w[k]).sum() for k in range(16))).backward()
    (v.sum() + sum((a[k] * w[k].to(DEV)).sum() for k in range(16))).backward()
    assert rel_l2(xd.grad, x.grad) <= 10 * tol
    check_grads(grads_of(mine), grads_of(ref), 10 * tol, "patch D")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_perceptual_training_step_vs_oracle(precision):
    B, S, NS = 2, 32, 6
    torch.manual_seed(0)
    ora = GANOracle("perceptual", dims=2, spatial=S, num_samples=NS)
    mine = GAN(1, S, S, variant="perceptual", num_samples=NS, precision=precision)
    mine.load_state_dict(ora.state_dict())
    batch = synthetic_batch(B, 2, S, seed=1)
    origins = sample_patch_origins(np.random.RandomState(2), B, NS, (S, S), 16)
    opts, _ = ora.configure_optimizers()
    ref_grads = {}
    ref_losses = lightning_step(ora, opts, batch, 0, patch_origins=origins, keep_grads=ref_grads)
    dbatch = to_dev(batch)
    my_opts, _ = mine.configure_optimizers()
    nets = (mine.generator, mine.discriminator)
    my_grads, my_losses = {}, []
    for opt_idx, opt in enumerate(my_opts):
        for i, net in enumerate(nets):
            for p in net.parameters():
                p.requires_grad_(i == opt_idx)
        loss = mine.training_step(dbatch, 0, opt_idx, patch_origins=origins)
        loss.backward()
        my_grads.update(grads_of(nets[opt_idx], "generator." if opt_idx == 0 else "discriminator."))
        opt.step(), opt.zero_grad()
        for net in nets:
            for p in net.parameters():
                p.requires_grad_(True)
        my_losses.append(float(loss))
    tol = TOL[precision]
    for got, ref in zip(my_losses, ref_losses):
        assert abs(got - float(ref)) <= 3 * tol * abs(float(ref))
    for k in ("g_perceptual_loss", "g_adv_loss", "g_recon_loss"):
        assert abs(float(mine.logged[k]) - float(ora.logged[k])) <= 3 * tol * max(1e-3, abs(float(ora.logged[k]))), k
    check_grads(my_grads, ref_grads, 10 * tol, "perceptual step grads")


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
