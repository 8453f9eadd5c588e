"""Pin the oracle against the reference's own classes and write tests/golden/*.pt (TEST INFRASTRUCTURE ONLY).

Run in the authoring container (needs /root/reference):

    python -m oracle.make_golden            # fast pins  (~1-2 min CPU)
    python -m oracle.make_golden --full     # + GAN_final Discriminator / training_step at the literal 128^3

Every fixture stores the seeds that regenerate its inputs plus the outputs *of the reference's classes*
(executed verbatim through oracle/ref_shim.py).  ``tests/test_oracle.py`` re-derives the same quantities with
the oracle on whatever machine runs the tests and compares against these files.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

from . import ref_shim
from .gan import GANOracle, lightning_step, synthetic_batch, sample_patch_origins
from .nets import CasNetGenerator, Discriminator, PatchDiscriminator

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _same(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), f"{what}: max|diff|={float((a - b).abs().max())}"


def _grad_summary(named_params):
    return {n: torch.stack([p.grad.double().norm(), p.grad.double().sum()]).float() for n, p in named_params}


def pin_patch_discriminator(ref_pgan):
    """reference test_runs/GAN.py:136-198 (3-D, 16^3 patches) == oracle PatchDiscriminator(dims=3)."""
    torch.manual_seed(0)
    ref = ref_pgan.Discriminator(img_shape=(1, 16, 16, 16))
    ora = PatchDiscriminator((1, 16, 16, 16), dims=3, spatial=16)
    ora.load_state_dict(ref.state_dict())
    x = torch.rand((3, 1, 16, 16, 16), generator=torch.Generator().manual_seed(1)) * 2 - 1
    v_ref, a_ref = ref(x.clone())
    v_ora, a_ora = ora(x.clone())
    _same(v_ref, v_ora, "patchD validity")
    assert len(a_ref) == len(a_ora) == 16
    for k in a_ref:
        _same(a_ref[k], a_ora[k], f"patchD act {k}")
    v_ref.sum().backward()
    fix = {"seed_weights": 0, "seed_input": 1, "input_shape": list(x.shape), "validity": v_ref.detach(),
           "act_sums": torch.stack([a_ref[k].double().sum() for k in range(16)]).float(),
           "act_shapes": [list(a_ref[k].shape) for k in range(16)],
           "grad_summary": _grad_summary(ref.named_parameters()),
           "bn_running_mean_0": ref.model_conv[1].running_mean.clone()}
    torch.save(fix, os.path.join(GOLDEN, "ref_patch_discriminator_3d.pt"))
    print("pinned: test_runs/GAN.py Discriminator (3-D 16^3) == oracle PatchDiscriminator; fixture written")


def pin_generator(ref_final):
    """reference GAN_final.py:92-122 CasNetGenerator wrapper == oracle CasNetGenerator (both on the restated UNet)."""
    torch.manual_seed(0)
    ref = ref_final.CasNetGenerator(img_shape=(1, 16, 16, 16))
    ora = CasNetGenerator((1, 16, 16, 16), n_unet_blocks=6, dims=3)
    ora.load_state_dict(ref.state_dict())
    x = torch.rand((2, 1, 16, 16, 16), generator=torch.Generator().manual_seed(1)) * 2 - 1
    y_ref, y_ora = ref(x), ora(x)
    _same(y_ref, y_ora, "generator out")
    assert y_ref.shape == x.shape  # generator_test.py:84-88
    n_tensors = len(list(ref.parameters()))
    n_params = sum(p.numel() for p in ref.parameters())
    fix = {"seed_weights": 0, "seed_input": 1, "input_shape": list(x.shape), "out": y_ref.detach(),
           "n_param_tensors": n_tensors, "n_params": n_params,
           "state_keys": list(ref.state_dict().keys())}
    torch.save(fix, os.path.join(GOLDEN, "ref_generator_3d.pt"))
    print(f"pinned: GAN_final.py CasNetGenerator (3-D 16^3): {n_tensors} tensors / {n_params} params")


def pin_perceptual_step(ref_pgan):
    """reference test_runs/GAN.py:300-438 training_step (both optimizer indices) == oracle, 3-D, 24^3 volume."""
    B, S, NS = 2, 24, 4
    torch.manual_seed(0)
    batch = synthetic_batch(B, 3, S, seed=1)
    ref = ref_pgan.GAN(1, S, S, S, example_data={"t1w": batch["t1w"][0]})
    ref.patch_transform = ref_shim.Compose([ref_shim.RandSpatialCropSamplesd(
        keys=["t2", "t2_gt"], roi_size=(16, 16, 16), num_samples=NS)])
    ora = GANOracle("perceptual", dims=3, spatial=S, num_samples=NS)
    ora.load_state_dict(ref.state_dict())
    origins = sample_patch_origins(np.random.RandomState(2), B, NS, (S, S, S), 16)
    ref_shim.RandSpatialCropSamplesd.origins = origins
    out = {}
    for opt_idx in (0, 1):
        for net, on in ((ref.generator, opt_idx == 0), (ref.discriminator, opt_idx == 1)):
            for p in net.parameters():
                p.requires_grad_(on)
        for net, on in ((ora.generator, opt_idx == 0), (ora.discriminator, opt_idx == 1)):
            for p in net.parameters():
                p.requires_grad_(on)
        l_ref = ref.training_step(batch, 0, opt_idx)
        l_ora = ora.training_step(batch, 0, opt_idx, patch_origins=origins)
        _same(l_ref.detach().reshape(-1), l_ora.detach().reshape(-1), f"perceptual step loss opt{opt_idx}")
        for k in ref.logged:
            _same(ref.logged[k].reshape(-1), ora.logged[k].reshape(-1), f"logged {k}")
        l_ref.backward()
        net = ref.generator if opt_idx == 0 else ref.discriminator
        out[f"loss{opt_idx}"] = l_ref.detach().reshape(-1)
        out[f"grad_summary{opt_idx}"] = _grad_summary(net.named_parameters())
        for p in list(ref.parameters()) + list(ora.parameters()):
            p.grad = None
    out["logged"] = {k: v.reshape(-1) for k, v in ref.logged.items()}
    out.update({"B": B, "S": S, "num_samples": NS, "seed_weights": 0, "seed_input": 1, "seed_origins": 2,
                "origins": torch.from_numpy(origins)})
    torch.save(out, os.path.join(GOLDEN, "ref_perceptual_step_3d.pt"))
    print("pinned: test_runs/GAN.py GAN.training_step opt 0/1 (3-D 24^3, 4 patches) == oracle; fixture written")


def pin_final_discriminator_128(ref_final):
    """reference GAN_final.py:159-209 at its literal 128^3 input == oracle Discriminator(dims=3, spatial=128)."""
    torch.manual_seed(0)
    ref = ref_final.Discriminator(img_shape=(1, 128, 128, 128))
    ora = Discriminator((1, 128, 128, 128), dims=3, spatial=128)
    ora.load_state_dict(ref.state_dict())
    x = torch.rand((2, 1, 128, 128, 128), generator=torch.Generator().manual_seed(1)) * 2 - 1
    with torch.no_grad():
        t = time.time()
        y_ref = ref(x.clone())
        y_ora = ora(x.clone())
        print(f"  two 128^3 D forwards: {time.time() - t:.1f}s")
    _same(y_ref, y_ora, "final D validity")
    fix = {"seed_weights": 0, "seed_input": 1, "input_shape": list(x.shape), "validity": y_ref,
           "n_params": sum(p.numel() for p in ref.parameters()),
           "bn_running_mean_0": ref.model_conv[1].running_mean.clone(),
           "state_keys": list(ref.state_dict().keys())}
    torch.save(fix, os.path.join(GOLDEN, "ref_final_discriminator_128.pt"))
    print(f"pinned: GAN_final.py Discriminator @128^3 ({fix['n_params']} params) == oracle; fixture written")


def pin_final_step_128(ref_final):
    """reference GAN_final.py:250-296 training_step at 128^3, B=1 == oracle (losses + grad summaries)."""
    S = 128
    torch.manual_seed(0)
    batch = synthetic_batch(1, 3, S, seed=1)
    ref = ref_final.GAN(1, S, S, S)
    ora = GANOracle("final", dims=3, spatial=S)
    ora.load_state_dict(ref.state_dict())
    out = {}
    for opt_idx in (0, 1):
        for m in (ref, ora):
            for net, on in ((m.generator, opt_idx == 0), (m.discriminator, opt_idx == 1)):
                for p in net.parameters():
                    p.requires_grad_(on)
        t = time.time()
        l_ref = ref.training_step(batch, 0, opt_idx)
        l_ora = ora.training_step(batch, 0, opt_idx)
        _same(l_ref.detach().reshape(-1), l_ora.detach().reshape(-1), f"final step loss opt{opt_idx}")
        l_ref.backward()
        print(f"  opt{opt_idx}: loss={float(l_ref):.6f}  ({time.time() - t:.1f}s)")
        net = ref.generator if opt_idx == 0 else ref.discriminator
        out[f"loss{opt_idx}"] = l_ref.detach().reshape(-1)
        out[f"grad_summary{opt_idx}"] = _grad_summary(net.named_parameters())
        for p in list(ref.parameters()) + list(ora.parameters()):
            p.grad = None
    out["logged"] = {k: v.reshape(-1) for k, v in ref.logged.items()}
    out.update({"S": S, "seed_weights": 0, "seed_input": 1})
    torch.save(out, os.path.join(GOLDEN, "ref_final_step_128.pt"))
    print("pinned: GAN_final.py GAN.training_step opt 0/1 @128^3 == oracle; fixture written")


def write_2d_twin_goldens():
    """Oracle-generated vectors for the 2-D twin (BASELINE cfg 1 shape family), small enough to commit.
    These are NOT reference outputs (the reference has no 2-D classes); they freeze the oracle's own numbers
    so a drift of the oracle (or of torch CPU kernels) is noticed, and give the GPU tests fixed targets."""
    B, S = 2, 64
    torch.manual_seed(0)
    m = GANOracle("final", dims=2, spatial=S)
    batch = synthetic_batch(B, 2, S, seed=1)
    opts, _ = m.configure_optimizers()
    grads = {}
    losses = lightning_step(m, opts, batch, 0, keep_grads=grads)
    with torch.no_grad():
        m.eval()
        y_eval = m(batch["t1w"])
    fix = {"B": B, "S": S, "seed_weights": 0, "seed_input": 1,
           "g_loss": losses[0].reshape(-1), "d_loss": losses[1].reshape(-1),
           "logged": {k: v.reshape(-1) for k, v in m.logged.items()},
           "grad_norms": {k: v.double().norm().float() for k, v in grads.items()},
           "eval_out_after_step": y_eval}
    torch.save(fix, os.path.join(GOLDEN, "oracle_final_step_2d_64.pt"))
    print(f"wrote 2-D twin golden: g_loss={float(losses[0]):.6f} d_loss={float(losses[1]):.6f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also pin the literal 128^3 GAN_final classes")
    args = ap.parse_args()
    if not ref_shim.available():
        sys.exit("reference not present at /root/reference; fixtures can only be generated in the authoring container")
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref_final = ref_shim.load_reference_module("code/GAN/GAN_final.py", "ref_GAN_final")
    ref_pgan = ref_shim.load_reference_module("test_runs/GAN.py", "ref_GAN_perceptual")
    pin_patch_discriminator(ref_pgan)
    pin_generator(ref_final)
    pin_perceptual_step(ref_pgan)
    write_2d_twin_goldens()
    if args.full:
        pin_final_discriminator_128(ref_final)
        pin_final_step_128(ref_final)


if __name__ == "__main__":
    main()
