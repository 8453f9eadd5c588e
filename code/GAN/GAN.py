"""Drop-in entry point the reference's README promises (``python code/GAN/GAN.py``; the reference tree only has
``GAN_final.py`` / ``test_runs/GAN.py``).  Same class names and constructor signatures as
/root/reference/code/GAN/GAN_final.py, backed by the B200 kernels in ``cross-modality-minipig-gan_b200``.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(_ROOT, "cross-modality-minipig-gan_b200"))

from mpgan import GAN, CasNetGenerator, Discriminator, PatchDiscriminator  # noqa: E402,F401

if __name__ == "__main__":
    import argparse

    import torch

    ap = argparse.ArgumentParser(description="synthetic-data training loop on the B200 kernels")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args()
    torch.manual_seed(0)
    model = GAN(1, a.size, a.size, precision=a.precision)
    g = torch.Generator().manual_seed(1)
    batch = {k: (torch.rand((a.batch, 1, a.size, a.size), generator=g) * 2 - 1).cuda() for k in ("t1w", "t2w")}
    for step in range(a.steps):
        logs = model.fused_step(batch).tolist()
        print(f"step {step}: g_adv={logs[0]:.4f} g_recon={logs[1]:.4f} d_loss={logs[2] + logs[3]:.4f}")
