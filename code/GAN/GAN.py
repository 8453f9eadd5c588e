"""Drop-in entry point the reference's README promises (``python code/GAN/GAN.py``; the reference tree only has
``GAN_final.py`` / ``test_runs/GAN.py``).  Same class names and constructor signatures as
/root/reference/code/GAN/GAN_final.py, backed by the B200 kernels in ``cross-modality-minipig-gan_b200``.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(_ROOT, "cross-modality-minipig-gan_b200"))

from mpgan import GAN, CasNetGenerator, Discriminator, PatchDiscriminator  # noqa: E402,F401
from mpgan import transforms  # noqa: E402,F401  (ScaleIntensityRangePercentilesd: the intensity transform either side of the path)


class HumanBrainDataModule:
    """/root/reference/code/GAN/GAN_final.py:321-437.  OUT OF SCOPE: it reads the authors' private
    PREDICT-HD volumes through ITK / MONAI ``CacheDataset`` and cannot run anywhere else.  Kept under its reference
    name so that a caller gets an explanation instead of an ImportError; feed ``GAN`` dict batches
    ``{"t1w": (B,1,...), "t2w": (B,1,...)}`` of fp32 CUDA tensors scaled to [-1, 1] from any loader
    (``mpgan.transforms.ScaleIntensityRangePercentilesd`` is the on-device 1/99-percentile rescale of :386-394)."""

    def __init__(self, spatial_size=(128, 128, 128)):
        raise NotImplementedError(
            "HumanBrainDataModule is the reference's ITK/MONAI loader for a private dataset and is outside the B200 hot "
            "path; build batches {'t1w','t2w'} with your own DataLoader (see INTEGRATION.md) -- the GAN / "
            "CasNetGenerator / Discriminator classes exported here are the drop-in part")

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
