"""Generator-only inference over a volume and the post-processing the reference applies to it.

Mirrors /root/reference/code/GAN/inferrence.py:147-204 (and minipig_inference.py:101-128): ``model.eval()``,
``model.freeze()``, ``model.generator.forward(t1)`` under ``torch.no_grad()``, then
``ScaleIntensityRangePercentilesd(lower=0, upper=100, b_min=0, b_max=255, clip=True)`` + ``np.round`` on the generated
and ground-truth volumes and ``MeanAbsoluteError`` between them.  The reference feeds one volume per call; here the
slices (2-D) or sub-volumes (3-D) are batched, the rescale / round / metric kernels run on the device and nothing is
copied to the host except the two percentile scalars.
"""
import torch

from .transforms import ScaleIntensityRangePercentiles, error_sums


@torch.no_grad()
def infer_volume(model, t1, batch=64):
    """``t1``: (S, 1, H, W) slices or (S, 1, D, H, W) sub-volumes, CUDA fp32 in [-1, 1].  Returns the generated T2
    tensor of the same shape.  ``model`` is a ``GAN`` or a ``CasNetGenerator``; BatchNorm uses running statistics."""
    gen = getattr(model, "generator", model)
    if not t1.is_cuda:
        raise RuntimeError("mpgan inference needs CUDA tensors on a B200: there is no CPU fallback")
    was_training = gen.training
    gen.eval()
    try:
        outs = []
        for s in range(0, t1.shape[0], batch):
            y, _ = gen.run_forward(t1[s:s + batch].contiguous(), save=False, need_wgrad=False)
            outs.append(y)
        return torch.cat(outs, dim=0) if len(outs) != 1 else outs[0]
    finally:
        gen.train(was_training)


def to_display_range(vol, out_dtype=torch.float32):
    """inferrence.py:152-161 / 190-199: 0/100-percentile rescale to [0, 255], clip, round half to even."""
    return ScaleIntensityRangePercentiles(0, 100, 0, 255, clip=True)(vol, round_half_even=True, out_dtype=out_dtype)


def evaluate(generated, truth):
    """{"mae", "mse"} between two volumes (torchmetrics MeanAbsoluteError / MeanSquaredError, one pass)."""
    s = error_sums(generated, truth).tolist()
    n = max(generated.numel(), 1)
    return {"mae": s[0] / n, "mse": s[1] / n}
