"""Device-side intensity transforms and volume metrics.

Mirrors the interface of the MONAI 0.4.0 transforms the reference applies either side of the generator:
``ScaleIntensityRangePercentiles(lower, upper, b_min, b_max, clip=False, relative=False)`` and its dictionary form
``ScaleIntensityRangePercentilesd(keys, ...)`` (/root/reference/code/GAN/GAN_final.py:386-394,
/root/reference/code/GAN/inferrence.py:152-161), plus torchmetrics' ``MeanAbsoluteError`` / ``MeanSquaredError``
(inferrence.py:170-176, metrics.py:213-218).  Inputs are CUDA fp32 tensors; everything runs in libmpgan_sm100
(exact radix-select order statistics, fp32 rescale in the reference's operation order) -- there is no CPU path.
"""
import math

import torch

from . import _lib
from .ops import check, ptr, _stream


def _require(x):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError("mpgan.transforms need CUDA tensors on a B200: there is no CPU fallback")
    return x.detach().contiguous().float()


def percentile_ranks(n, q):
    """np.percentile's linear interpolation: position q/100*(n-1) between two neighbouring order statistics."""
    pos = (q / 100.0) * (n - 1)
    lo = int(math.floor(pos))
    frac = pos - lo
    hi = lo if frac == 0 else min(lo + 1, n - 1)   # exact percentiles (0, 100, ...) need one order statistic only
    return lo, hi, frac


def order_statistics(x, ranks):
    """The ``ranks``-th smallest values (0-based, exact) of a CUDA fp32 tensor; returns a python list of floats."""
    x = _require(x)
    if x.numel() == 0:
        raise RuntimeError("order_statistics of an empty tensor")
    lib = _lib.require_device()
    out = []
    for i in range(0, len(ranks), 4):
        grp = [int(r) for r in ranks[i:i + 4]]
        if any(r < 0 or r >= x.numel() for r in grp):
            raise RuntimeError(f"rank out of range for {x.numel()} elements: {grp}")
        r_dev = torch.tensor(grp, dtype=torch.int64, device=x.device)
        o_dev = torch.empty(len(grp), dtype=torch.float32, device=x.device)
        nbytes = int(lib.mpgan_order_stats_workspace(len(grp)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        if all(r in (0, x.numel() - 1) for r in grp):   # 0 / 100 percentiles: one min/max pass
            check(lib.mpgan_minmax(ptr(x), x.numel(), ptr(r_dev), len(grp), ptr(o_dev), ptr(ws), nbytes, _stream()),
                  "mpgan_minmax")
        else:
            check(lib.mpgan_order_stats(ptr(x), x.numel(), ptr(r_dev), len(grp), ptr(o_dev), ptr(ws), nbytes, _stream()),
                  "mpgan_order_stats")
        out += o_dev.tolist()
    return out


def percentiles(x, qs):
    """np.percentile(x, q) (linear interpolation, evaluated in float64, rounded to float32) for each q."""
    n = x.numel()
    spec = [percentile_ranks(n, q) for q in qs]
    ranks = sorted({r for lo, hi, _ in spec for r in (lo, hi)})
    vals = dict(zip(ranks, order_statistics(x, ranks)))
    res = []
    for lo, hi, frac in spec:
        v = vals[lo] + (vals[hi] - vals[lo]) * frac
        res.append(float(torch.tensor(v, dtype=torch.float64).float()))
    return res


def rescale_intensity(x, a_min, a_max, b_min, b_max, clip=None, round_half_even=False, out_dtype=torch.float32):
    """MONAI ScaleIntensityRange arithmetic in fp32; ``clip`` = (lo, hi) or None."""
    x = _require(x)
    lib = _lib.require_device()
    if out_dtype not in (torch.float32, torch.float16):
        raise RuntimeError("rescale_intensity writes float32 or float16")
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    lo, hi = clip if clip is not None else (0.0, 0.0)
    check(lib.mpgan_rescale_intensity(ptr(x), x.numel(), float(a_min), float(a_max), float(b_min), float(b_max),
                                      0 if clip is None else 1, float(lo), float(hi), 1 if round_half_even else 0,
                                      0 if out_dtype == torch.float32 else 2, ptr(y), _stream()),
          "mpgan_rescale_intensity")
    return y


class ScaleIntensityRangePercentiles:
    """monai.transforms.ScaleIntensityRangePercentiles (0.4.0) on a CUDA tensor."""

    def __init__(self, lower, upper, b_min, b_max, clip=False, relative=False):
        if not (0.0 <= lower <= 100.0 and 0.0 <= upper <= 100.0):
            raise ValueError("Percentiles must be in the range [0, 100]")
        self.lower, self.upper, self.b_min, self.b_max, self.clip, self.relative = lower, upper, b_min, b_max, clip, relative

    def __call__(self, img, round_half_even=False, out_dtype=torch.float32):
        a_min, a_max = percentiles(img, [self.lower, self.upper])
        b_min, b_max = self.b_min, self.b_max
        if self.relative:
            b_min = ((self.b_max - self.b_min) * (self.lower / 100.0)) + self.b_min
            b_max = ((self.b_max - self.b_min) * (self.upper / 100.0)) + self.b_min
        return rescale_intensity(img, a_min, a_max, b_min, b_max, clip=(self.b_min, self.b_max) if self.clip else None,
                                 round_half_even=round_half_even, out_dtype=out_dtype)


class ScaleIntensityRangePercentilesd:
    """Dictionary form (keys are transformed independently, like MONAI's MapTransform)."""

    def __init__(self, keys, lower, upper, b_min, b_max, clip=False, relative=False):
        self.keys = [keys] if isinstance(keys, str) else list(keys)
        self.scaler = ScaleIntensityRangePercentiles(lower, upper, b_min, b_max, clip, relative)

    def __call__(self, data):
        d = dict(data)
        for k in self.keys:
            d[k] = self.scaler(d[k])
        return d


def error_sums(a, b):
    """(sum |a-b|, sum (a-b)^2) as a device float64 pair."""
    a, b = _require(a), _require(b)
    if a.shape != b.shape:
        raise RuntimeError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    lib = _lib.require_device()
    out = torch.zeros(2, dtype=torch.float64, device=a.device)
    check(lib.mpgan_err_sums(ptr(a), ptr(b), a.numel(), ptr(out), _stream()), "mpgan_err_sums")
    return out


def mean_absolute_error(a, b):
    return float(error_sums(a, b)[0]) / max(a.numel(), 1)


def mean_squared_error(a, b):
    return float(error_sums(a, b)[1]) / max(a.numel(), 1)
