"""TEST INFRASTRUCTURE (oracle): numpy restatement of the intensity transforms and volume metrics either side of the
generator.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows
* monai==0.4.0 ``monai.transforms.ScaleIntensityRangePercentiles`` / ``ScaleIntensityRange`` (external, not vendored
  by the reference; call sites /root/reference/code/GAN/GAN_final.py:386-394 -- lower=1, upper=99, b_min=-1, b_max=1,
  clip=True -- and /root/reference/code/GAN/inferrence.py:152-160,190-198 -- lower=0, upper=100, b_min=0, b_max=255,
  clip=True, followed by ``np.round`` at :161,:199):

      a_min = np.percentile(img, lower); a_max = np.percentile(img, upper)
      b_min, b_max = self.b_min, self.b_max
      if relative: b_min = (self.b_max - self.b_min) * (lower / 100) + self.b_min   (b_max likewise with upper)
      img = ScaleIntensityRange(a_min, a_max, b_min, b_max, clip=False)(img)
      if clip: img = np.clip(img, self.b_min, self.b_max)
  ScaleIntensityRange:  a_max == a_min -> img - a_min + b_min;  else ((img - a_min) / (a_max - a_min)) * (b_max - b_min) + b_min

* torchmetrics ``MeanAbsoluteError`` / ``MeanSquaredError`` (inferrence.py:170-176, metrics.py:213-218).

Parity unpinned for the MONAI part (MONAI is not installable here; the class is restated from its published source).
Numeric convention, stated because numpy changed it: the images are float32 and, under the numpy 1.19 value-based
casting MONAI 0.4.0 ran on, a float32 array combined with the python / float64 percentile scalars stays float32 --
so every arithmetic step below is rounded to float32.  The percentile is numpy's default linear interpolation between
the two neighbouring order statistics, evaluated in float64 and then rounded to float32.
"""
import numpy as np


def percentile_ranks(n, q):
    """(lower rank, upper rank, fraction) of np.percentile(..., q) with linear interpolation."""
    pos = (q / 100.0) * (n - 1)
    lo = int(np.floor(pos))
    frac = pos - lo
    hi = lo if frac == 0 else min(lo + 1, n - 1)   # exact percentiles (0, 100, ...) need one order statistic only
    return lo, hi, frac


def percentile_f32(img, q):
    flat = np.sort(np.asarray(img, dtype=np.float32).ravel())
    lo, hi, frac = percentile_ranks(flat.size, q)
    v = float(flat[lo]) + (float(flat[hi]) - float(flat[lo])) * frac
    return np.float32(v)


def scale_intensity_range(img, a_min, a_max, b_min, b_max, clip=False):
    img = np.asarray(img, dtype=np.float32)
    a_min, a_max, b_min, b_max = (np.float32(v) for v in (a_min, a_max, b_min, b_max))
    if a_max - a_min == 0:
        out = img - a_min + b_min
    else:
        out = (img - a_min) / (a_max - a_min)
        out = out * (b_max - b_min) + b_min
    if clip:
        out = np.clip(out, b_min, b_max)
    return out.astype(np.float32)


def scale_intensity_range_percentiles(img, lower, upper, b_min, b_max, clip=False, relative=False):
    a_min, a_max = percentile_f32(img, lower), percentile_f32(img, upper)
    bl, bu = b_min, b_max
    if relative:
        bl = ((b_max - b_min) * (lower / 100.0)) + b_min
        bu = ((b_max - b_min) * (upper / 100.0)) + b_min
    out = scale_intensity_range(img, a_min, a_max, bl, bu, clip=False)
    if clip:
        out = np.clip(out, np.float32(b_min), np.float32(b_max))
    return out.astype(np.float32)


def to_display_range(img):
    """inferrence.py:152-161: percentiles 0/100 -> [0, 255], clip, np.round."""
    return np.round(scale_intensity_range_percentiles(img, 0, 100, 0, 255, clip=True)).astype(np.float32)


def mean_absolute_error(a, b):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return float(np.abs(a - b).astype(np.float64).sum() / a.size)


def mean_squared_error(a, b):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    d = (a - b).astype(np.float64)
    return float((d * d).sum() / a.size)
