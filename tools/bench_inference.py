"""BASELINE.json configs[4]: generator-only inference T1->T2 over a synthetic 256-slice volume at 512x512, batch 64,
bf16, plus the reference's post-processing (0/100-percentile rescale to 0..255, round, MAE) and the 1/99-percentile
pre-processing transform on a 128^3 volume.  CUDA-event timed; one JSON line.
usage: python tools/bench_inference.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import mpgan  # noqa: E402
from mpgan import inference, transforms as T  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 512, 512, precision="bf16").to(dev)
model.freeze()
g = torch.Generator().manual_seed(1)
vol = (torch.rand((256, 1, 512, 512), generator=g) * 2 - 1).to(dev)
truth = (torch.rand((256, 1, 512, 512), generator=g) * 2 - 1).to(dev)


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


gg = inference.GraphedGenerator(model, (64, 1, 512, 512), dev)
ms_eager, _ = timed(lambda: inference.infer_volume(model, vol, batch=64), reps)
ms_fwd, out = timed(lambda: inference.infer_volume(model, vol, batch=64, graphed=gg), reps)
ms_post, disp = timed(lambda: inference.to_display_range(out), reps)
ms_mae, _ = timed(lambda: inference.evaluate(out, truth), reps)
v128 = torch.rand((128, 128, 128), generator=g).to(dev) * 900
ms_pre, _ = timed(lambda: T.ScaleIntensityRangePercentiles(1.0, 99.0, -1.0, 1.0, clip=True)(v128), reps)
FLOP_PER_SLICE, BYTES_PER_SLICE = 2.897e10, 276.8e6   # SURVEY.md section 8d (cfg 5)
sl_s = 256 / (ms_fwd * 1e-3)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
print(json.dumps({
    "metric": "generator_inference_slices_per_sec", "value": sl_s, "unit": "slices/s", "dtype": "bf16",
    "config": {"workload": "G 6xUNet(16,32,64,128)+tanh eval forward, 256 slices of 512x512, batch 64"},
    "ms_per_volume": ms_fwd, "volumes_per_sec": 1e3 / ms_fwd, "ms_per_volume_eager": ms_eager, "cuda_graph": True,
    "algorithmic": {"tflops": sl_s * FLOP_PER_SLICE / 1e12, "gbs": sl_s * BYTES_PER_SLICE / 1e9,
                    "hbm_peak_gbs": peaks.get("hbm_gbs"), "frac_hbm": sl_s * BYTES_PER_SLICE / 1e9 / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None},
    "postprocess_ms_per_volume": ms_post, "postprocess_gbs": vol.numel() * 4 * 4 / (ms_post * 1e-3) / 1e9,
    "mae_mse_ms_per_volume": ms_mae, "mae_mse_gbs": vol.numel() * 8 / (ms_mae * 1e-3) / 1e9,
    "preprocess_128cube_ms": ms_pre, "preprocess_gbs": v128.numel() * 4 * 4 / (ms_pre * 1e-3) / 1e9,
}))
