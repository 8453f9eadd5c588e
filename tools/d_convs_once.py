"""The nine FLOP-dominant launches of the training step -- D layers 2-4 (batch 32, 256x256 input) x {fprop with fused
BatchNorm statistics, data gradient, weight gradient} -- three eager launches each; the third sits between
cudaProfilerStart / Stop, so `ncu --profile-from-start off --set full` (tools/gpu_r2.sh ncud) captures exactly nine warm
launches.  Inputs exceed L2 (264-520 MB per tensor)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from mpgan import ops  # noqa: E402

DEV, B = "cuda", 32
torch.manual_seed(0)
for name, cin, cout, k, s, xs in (("D2", 64, 128, 3, 1, 254), ("D3", 128, 256, 4, 2, 252), ("D4", 256, 256, 4, 2, 125)):
    spec = ops.ConvSpec(2, cin, cout, k, s, 0)
    ys = spec.y_of_x((xs, xs))[0]
    x = (torch.rand((B, xs, xs, cin), device=DEV) * 2 - 1).bfloat16()
    y = (torch.rand((B, ys, ys, cout), device=DEV) * 2 - 1).bfloat16()
    w = ((torch.rand((cout, k * k, cin), device=DEV) * 2 - 1) * 0.05).bfloat16()
    wt = w.permute(2, 1, 0).contiguous()
    dw = torch.zeros(cout, k * k, cin, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    for fn in (lambda: ops.conv_fprop(spec, x, w, None, out=y, stats=stats),
               lambda: ops.conv_bprop(spec, y, w, wt, None, xs=(xs, xs), out=x),
               lambda: ops.conv_wgrad(spec, x, y, dw)):
        fn(), fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print(name, "ok")
