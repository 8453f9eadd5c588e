import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch
from mpgan import ops
DEV, B = "cuda", 32
torch.manual_seed(0)
cin, cout, k, s, xs = 64, 128, 3, 1, 254
spec = ops.ConvSpec(2, cin, cout, k, s, 0)
ys = spec.y_of_x((xs, xs))[0]
x = (torch.rand((B, xs, xs, cin), device=DEV) * 2 - 1).bfloat16()
y = (torch.rand((B, ys, ys, cout), device=DEV) * 2 - 1).bfloat16()
w = ((torch.rand((cout, k * k, cin), device=DEV) * 2 - 1) * 0.05).bfloat16()
stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
mode = sys.argv[1] if len(sys.argv) > 1 else "stats"
fn = (lambda: ops.conv_fprop(spec, x, w, None, out=y, stats=stats)) if mode == "stats" else (lambda: ops.conv_fprop(spec, x, w, None, out=y, stats=None))
fn(); fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
print(mode, "us", e0.elapsed_time(e1) * 1e3)
torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
