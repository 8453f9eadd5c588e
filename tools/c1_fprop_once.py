"""One-channel stride-2 forward convolution (1 -> 16, the UNet's first layer) at the inference size, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch
from mpgan import ops
DEV = "cuda"
B, xs, cout = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 512, 16
spec = ops.ConvSpec(2, 1, cout, 3, 2, 1)
x = (torch.rand((B, xs, xs, 1), device=DEV) * 2 - 1).bfloat16()
y = torch.empty((B, xs // 2, xs // 2, cout), device=DEV, dtype=torch.bfloat16)
w = ((torch.rand((cout, 9, 1), device=DEV) * 2 - 1) * 0.2).bfloat16()
fn = lambda: ops.conv_fprop(spec, x, w, None, out=y, stats=None)
fn(); fn(); torch.cuda.synchronize()
torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
