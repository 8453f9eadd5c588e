"""G backward: the data-gradient chain alone vs with the weight gradients (how much of the pass is the side stream)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch, mpgan
from bench import synthetic_batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 256, 256, precision="bf16")
batch = {k: v.to(dev) for k, v in synthetic_batch(32, 2, 256, seed=1).items()}
G = model.generator
model.fused_step(batch); torch.cuda.synchronize()
dgen = torch.rand(32, 1, 256, 256, device=dev) - 0.5
for wg in (True, False):
    st = {}
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1):
        st["gen"], st["plan"] = G.run_forward(batch["t1w"], save=True, need_wgrad=wg)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        G.run_backward(st["plan"], dgen, need_dx=False)
    g1.replay(); g2.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g2.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"G backward, weight gradients {'on ' if wg else 'off'}: {e0.elapsed_time(e1) / 10:.3f} ms")
