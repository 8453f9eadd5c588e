"""Generator phases of one training step run eagerly once (after a warm-up step), each between cudaProfilerStart / Stop
markers -- for an ncu launch list with warm caches:
    ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --profile-from-start off --csv ...
usage: python tools/g_phases_once.py [fwd|bwd|both]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import mpgan  # noqa: E402
from bench import synthetic_batch  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 256, 256, precision="bf16")
batch = {k: v.to(dev) for k, v in synthetic_batch(32, 2, 256, seed=1).items()}
G = model.generator
logs = torch.zeros(4, device=dev)
model.fused_step(batch, logs)
torch.cuda.synchronize()
t1 = batch["t1w"]
for rep in range(2):
    prof = rep == 1
    if prof and what in ("fwd", "both"):
        torch.cuda.profiler.start()
    gen, gplan = G.run_forward(t1, save=True, need_wgrad=True)
    torch.cuda.synchronize()
    if prof and what == "fwd":
        torch.cuda.profiler.stop()
    if prof and what == "bwd":
        torch.cuda.profiler.start()
    G.run_backward(gplan, torch.ones_like(gen) * 1e-3, need_dx=False)
    torch.cuda.synchronize()
    if prof and what in ("bwd", "both"):
        torch.cuda.profiler.stop()
    G.runtime.zero_grad()
print("ok")
