"""One eager generator inference forward (BASELINE configs[4]: batch 64 of 512x512, bf16, eval-mode BatchNorm folded),
after two warm-up forwards -- for an ncu launch list (`tools/gpu_r2.sh <tag> ncui`).
usage: python tools/infer_once.py [batch] [size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import mpgan  # noqa: E402
from mpgan import inference  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, size, size, precision="bf16").to(dev)
model.freeze()
vol = (torch.rand((batch, 1, size, size)) * 2 - 1).to(dev)
for _ in range(2):
    out = inference.infer_volume(model, vol, batch=batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()      # ncu --profile-from-start off: only the third forward is recorded
out = inference.infer_volume(model, vol, batch=batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(out.shape), bool(torch.isfinite(out).all()))
