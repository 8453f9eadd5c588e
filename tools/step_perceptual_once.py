"""One eager perceptual fused step (BASELINE configs[2] shape) for launch-list profiling: warm-up steps, then one more."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mpgan  # noqa: E402
from bench import synthetic_batch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 256, 256, variant="perceptual", precision="bf16")
batch = {k: v.to(dev) for k, v in synthetic_batch(32, 2, 256, seed=1).items()}
origins = np.random.RandomState(2).randint(0, 256 - 16 + 1, size=(32, 128, 2))
batch["origins"] = torch.as_tensor(origins.reshape(-1, 2), dtype=torch.int32, device=dev)
for _ in range(n):
    logs = model.fused_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()      # ncu --profile-from-start off: one more step is recorded
logs = model.fused_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(logs.tolist())
