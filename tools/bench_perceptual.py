"""BASELINE.json configs[2]: the test_runs/GAN.py training step (G = 4 x UNet(32,64,128,256), 128 random 16x16 patches
per image, patch discriminator with its 16 activations, adversarial + L1(patches) + discriminator-feature "perceptual"
loss) at batch 32 of 256x256 slices, bf16, driven through the reference's own two-optimizer protocol
(``GAN.fit_batch``: training_step -> backward -> step -> zero_grad per optimizer).  CUDA-event timed; one JSON line.
usage: python tools/bench_perceptual.py [steps] [graph|protocol]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mpgan  # noqa: E402
from bench import synthetic_batch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 256, 256, variant="perceptual", precision="bf16").to(dev)
batch = {k: v.to(dev) for k, v in synthetic_batch(32, 2, 256, seed=1).items()}
origins = np.random.RandomState(2).randint(0, 256 - 16 + 1, size=(32, 128, 2))
mode = sys.argv[2] if len(sys.argv) > 2 else "graph"
o_dev = torch.as_tensor(origins.reshape(-1, 2), dtype=torch.int32, device=dev)
if mode == "protocol":      # the reference's own two-optimizer protocol through autograd (eager)
    def step():
        return model.fit_batch(batch, patch_origins=origins)
else:                       # the same arithmetic as one static kernel sequence, replayed from a CUDA graph
    graph, static, logs = model.capture(dict(batch, origins=o_dev))

    def step():
        graph.replay()
        return logs
for _ in range(3):
    losses = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    losses = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if mode != "protocol":
    losses = [losses[0] + losses[1] + losses[2], losses[3] + losses[4]]
FLOP_PER_PAIR = 3.480e11   # SURVEY.md section 8d (cfg 3)
print(json.dumps({"metric": "gan_train_image_pairs_per_sec", "value": 32 / (ms * 1e-3), "unit": "image-pairs/s",
                  "ms_per_step": ms, "dtype": "bf16", "steps": steps,
                  "config": {"workload": "test_runs/GAN.py step: G 4xUNet(32,64,128,256), 128 patches of 16x16 per image, patch-D "
                                         "with activations, adv + L1 + perceptual; batch 32 of 256x256; eager two-optimizer protocol"},
                  "step_tflops": 32 / (ms * 1e-3) * FLOP_PER_PAIR / 1e12,
                  "losses": {"g_loss": float(losses[0]), "d_loss": float(losses[1])}}))
