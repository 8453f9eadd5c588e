"""The reference's LITERAL configuration: GAN_final.py's two-optimizer training step on 3-D volumes
-- G = 6 x UNet3D(16,32,64,128) + tanh, D = Conv3d 1->64->128 (k3) ->256->256 (k4 s2) + Linear(256*29^3, 1) on 128^3 --
bf16 on the rank-3 tcgen05 path (NDHWC, 5-D TMA boxes), CUDA-event timed.  One JSON line.
usage: python tools/bench_3d.py [batch] [steps] [size]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import mpgan  # noqa: E402
from bench import synthetic_batch, peaks  # noqa: E402

batch_n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
size = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, size, size, size, precision="bf16")
batch = {k: v.to(dev) for k, v in synthetic_batch(batch_n, 3, size, seed=1).items()}
use_graph = os.environ.get("MPGAN_3D_NOGRAPH", "0") != "1"
if use_graph:
    try:
        graph, static, logs = model.capture(batch)
        step = graph.replay
    except Exception as e:  # noqa: BLE001
        print(f"# graph capture failed: {e!r}", file=sys.stderr)
        use_graph = False
if not use_graph:
    logs = torch.zeros(4, device=dev)

    def step():
        model.fused_step(batch, logs)
step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# SURVEY.md section 8d (3-D reference-literal @128^3): 4 G_fwd + 8 D_fwd = 8.318e12 MAC per volume pair
G_MAC, D_MAC = 72.57e9 * (size / 128) ** 3, 1003.5e9 * (size / 128) ** 3
flop = 2 * (4 * G_MAC + 8 * D_MAC)
pk, _ = peaks()
v = batch_n / (ms * 1e-3)
lg = logs.tolist()
print(json.dumps({"metric": "gan_train_volume_pairs_per_sec", "value": v, "unit": "volume-pairs/s", "ms_per_step": ms,
                  "dtype": "bf16", "steps": steps, "cuda_graph": use_graph,
                  "config": {"workload": f"GAN_final.py literal 3-D step, batch {batch_n} of {size}^3 volumes, rank-3 tcgen05 path"},
                  "step_tflops": v * flop / 1e12,
                  "frac_of_sustained_compute_roofline": v * flop / 1e12 / float(pk["bf16_tflops_sustained"]),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "losses": {"g_adv": lg[0], "g_recon": lg[1], "d_loss": lg[2] + lg[3]}}))
