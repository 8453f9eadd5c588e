"""Where the end-to-end step loses time against the device-resident replay: variants of the per-step host sequence around
the same captured graph (batch 32, 256x256, bf16), CUDA-event timed over 20 steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import mpgan  # noqa: E402
from bench import synthetic_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 256, 256, precision="bf16")
host = synthetic_batch(32, 2, 256, seed=1)
batch = {k: v.to(dev) for k, v in host.items()}
graph, static, logs = model.capture(batch)
pin = {k: v.pin_memory() for k, v in host.items()}
logs_host = torch.zeros(4).pin_memory()
stage = {k: torch.empty_like(static[k]) for k in static if k in pin}
fed = mpgan.HostFedStep(model, batch)
side = torch.cuda.Stream()


def v_replay():
    graph.replay()


def v_d2h():
    graph.replay()
    logs_host.copy_(logs, non_blocking=True)


def v_d2d():
    for k in stage:
        static[k].copy_(stage[k], non_blocking=True)
    graph.replay()
    logs_host.copy_(logs, non_blocking=True)


def v_inline():
    for k in pin:
        static[k].copy_(pin[k], non_blocking=True)
    graph.replay()
    logs_host.copy_(logs, non_blocking=True)


def v_fed():
    fed.step(pin)


def v_h2d_side_only():   # H2D into staging on a side stream, never consumed: pure interference
    with torch.cuda.stream(side):
        for k in pin:
            stage[k].copy_(pin[k], non_blocking=True)
    graph.replay()


ring = torch.zeros(4, 4, device=dev)
ev = [torch.cuda.Event() for _ in range(4)]
cnt = [0]


def v_d2h_side():      # losses -> ring slot by a KERNEL on the main stream, D2H of the slot on the side stream
    j = cnt[0] & 3
    graph.replay()
    torch.add(logs, 0.0, out=ring[j])
    ev[j].record()
    with torch.cuda.stream(side):
        side.wait_event(ev[j])
        logs_host.copy_(ring[j], non_blocking=True)
    cnt[0] += 1


def v_kernel_only():   # replay + one tiny kernel on the main stream
    graph.replay()
    torch.add(logs, 0.0, out=ring[0])


rdy = [torch.cuda.Event() for _ in range(2)]
fre = [torch.cuda.Event() for _ in range(2)]
stg2 = [{k: torch.empty_like(static[k]) for k in stage} for _ in range(2)]


def v_full_new():      # H2D on the side stream into double-buffered staging, kernel D2D into the inputs, losses via the ring
    i = cnt[0]
    j = i & 1
    with torch.cuda.stream(side):
        side.wait_event(fre[j])
        for k in stage:
            stg2[j][k].copy_(pin[k], non_blocking=True)
        rdy[j].record(side)
    cur = torch.cuda.current_stream()
    cur.wait_event(rdy[j])
    for k in stage:
        torch.add(stg2[j][k], 0.0, out=static[k])
    fre[j].record(cur)
    graph.replay()
    r = i & 3
    torch.add(logs, 0.0, out=ring[r])
    ev[r].record()
    with torch.cuda.stream(side):
        side.wait_event(ev[r])
        logs_host.copy_(ring[r], non_blocking=True)
    cnt[0] += 1


for name, fn in (("replay + tiny kernel", v_kernel_only), ("losses via ring kernel + side-stream D2H", v_d2h_side),
                 ("full new design", v_full_new), ("replay only", v_replay), ("+ D2H losses", v_d2h), ("+ D2D inputs", v_d2d), ("in-line H2D", v_inline),
                 ("HostFedStep", v_fed), ("replay + unconsumed H2D on a side stream", v_h2d_side_only), ("replay only", v_replay)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:45s} {e0.elapsed_time(e1) / 20:8.3f} ms/step", flush=True)
