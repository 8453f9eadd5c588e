"""Where one training step spends its time: every phase of GAN.fused_step captured as its own CUDA graph and timed
over replays with CUDA events (batch 32, 256x256, bf16).  usage: python tools/phase_times.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import mpgan  # noqa: E402
from mpgan import ops, _lib  # noqa: E402
from bench import synthetic_batch  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mpgan.GAN(1, 256, 256, precision="bf16")
batch = {k: v.to(dev) for k, v in synthetic_batch(32, 2, 256, seed=1).items()}
G, D, hp = model.generator, model.discriminator, model.hparams
t1, t2 = batch["t1w"], batch["t2w"]
logs = torch.zeros(4, device=dev)
model.fused_step(batch, logs)  # warm: lazy module load, func attributes
torch.cuda.synchronize()
ones, soft, zeros = model._consts(32, dev)
state = {}


def ph_g_fwd_save():
    state["gen"], state["gplan"] = G.run_forward(t1, save=True, need_wgrad=True)


def ph_d_fwd_nowg():
    state["p"], state["dplan"] = D.run_forward(state["gen"], save=True, need_wgrad=False)


def ph_d_bwd_dx():
    p = state["p"]
    dprob = ops.bce_bwd(p, ones, 1.0, None, torch.empty_like(p))
    state["dgen"] = D.run_backward(state["dplan"], dprob, need_dx=True)


def ph_g_bwd():
    G.run_backward(state["gplan"], state["dgen"], need_dx=False)


def ph_g_adam():
    G.runtime.adam_step(hp.g_lr, hp.b1, hp.b2)
    G.runtime.zero_grad()


def ph_d_fwd_wg():
    state["p2"], state["dplan2"] = D.run_forward(t2, save=True, need_wgrad=True)


def ph_g_fwd_nosave():
    state["gen2"], _ = G.run_forward(t1, save=False, need_wgrad=False)


def ph_d_bwd_wg():
    p = state["p2"]
    D.run_backward(state["dplan2"], ops.bce_bwd(p, zeros, 0.5, None, torch.empty_like(p)), need_dx=False)


def ph_d_adam():
    D.runtime.adam_step(hp.d_lr, hp.b1, hp.b2)
    D.runtime.zero_grad()


PHASES = [("G fwd (tape)            x1", ph_g_fwd_save, 1), ("D fwd (dgrad-only plan) x1", ph_d_fwd_nowg, 1),
          ("D bwd dx only           x1", ph_d_bwd_dx, 1), ("G bwd                   x1", ph_g_bwd, 1),
          ("G adam+zero             x1", ph_g_adam, 1), ("D fwd (tape)            x2", ph_d_fwd_wg, 2),
          ("G fwd (no tape)         x1", ph_g_fwd_nosave, 1), ("D bwd wgrad only        x2", ph_d_bwd_wg, 2),
          ("D adam+zero             x1", ph_d_adam, 1)]

total = 0.0
keep = []
for name, fn, mult in PHASES:
    c0 = _lib.ABI_CALLS
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    calls = _lib.ABI_CALLS - c0
    keep.append(g)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    total += ms * mult
    print(f"{name}: {ms:8.3f} ms  ({calls} ABI calls)  -> {ms * mult:8.3f} ms/step", flush=True)
print(f"sum over the step: {total:.3f} ms")
