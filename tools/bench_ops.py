"""Per-op CUDA-event timings of the layer shapes of the bench workload (batch 32, 256x256), through the C ABI.

    python tools/bench_ops.py [filter-substring ...]

Small (generator) tensors are timed back to back on the same buffers, i.e. L2-resident as they are inside the real
step (their producer has just written them); the discriminator tensors exceed the 126 MB L2 on their own.
Prints one line per op: microseconds, algorithmic GB/s (inputs read once + outputs written once) and TFLOP/s.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from mpgan import ops  # noqa: E402
from mpgan._lib import ACT_LEAKY, ACT_PRELU  # noqa: E402

DEV = "cuda"
B = int(os.environ.get("OPS_BATCH", "32"))      # OPS_BATCH=64 OPS_SCALE=2: the inference shapes of BASELINE configs[4]
SCALE = int(os.environ.get("OPS_SCALE", "1"))


def timeit(fn, reps=20, warm=3):
    """`reps` back-to-back launches replayed from a CUDA graph (no host launch overhead in the number)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def rnd(*shape, dtype=torch.bfloat16):
    return (torch.rand(shape, device=DEV) * 2 - 1).to(dtype)


def report(name, us, nbytes, flops=0.0):
    print(f"{name:58s} {us:9.1f} us  {nbytes / us / 1e3:8.1f} GB/s  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)


def conv_case(name, cin, cout, k, s, p, xs, transposed=False, dirs=("f", "b", "w"), rank=2, batch=None):
    """Underlying conv X(cin) -> Y(cout); for a ConvTranspose pass the underlying conv's cin/cout.  rank 3: the
    reference's literal volumes (NDHWC, rank-3 tcgen05 path)."""
    n = batch or B
    if rank == 2 and name.startswith("G"):
        xs *= SCALE
    spec = ops.ConvSpec(rank, cin, cout, k, s, p, transposed, s - 1 if transposed else 0)
    ys = spec.y_of_x((xs,) * rank)[0]
    x = rnd(n, *([xs] * rank), cin)
    y = rnd(n, *([ys] * rank), cout)
    w = rnd(cout, k ** rank, cin) * 0.1
    wt = w.permute(2, 1, 0).contiguous()
    dw = torch.zeros(cout, k ** rank, cin, device=DEV)
    stats = None if os.environ.get("NOSTATS") else torch.zeros(2 * max(cin, cout), dtype=torch.float64, device=DEV)
    flops = 2.0 * n * ys ** rank * cout * k ** rank * cin
    nb = (x.numel() + y.numel()) * 2
    # BatchNorm statistics are fused only where the training step fuses them: in the layer's FORWARD direction (fprop of a
    # convolution, bprop of a ConvTranspose); the data gradients run without (round 1's table timed them with)
    sf, sb = (None, stats) if transposed else (stats, None)
    if "f" in dirs:
        report(f"{name} fprop {cin}->{cout} k{k}s{s} @{xs}", timeit(lambda: ops.conv_fprop(spec, x, w, None, out=y, stats=sf)), nb, flops)
    if "b" in dirs:
        report(f"{name} bprop {cout}->{cin} k{k}s{s} @{ys}", timeit(lambda: ops.conv_bprop(spec, y, w, wt, None, xs=(xs,) * rank, out=x, stats=sb)), nb, flops)
    if "w" in dirs:
        report(f"{name} wgrad {cin}x{cout} k{k}s{s} @{xs}", timeit(lambda: ops.conv_wgrad(spec, x, y, dw)), nb, flops)


def bn_case(name, c, hw, act):
    x, res, y, dy = rnd(B, hw, hw, c), rnd(B, hw, hw, c), rnd(B, hw, hw, c), rnd(B, hw, hw, c)
    bn = torch.nn.BatchNorm2d(c).to(DEV)
    alpha = torch.full((1,), 0.25, device=DEV)
    stats = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    ops.bn_stats(x, stats)
    saved = torch.empty(4, c, device=DEV)
    nb = x.numel() * 2
    a = alpha if act == ACT_PRELU else None
    report(f"{name} bn_stats c{c} @{hw}", timeit(lambda: ops.bn_stats(x, stats)), nb)
    stats.zero_(), ops.bn_stats(x, stats)
    report(f"{name} bn_train_apply c{c} @{hw}", timeit(lambda: ops.bn_train_apply(x, stats, bn, saved, act, a, 0.2, None, y)), 2 * nb)
    sums = torch.zeros(2 * c + 1, dtype=torch.float64, device=DEV)
    dg, db, da, dbias = (torch.zeros(c, device=DEV) for _ in range(4))
    lib = ops._lib.require_device()
    from mpgan.ops import check, dt, ld, pixels, ptr, _stream

    def red():
        check(lib.mpgan_bn_act_bwd_reduce(dt(x), ptr(dy), ld(dy), ptr(x), ld(x), pixels(x), c, ptr(saved[0]), ptr(saved[1]),
                                          ptr(saved[2]), ptr(saved[3]), act, ptr(a), 0.2, ptr(sums), _stream()), "r")

    def app():
        check(lib.mpgan_bn_act_bwd_apply(dt(x), ptr(dy), ld(dy), ptr(x), ld(x), pixels(x), c, ptr(saved[0]), ptr(saved[1]),
                                         ptr(saved[2]), ptr(saved[3]), act, ptr(a), 0.2, ptr(sums), ptr(dg), ptr(db),
                                         ptr(da[:1]), ptr(dbias), ptr(y), ld(y), _stream()), "a")
    report(f"{name} bn_bwd_reduce c{c} @{hw}", timeit(red), 2 * nb)
    report(f"{name} bn_bwd_apply c{c} @{hw}", timeit(app), 3 * nb)


CASES = {
    # one-channel edge layers
    "c1_d1": lambda: conv_case("D1", 1, 64, 3, 1, 0, 256),
    "c1_g1": lambda: conv_case("G first", 1, 16, 3, 2, 1, 256),
    "c1_gT": lambda: conv_case("G ConvT32->1", 1, 32, 3, 2, 1, 256, transposed=True),
    "c1_11": lambda: conv_case("G 1->1", 1, 1, 3, 1, 1, 256),
    # generator tensor-core layers
    "g_16_16": lambda: conv_case("G", 16, 16, 3, 1, 1, 128),
    "g_16_32": lambda: conv_case("G", 16, 32, 3, 2, 1, 128),
    "g_32_32": lambda: conv_case("G", 32, 32, 3, 1, 1, 64),
    "g_32_64": lambda: conv_case("G", 32, 64, 3, 2, 1, 64),
    "g_64_64": lambda: conv_case("G", 64, 64, 3, 1, 1, 32),
    "g_64_128": lambda: conv_case("G", 64, 128, 3, 1, 1, 32),
    "g_128_128": lambda: conv_case("G", 128, 128, 3, 1, 1, 32),
    "g_64_128_1x1": lambda: conv_case("G res", 64, 128, 1, 1, 0, 32),
    "gT_32_192": lambda: conv_case("G ConvT192->32", 32, 192, 3, 2, 1, 64, transposed=True),
    "gT_16_64": lambda: conv_case("G ConvT64->16", 16, 64, 3, 2, 1, 128, transposed=True),
    # discriminator
    "d2": lambda: conv_case("D2", 64, 128, 3, 1, 0, 254),
    "d3": lambda: conv_case("D3", 128, 256, 4, 2, 0, 252),
    "d4": lambda: conv_case("D4", 256, 256, 4, 2, 0, 125),
    # rank 3: the reference's literal 128^3 volumes (GAN_final.py:107,167-201), batch 1
    "v_d1": lambda: conv_case("3D D1", 1, 64, 3, 1, 0, 128, rank=3, batch=1),
    "v_d2": lambda: conv_case("3D D2", 64, 128, 3, 1, 0, 126, rank=3, batch=1),
    "v_d3": lambda: conv_case("3D D3", 128, 256, 4, 2, 0, 124, rank=3, batch=1),
    "v_d4": lambda: conv_case("3D D4", 256, 256, 4, 2, 0, 61, rank=3, batch=1),
    "v_g1": lambda: conv_case("3D G first", 1, 16, 3, 2, 1, 128, rank=3, batch=1),
    "v_g16": lambda: conv_case("3D G", 16, 16, 3, 1, 1, 64, rank=3, batch=1),
    "v_g16_32": lambda: conv_case("3D G", 16, 32, 3, 2, 1, 64, rank=3, batch=1),
    "v_g32": lambda: conv_case("3D G", 32, 32, 3, 1, 1, 32, rank=3, batch=1),
    "v_g64": lambda: conv_case("3D G", 64, 64, 3, 1, 1, 16, rank=3, batch=1),
    "v_g128": lambda: conv_case("3D G", 128, 128, 3, 1, 1, 16, rank=3, batch=1),
    "v_gT16": lambda: conv_case("3D G ConvT64->16", 16, 64, 3, 2, 1, 64, transposed=True, rank=3, batch=1),
    "v_gT1": lambda: conv_case("3D G ConvT32->1", 1, 32, 3, 2, 1, 128, transposed=True, rank=3, batch=1),
    "v_g11": lambda: conv_case("3D G 1->1", 1, 1, 3, 1, 1, 128, rank=3, batch=1),
    # batch norm
    "bn_g16": lambda: bn_case("G", 16, 128, ACT_PRELU),
    "bn_g64": lambda: bn_case("G", 64, 32, ACT_PRELU),
    "bn_d64": lambda: bn_case("D", 64, 254, ACT_LEAKY),
    "bn_d128": lambda: bn_case("D", 128, 252, ACT_LEAKY),
}

if __name__ == "__main__":
    want = sys.argv[1:]
    for key, fn in CASES.items():
        if want and not any(w in key for w in want):
            continue
        fn()
        torch.cuda.empty_cache()
