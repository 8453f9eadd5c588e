#!/usr/bin/env python
"""CPU experiment (uses oracle/: test infrastructure, not product): what is the precision FLOOR of the generator
cascade for an implementation that stores activations / feeds the tensor cores in a reduced format?

The fp32 oracle generator (6 x UNet + tanh, train-mode BatchNorm) is run with rounding injected exactly where a
reduced-precision implementation has to round, everything else (accumulation, BatchNorm, PReLU) in fp32:

  w        conv / conv-transpose weights rounded (tensor-core operand)
  a        every convolution INPUT rounded (the stored activation `a` = PReLU(BN(c)) [+ residual])
  c        every convolution OUTPUT rounded (the stored pre-BatchNorm tensor `c`; BatchNorm statistics are then the
           statistics of the rounded values, as in the tcgen05 epilogue)
  grads    the same roundings on the gradient tensors flowing through those points (dgrad outputs, dc)

formats: bf16 (8-bit mantissa incl. hidden bit), tf32 (11-bit, round-to-nearest here; the tensor core truncates),
fp16-like 11-bit is the same as tf32 for this purpose.

Prints relative L2 of the generator output and of the global parameter-gradient vector against the fp32 run, for
1 UNet and for the 6-UNet cascade, at several sizes -- the numbers quoted in profiles/parity_r2.md.

    python tools/precision_floor.py [--sizes 64:2,128:4,256:8]
"""
import argparse
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from oracle.gan import synthetic_batch  # noqa: E402
from oracle.monai_unet import UNet  # noqa: E402
from oracle.nets import CasNetGenerator  # noqa: E402
from oracle.rounding import Round as _Round, rnd  # noqa: E402

_CONVS = (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d)


def instrument(net, fmt, w, a, c, grads):
    net = copy.deepcopy(net)
    hooks = []
    for m in net.modules():
        if isinstance(m, _CONVS):
            if w:
                with torch.no_grad():
                    m.weight.copy_(rnd(m.weight, fmt))
            if a:
                hooks.append(m.register_forward_pre_hook(lambda mod, inp: (_Round.apply(inp[0], fmt, grads),)))
            if c:
                hooks.append(m.register_forward_hook(lambda mod, inp, out: _Round.apply(out, fmt, grads)))
        if isinstance(m, UNet) and a:   # the one-channel trunk between UNets / before tanh is stored rounded too
            hooks.append(m.register_forward_hook(lambda mod, inp, out: _Round.apply(out, fmt, grads)))
    return net


def run(net, x, dy):
    for p in net.parameters():
        p.grad = None
    y = net(x)
    y.backward(dy)
    g = torch.cat([p.grad.flatten() for p in net.parameters()])
    return y.detach(), g


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm())


CASES = [
    ("bf16 w+a+c (+grads)   [this library's bf16 mode]", "bf16", True, True, True, True),
    ("bf16 w+a, c fp32      [c kept in TMEM / fp32]", "bf16", True, True, False, True),
    ("bf16 w only", "bf16", True, False, False, False),
    ("bf16 a only", "bf16", False, True, False, False),
    ("bf16 c only", "bf16", False, False, True, False),
    ("tf32 w+a+c (+grads)   [kind::tf32 operands, fp32 storage]", "tf32", True, True, True, True),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64:2,128:4,256:8")
    ap.add_argument("--autocast", action="store_true", help="also run torch's own CPU autocast(bf16) as a yardstick")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    print("| size x batch | UNets | rounding | rel-L2 G output | rel-L2 global grad |")
    print("|---|---:|---|---:|---:|")
    for item in args.sizes.split(","):
        size, batch = (int(v) for v in item.split(":"))
        for nblocks in (1, 6):
            torch.manual_seed(0)
            ref = CasNetGenerator((1, size, size), nblocks, 2)
            x = synthetic_batch(batch, 2, size, seed=1)["t1w"]
            dy = synthetic_batch(batch, 2, size, seed=5)["t2w"]
            y0, g0 = run(copy.deepcopy(ref), x, dy)
            for name, fmt, w, a, c, gr in CASES:
                y, g = run(instrument(ref, fmt, w, a, c, gr), x, dy)
                print(f"| {size}^2 x {batch} | {nblocks} | {name} | {rel(y, y0):.2e} | {rel(g, g0):.2e} |", flush=True)
            if args.autocast:
                net = copy.deepcopy(ref)
                for p in net.parameters():
                    p.grad = None
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    y = net(x)
                y = y.float()
                y.backward(dy)
                g = torch.cat([p.grad.flatten() for p in net.parameters()])
                print(f"| {size}^2 x {batch} | {nblocks} | torch autocast(bf16) on CPU | {rel(y, y0):.2e} | {rel(g, g0):.2e} |",
                      flush=True)


if __name__ == "__main__":
    main()
