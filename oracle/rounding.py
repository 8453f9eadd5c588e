"""Rounding-matched oracle (TEST INFRASTRUCTURE ONLY): the fp32 CPU oracle with bf16 rounding injected at exactly the
points where the library's bf16 mode stores a tensor or feeds the tensor cores -- and nowhere else.

Why: the randomly initialised 6-UNet cascade with batch-statistics BatchNorm amplifies ANY operand rounding (bf16
weights alone move the generator output by 6e-2 and the parameter gradients by 70 %, tf32 everywhere still 1.5e-2 /
35 %: ``tools/precision_floor.py``, ``profiles/precision_floor_r2.md``), so a comparison of the bf16 path with the
plain fp32 oracle cannot separate "rounds where bf16 must round" from "computes something else".  The distance of THIS
oracle from the plain fp32 oracle is the FLOOR of the library's rounding scheme: what an ideal implementation that
rounds at these points must show.  The tests gate "bf16-mode error <= 1.2 x floor" (errors add in quadrature, so the
implementation's own error is bounded by 0.66 x the rounding noise); measured: 1.13e-1 vs 1.12e-1 (6 UNets), 1.18e-2 vs
1.18e-2 (one UNet) -- ``profiles/parity_r2.md``.

What it can NOT do (measured, same file): serve as an element-wise 1e-2 reference for the deep cascade.  bf16 rounding
is a chaotic map -- two evaluations that differ by d before a rounding point differ by ~sqrt(d * ulp) after it -- so a
1e-7 summation-order difference decorrelates the rounding noise completely within ~4 stored tensors; only shallow paths
(the discriminators: 3e-4 against this oracle) stay tight.

Rounding points of the bf16 mode (cross-modality-minipig-gan_b200/mpgan/nets.py, csrc/conv_tc.cu epilogues):
  * conv / conv-transpose / first-Linear weights: bf16 shadow of the fp32 master (round to nearest even);
  * every convolution output ``c`` (+ bias) is stored in bf16; BatchNorm statistics are those of the stored values;
  * ``a = PReLU(BN(c))`` is stored in bf16 -- except for the last unit of a ResidualUnit, whose residual is added in
    the same epilogue BEFORE the single rounding of the unit's output;
  * the network input and every UNet output (the one-channel trunk) are bf16; tanh / sigmoid / losses are fp32;
  * discriminator: ``LeakyReLU(BN(c))`` stored in bf16, Linear accumulates in fp32.
Everything else (accumulation, BatchNorm arithmetic, biases, PReLU slopes) is fp32 in both.  The backward pass is NOT
matched (its fused gradient sums round in a different association than autograd's), so only forward quantities --
outputs, losses, BatchNorm running statistics -- are compared against this oracle.
"""
import copy

import torch
import torch.nn as nn

from . import monai_unet

_CONVS = (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d)


def rnd(t, fmt="bf16"):
    if fmt == "bf16":
        return t.bfloat16().float()
    if fmt == "tf32":  # keep 10 explicit mantissa bits, round to nearest even (the tensor core itself truncates)
        i = t.contiguous().view(torch.int32)
        lsb = (i >> 13) & 1
        i = (i + 0xFFF + lsb) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(fmt)


class Round(torch.autograd.Function):
    """y = round(x); the gradient passes through (optionally rounded too)."""

    @staticmethod
    def forward(ctx, x, fmt, round_grad):
        ctx.fmt, ctx.g = fmt, round_grad
        return rnd(x, fmt)

    @staticmethod
    def backward(ctx, dy):
        return (rnd(dy, ctx.fmt) if ctx.g else dy), None, None


def _r(x):
    return Round.apply(x, "bf16", False)


def bf16_matched(net):
    """Deep copy of an oracle network (CasNetGenerator / Discriminator / PatchDiscriminator / GANOracle) that rounds
    to bf16 where the library's bf16 mode does (see the module docstring)."""
    net = copy.deepcopy(net)
    last_units = set()
    for m in net.modules():
        if isinstance(m, monai_unet.ResidualUnit):
            last_units.add(list(m.conv)[-1])
    for m in net.modules():
        if isinstance(m, _CONVS):
            with torch.no_grad():
                m.weight.copy_(rnd(m.weight))
            m.register_forward_hook(lambda mod, inp, out: _r(out))
        elif isinstance(m, monai_unet.Convolution) and m not in last_units and hasattr(m, "act"):
            m.register_forward_hook(lambda mod, inp, out: _r(out))
        elif isinstance(m, monai_unet.ResidualUnit):
            m.register_forward_hook(lambda mod, inp, out: _r(out))
        elif isinstance(m, nn.LeakyReLU):
            m.register_forward_hook(lambda mod, inp, out: _r(out))
        elif isinstance(m, monai_unet.UNet):
            m.register_forward_pre_hook(lambda mod, inp: (_r(inp[0]),))
    for name in ("discriminator", None):
        d = getattr(net, name, None) if name else net
        if d is not None and hasattr(d, "model_conv"):
            # (PatchDiscriminator.forward walks the modules itself, so the hook sits on the first conv, not the Sequential)
            d.model_conv[0].register_forward_pre_hook(lambda mod, inp: (_r(inp[0]),))
            first = [m for m in d.model_linear if isinstance(m, nn.Linear)][0]
            with torch.no_grad():
                first.weight.copy_(rnd(first.weight))
            if type(d).__name__ == "PatchDiscriminator" and getattr(d, "use_perceptual", False):
                # the 16 exposed activations include the BatchNorm outputs: the library materialises them (bf16)
                for m in d.model_conv:
                    if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                        m.register_forward_hook(lambda mod, inp, out: _r(out))
    return net
