"""Generator / discriminator with the reference's module tree and explicit B200 execution plans.

Module classes, constructor signatures, attribute names and state-dict keys mirror
* /root/reference/code/GAN/GAN_final.py:92-122 (``CasNetGenerator``), :159-209 (``Discriminator``)
* /root/reference/test_runs/GAN.py:94-129, :136-198 (4-UNet generator, patch discriminator with activations)
* monai==0.4.0 ``UNet`` / ``Convolution`` / ``ResidualUnit`` / ``SkipConnection``
(torch.nn modules are used as parameter containers so default initialisation consumes the RNG exactly like the
reference).  Arithmetic never goes through torch: ``forward`` runs a static plan of libmpgan_sm100 kernels
(channels-last, bf16 or fp32) and records a tape; ``backward`` replays the tape in reverse with hand-written
gradient kernels, writing parameter gradients straight into the flat gradient buffer.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_LEAKY, ACT_NONE, ACT_PRELU
from .runtime import Runtime

_CONV = {2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}
_BN = {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}

DEFAULT_PRECISION = "bf16"


def remap_monai_keys(state_dict, to="legacy"):
    """MONAI changed ``Convolution``'s sub-module names when it introduced the ADN block: ``...unitN.norm.* / .act.*``
    (pre-ADN, the names this module tree uses) vs ``...unitN.adn.N.* / .adn.A.*`` (ADN).  The arithmetic is the same
    (conv -> norm -> act).  ``to="legacy"`` accepts either naming; ``to="adn"`` emits the ADN naming."""
    out = type(state_dict)() if isinstance(state_dict, dict) else {}
    for k, v in state_dict.items():
        if to == "legacy":
            k = k.replace(".adn.N.", ".norm.").replace(".adn.A.", ".act.")
        elif "generator." in k or k.startswith("model."):
            k = k.replace(".norm.", ".adn.N.").replace(".act.", ".adn.A.")
        out[k] = v
    return out


def _no_direct_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is executed by its owning network's kernel plan; call the "
                       "CasNetGenerator / Discriminator instead")


class Plan:
    """One forward pass worth of state: the tape the backward pops."""

    def __init__(self, rt, training, save, need_wgrad):
        self.rt, self.training, self.save, self.need_wgrad = rt, training, save, need_wgrad
        self.dtype = rt.dtype
        self.tape = []
        self.extra = {}
        self.fork = os.environ.get("MPGAN_NO_FORK", "0") != "1"
        self._arena = None
        self._arena_used = 0
        self.keepalive = []        # tensors read by weight-gradient kernels still in flight on the "wgrad" stream
        self.wgrad_forked = False
        self.wgrad_count = 0

    ARENA = 1 << 15  # fp64 slots: every BatchNorm statistic / backward sum of one pass, zeroed by ONE memset

    def zeros64(self, n, device):
        """n zeroed fp64 accumulators (a slice of the per-pass arena)."""
        if self._arena is None:
            self._arena = torch.zeros(self.ARENA, dtype=torch.float64, device=device)
        n8 = (n + 7) // 8 * 8
        if self._arena_used + n8 > self.ARENA:
            return torch.zeros(n, dtype=torch.float64, device=device)
        out = self._arena[self._arena_used:self._arena_used + n]
        self._arena_used += n8
        return out


def _new(ref, shape, dtype=None):
    return torch.empty(shape, dtype=dtype or ref.dtype, device=ref.device)


# -------------------------------------------------------------------------------------------------------------
# conv helpers shared by all layers
# -------------------------------------------------------------------------------------------------------------
def conv_apply(rec, x, out=None, stats=None):
    if rec.spec.transposed:
        return ops.conv_bprop(rec.spec, x, rec.w, rec.wt, rec.bias, out=out, stats=stats)
    return ops.conv_fprop(rec.spec, x, rec.w, rec.bias, out=out, stats=stats)


_SIDE = {}
N_WGRAD_STREAMS = int(os.environ.get("MPGAN_WGRAD_STREAMS", "1"))   # measured: 2-4 streams give nothing


def side_stream(device, which="branch"):
    """Auxiliary streams per device.  "branch": the residual branch of a ResidualUnit (forked and joined inside the
    unit).  "wgrad": every weight-gradient kernel of a backward pass -- they are off the critical path (nothing in
    the backward chain reads a weight gradient), so they are forked layer by layer and joined ONCE, at the end of
    the pass (``join_wgrad``).  Under CUDA-graph capture the streams become parallel graph branches."""
    key = (torch.device(device).index, which)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def join_wgrad(plan, device):
    """End of a backward pass: the optimizer (or the caller) may read the weight gradients after this."""
    if plan.wgrad_forked:
        for i in range(N_WGRAD_STREAMS):
            torch.cuda.current_stream().wait_stream(side_stream(device, f"wgrad{i}"))
        plan.wgrad_forked = False
    plan.keepalive.clear()


def conv_backward(rec, x, dy, plan, need_dx, bias_done=False, res=None):
    """x: the layer's input, dy: gradient of its output (contiguous or sliced).  Returns dx or None.
    bias_done: the bias gradient was already accumulated by the fused BatchNorm backward.
    res: a tensor of dx's shape added to it (gradient of a parallel branch) -- fused into the convolution's epilogue.
    The weight gradient only shares inputs with the data gradient: it is issued on the "wgrad" stream and not
    waited for until the end of the pass; (x, dy) are kept alive until then so the allocator cannot recycle them."""
    spec = rec.spec

    def wgrad():
        if spec.transposed:
            ops.conv_wgrad(spec, dy, x, rec.dw)
        else:
            ops.conv_wgrad(spec, x, dy, rec.dw)
        if rec.db is not None and not bias_done:
            ops.colsum(dy, rec.db)

    def dgrad():
        if spec.transposed:
            dx = ops.conv_fprop(spec, dy, rec.w, None)[0]
            return ops.add_copy(dx, res, dx) if res is not None else dx
        return ops.conv_bprop(spec, dy, rec.w, rec.wt, None, xs=tuple(x.shape[1:-1]), res=res)[0]

    if not plan.need_wgrad:
        return dgrad() if need_dx else None
    if not plan.fork:
        wgrad()
        return dgrad() if need_dx else None
    cur = torch.cuda.current_stream()
    side = side_stream(dy.device, f"wgrad{plan.wgrad_count % N_WGRAD_STREAMS}")   # round robin: wgrads are independent
    plan.wgrad_count += 1
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        wgrad()
    plan.keepalive.append((x, dy))
    plan.wgrad_forked = True
    return dgrad() if need_dx else None


def bn_act_forward(c, bn, act, alpha, leaky, res, out, plan, stats, fused_stats):
    ch = c.shape[-1]
    if plan.training and not fused_stats:
        ops.bn_stats(c, stats)
    buf = torch.empty((4, ch), dtype=torch.float32, device=c.device)
    mean, invstd, scale, shift = buf[0], buf[1], buf[2], buf[3]
    if out is None:
        out = _new(c, c.shape)
    if plan.training:
        ops.bn_train_apply(c, stats, bn, buf, act, alpha, leaky, res, out)
        plan.rt.bn_version += 1      # running statistics moved (raw-pointer write): folded inference weights are stale
    else:
        ops.bn_finalize(None, ops.pixels(c), bn, False, mean, invstd, scale, shift)
        ops.bn_act_apply(c, scale, shift, act, alpha, leaky, res, out)
    return out, (mean, invstd, scale, shift)


def bn_act_backward(dy, c, saved, bn, act, alpha_param, leaky, plan, trained, conv_db=None):
    """conv_db: bias-gradient buffer of the convolution that produced c (accumulated in the same pass)."""
    mean, invstd, scale, shift = saved
    ch = c.shape[-1]
    sums = plan.zeros64(2 * ch + 1, c.device)
    dc = _new(c, c.shape)
    wg = plan.need_wgrad
    ops.bn_act_bwd(dy, c, mean if trained else None, invstd if trained else None, scale, shift, act, alpha_param,
                   leaky, sums, bn.weight.grad if wg else None, bn.bias.grad if wg else None,
                   alpha_param.grad if (wg and alpha_param is not None) else None, dc,
                   dbias=conv_db if wg else None)
    return dc


# -------------------------------------------------------------------------------------------------------------
# MONAI-mirror blocks
# -------------------------------------------------------------------------------------------------------------
class Convolution(nn.Sequential):
    """monai.networks.blocks.Convolution (0.4.0): conv [-> BatchNorm -> PReLU]."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3, conv_only=False,
                 is_transposed=False):
        super().__init__()
        padding = (kernel_size - 1) // 2
        if is_transposed:
            conv = _CONVT[dimensions](in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                      padding=padding, output_padding=strides - 1, bias=True)
        else:
            conv = _CONV[dimensions](in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                     padding=padding, bias=True)
        self.add_module("conv", conv)
        self.conv_only = conv_only
        self.out_channels = out_channels
        if not conv_only:
            self.add_module("norm", _BN[dimensions](out_channels))
            self.add_module("act", nn.PReLU())

    forward = _no_direct_forward

    def out_spatial(self, sp):
        spec = ops.ConvSpec.from_module(self.conv)
        return spec.x_of_y(sp) if spec.transposed else spec.y_of_x(sp)

    def _fwd(self, x, plan, out=None, res=None):
        rec = plan.rt.rec[self.conv]
        if self.conv_only:
            if res is None:
                y, _ = conv_apply(rec, x, out=out)
            else:
                c, _ = conv_apply(rec, x)
                y = ops.add_copy(c, res, out if out is not None else _new(c, c.shape))
            if plan.save:
                plan.tape.append((x,))
            return y
        ch = self.out_channels
        if (not plan.training and not plan.save and plan.dtype == torch.bfloat16
                and os.environ.get("MPGAN_NO_EVAL_FUSION", "0") != "1"):
            # inference: BatchNorm folded into the weights, PReLU and the residual in the convolution's epilogue
            wf, bf = plan.rt.folded(self.conv, self.norm)
            y = ops.conv_act(rec.spec, x, wf, bf, self.act.weight, res=res, out=out)
            if y is not None:
                return y
        stats = plan.zeros64(2 * ch, x.device) if plan.training else None
        c, fused = conv_apply(rec, x, stats=stats)
        y, saved = bn_act_forward(c, self.norm, ACT_PRELU, self.act.weight, 0.0, res, out, plan, stats, fused)
        if plan.save:
            plan.tape.append((x, c, saved, plan.training))
        return y

    def _bwd(self, dy, plan, need_dx=True, res=None):
        rec = plan.rt.rec[self.conv]
        if self.conv_only:
            (x,) = plan.tape.pop()
            return conv_backward(rec, x, dy, plan, need_dx, res=res)
        x, c, saved, trained = plan.tape.pop()
        dc = bn_act_backward(dy, c, saved, self.norm, ACT_PRELU, self.act.weight, 0.0, plan, trained, conv_db=rec.db)
        return conv_backward(rec, x, dc, plan, need_dx, bias_done=True, res=res)


class ResidualUnit(nn.Module):
    """monai.networks.blocks.ResidualUnit (0.4.0)."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3, subunits=2,
                 last_conv_only=False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual = nn.Identity()
        self.out_channels = out_channels
        padding = (kernel_size - 1) // 2
        schannels, sstrides = in_channels, strides
        subunits = max(1, subunits)
        for su in range(subunits):
            conv_only = last_conv_only and su == (subunits - 1)
            self.conv.add_module(f"unit{su:d}", Convolution(dimensions, schannels, out_channels, strides=sstrides,
                                                            kernel_size=kernel_size, conv_only=conv_only))
            schannels, sstrides = out_channels, 1
        if np.prod(strides) != 1 or in_channels != out_channels:
            rkernel, rpad = kernel_size, padding
            if np.prod(strides) == 1:
                rkernel, rpad = 1, 0
            self.residual = _CONV[dimensions](in_channels, out_channels, rkernel, strides, rpad, bias=True)

    forward = _no_direct_forward

    def out_spatial(self, sp):
        return self.conv[0].out_spatial(sp)

    def _fwd(self, x, plan, out=None):
        has_res_conv = not isinstance(self.residual, nn.Identity)
        forked = False
        if has_res_conv:  # the residual convolution only shares its input with the main path: run it alongside
            rrec = plan.rt.rec[self.residual]
            sp = rrec.spec.y_of_x(tuple(x.shape[1:-1]))
            r = _new(x, (x.shape[0],) + tuple(sp) + (rrec.spec.cy,))
            if plan.fork:
                cur, side = torch.cuda.current_stream(), side_stream(x.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    conv_apply(rrec, x, out=r)
                forked = True
            else:
                conv_apply(rrec, x, out=r)
        else:
            r = x
        units = list(self.conv)
        h = x
        for u in units[:-1]:
            h = u._fwd(h, plan)
        if forked:
            cur.wait_stream(side)
        y = units[-1]._fwd(h, plan, out=out, res=r)
        if plan.save:
            plan.tape.append((x,))
        return y

    def _bwd(self, dy, plan, need_dx=True, res=None):
        """res: gradient of a branch parallel to the whole unit, added to the returned dx.  The branch sums
        (dx = main-path dgrad + residual-branch dgrad [+ res]) ride in the epilogues of the data-gradient kernels."""
        (x,) = plan.tape.pop()
        has_res_conv = not isinstance(self.residual, nn.Identity)
        units = list(self.conv)
        forked = False
        if has_res_conv:
            if plan.fork:
                cur, side = torch.cuda.current_stream(), side_stream(dy.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    dr = conv_backward(plan.rt.rec[self.residual], x, dy, plan, need_dx, res=res)
                forked = True
            else:
                dr = conv_backward(plan.rt.rec[self.residual], x, dy, plan, need_dx, res=res)
        elif res is not None and need_dx:
            dr = ops.add_copy(dy, res, _new(dy, dy.shape))
        else:
            dr = dy
        dh = dy
        for i in range(len(units) - 1, -1, -1):
            if i == 0 and forked:
                cur.wait_stream(side)   # the first unit's data gradient accumulates onto the residual branch's
            dh = units[i]._bwd(dh, plan, need_dx=(need_dx or i > 0), res=dr if (i == 0 and need_dx) else None)
        if not need_dx:
            return None
        return dh


class SkipConnection(nn.Module):
    """monai.networks.layers.SkipConnection (0.4.0): cat([x, submodule(x)], dim=1) -- realised by having the
    producers write straight into channel slices of one concat buffer."""

    def __init__(self, submodule, cat_dim=1):
        super().__init__()
        self.submodule = submodule
        self.cat_dim = cat_dim

    forward = _no_direct_forward


class UNet(nn.Module):
    """monai.networks.nets.UNet (0.4.0), act=PRELU, norm=BATCH, dropout=0."""

    def __init__(self, dimensions, in_channels, out_channels, channels, strides, kernel_size=3, up_kernel_size=3,
                 num_res_units=0, norm="batch", **_ignored):
        super().__init__()
        assert str(norm).lower().endswith("batch"), "the reference uses Norm.BATCH only"
        assert num_res_units > 0, "the reference uses num_res_units=2"
        self.dimensions = dimensions
        self.in_channels, self.out_channels = in_channels, out_channels
        self.channels, self.strides = tuple(channels), tuple(strides)
        self.kernel_size, self.up_kernel_size, self.num_res_units = kernel_size, up_kernel_size, num_res_units

        def _create_block(inc, outc, channels, strides, is_top):
            c, s = channels[0], strides[0]
            if len(channels) > 2:
                subblock = _create_block(c, c, channels[1:], strides[1:], False)
                upc = c * 2
            else:
                subblock = self._get_down_layer(c, channels[1], 1)
                upc = c + channels[1]
            down = self._get_down_layer(inc, c, s)
            up = self._get_up_layer(upc, outc, s, is_top)
            return nn.Sequential(down, SkipConnection(subblock), up)

        self.model = _create_block(in_channels, out_channels, self.channels, self.strides, True)

    def _get_down_layer(self, in_channels, out_channels, strides):
        return ResidualUnit(self.dimensions, in_channels, out_channels, strides=strides,
                            kernel_size=self.kernel_size, subunits=self.num_res_units)

    def _get_up_layer(self, in_channels, out_channels, strides, is_top):
        conv = Convolution(self.dimensions, in_channels, out_channels, strides=strides,
                           kernel_size=self.up_kernel_size, conv_only=False, is_transposed=True)
        ru = ResidualUnit(self.dimensions, out_channels, out_channels, strides=1, kernel_size=self.kernel_size,
                          subunits=1, last_conv_only=is_top)
        return nn.Sequential(conv, ru)

    forward = _no_direct_forward

    @staticmethod
    def _block_fwd(block, x, plan, out=None):
        down, skip, up = block[0], block[1], block[2]
        sub = skip.submodule
        cd = down.out_channels
        cs = sub.out_channels if isinstance(sub, ResidualUnit) else sub[2][1].out_channels
        sp = down.out_spatial(tuple(x.shape[1:-1]))
        cat = _new(x, (x.shape[0],) + tuple(sp) + (cd + cs,), plan.dtype)
        xd = down._fwd(x, plan, out=cat[..., :cd])
        if isinstance(sub, ResidualUnit):
            sub._fwd(xd, plan, out=cat[..., cd:])
        else:
            UNet._block_fwd(sub, xd, plan, out=cat[..., cd:])
        fused = UNet._tail_fwd(up, cat, plan, out)
        if fused is not None:
            return fused
        h = up[0]._fwd(cat, plan)
        return up[1]._fwd(h, plan, out=out)

    @staticmethod
    def _tail_fwd(up, cat, plan, out):
        """Top level of a 2-D bf16 UNet in training mode: ConvT(->1) + BatchNorm(1) + PReLU + ResidualUnit(1->1, conv
        only) run as the ConvTranspose (with fused statistics) plus ONE stencil kernel; the tape is the one the
        unfused layers would have recorded.  Returns None when the pattern does not apply."""
        conv0, ru = up[0], up[1]
        if (out is not None or plan.dtype != torch.bfloat16 or conv0.out_channels != 1
                or conv0.conv_only or cat.dim() != 4 or len(ru.conv) != 1 or not ru.conv[0].conv_only
                or not isinstance(ru.residual, nn.Identity) or os.environ.get("MPGAN_NO_TAIL_FUSION", "0") == "1"):
            return None
        rec0, rec1 = plan.rt.rec[conv0.conv], plan.rt.rec[ru.conv[0].conv]
        if not plan.training:
            # evaluation mode (inference): the same stencil kernel with the scale / shift of the running statistics
            if plan.save:
                return None
            c, _ = conv_apply(rec0, cat)
            saved = torch.empty((4, 1), dtype=torch.float32, device=c.device)
            ops.bn_finalize(None, ops.pixels(c), conv0.norm, False, saved[0], saved[1], saved[2], saved[3])
            _, y = ops.c1_tail_fwd(c, None, conv0.norm, saved, conv0.act.weight, rec1.w, rec1.bias, False)
            return y
        stats = plan.zeros64(2, cat.device)
        c, fused_stats = conv_apply(rec0, cat, stats=stats)
        if not fused_stats:
            ops.bn_stats(c, stats)
        saved = torch.empty((4, 1), dtype=torch.float32, device=c.device)
        h, y = ops.c1_tail_fwd(c, stats, conv0.norm, saved, conv0.act.weight, rec1.w, rec1.bias, plan.save)
        plan.rt.bn_version += 1
        if plan.save:
            plan.tape.append((cat, c, (saved[0], saved[1], saved[2], saved[3]), plan.training))   # Convolution (ConvT)
            plan.tape.append((h,))                                                              # conv-only Convolution
            plan.tape.append((h,))                                                              # ResidualUnit
        return y

    @staticmethod
    def _tail_bwd(up, dy, plan):
        """Backward of the fused tail (see _tail_fwd): the 1->1 conv's weight gradient goes to the weight-gradient
        stream as usual; its data gradient, the residual add and the BatchNorm(1) backward reduction are one kernel,
        the BatchNorm apply the second; then the ConvTranspose backward.  None when the pattern does not apply."""
        conv0, ru = up[0], up[1]
        if (plan.dtype != torch.bfloat16 or conv0.out_channels != 1 or conv0.conv_only or dy.dim() != 4
                or len(ru.conv) != 1 or not ru.conv[0].conv_only or not isinstance(ru.residual, nn.Identity)
                or len(plan.tape) < 3 or len(plan.tape[-3]) != 4 or not plan.tape[-3][3]      # trained BatchNorm only
                or os.environ.get("MPGAN_NO_TAIL_FUSION", "0") == "1"):
            return None
        rec0, rec1 = plan.rt.rec[conv0.conv], plan.rt.rec[ru.conv[0].conv]
        plan.tape.pop()                       # ResidualUnit (h,)
        (h,) = plan.tape.pop()                # conv-only Convolution (h,)
        cat, c, saved, _ = plan.tape.pop()    # ConvTranspose Convolution
        dyc = dy if dy.is_contiguous() else dy.contiguous()
        conv_backward(rec1, h, dyc, plan, need_dx=False)          # weight / bias gradient of the 1->1 conv
        wg = plan.need_wgrad
        sums = plan.zeros64(3, dy.device)
        dc = ops.c1_tail_bwd(dyc, c, saved, conv0.norm, conv0.act.weight, rec1.w, sums,
                             conv0.norm.weight.grad if wg else None, conv0.norm.bias.grad if wg else None,
                             conv0.act.weight.grad if wg else None, rec0.db if wg else None)
        return conv_backward(rec0, cat, dc, plan, need_dx=True, bias_done=True)

    @staticmethod
    def _block_bwd(block, dy, plan, need_dx=True, res=None):
        down, skip, up = block[0], block[1], block[2]
        sub = skip.submodule
        cd = down.out_channels
        dcat = UNet._tail_bwd(up, dy, plan)
        if dcat is None:
            dh = up[1]._bwd(dy, plan)
            dcat = up[0]._bwd(dh, plan)
        # the sub-block's gradient is the upper channel slice of dcat: one packing copy (7 us) lets its BatchNorm
        # backward run on the contiguous streaming kernels (2 x 8 us) instead of the strided ones (2 x 20 us)
        dsub = dcat[..., cd:]
        if dsub.dtype == torch.bfloat16 and os.environ.get("MPGAN_NO_PACK_DSUB", "0") != "1":
            dsub = ops.add_copy(dsub, None, _new(dsub, dsub.shape))
        # d(down output) = skip slice of dcat + the sub-block's input gradient: the slice is handed down as `res` and
        # added in the epilogue of the sub-block's last data-gradient kernel
        dskip = dcat[..., :cd]
        if isinstance(sub, ResidualUnit):
            dxd = sub._bwd(dsub, plan, res=dskip)
        else:
            dxd = UNet._block_bwd(sub, dsub, plan, res=dskip)
        return down._bwd(dxd, plan, need_dx=need_dx, res=res)

    def _fwd(self, x, plan, out=None):
        return UNet._block_fwd(self.model, x, plan, out=out)

    def _bwd(self, dy, plan, need_dx=True):
        return UNet._block_bwd(self.model, dy, plan, need_dx=need_dx)


# -------------------------------------------------------------------------------------------------------------
# autograd bridge: one Function per network call
# -------------------------------------------------------------------------------------------------------------
class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, net, want_acts, save):
        y, plan = net.run_forward(x, save=save)
        ctx.net, ctx.plan = net, plan
        ctx.set_materialize_grads(False)
        if want_acts:
            acts = plan.extra["acts"]
            ctx.n_acts = len(acts)
            return (y,) + tuple(acts)
        ctx.n_acts = 0
        return y

    @staticmethod
    def backward(ctx, dy, *dacts):
        if ctx.plan is None or not ctx.plan.save:
            raise RuntimeError("backward through a network call that recorded no tape")
        extra = {i: g for i, g in enumerate(dacts) if g is not None} if ctx.n_acts else None
        dx = ctx.net.run_backward(ctx.plan, dy, need_dx=ctx.needs_input_grad[0], act_grads=extra)
        ctx.plan = None
        return dx, None, None, None, None


class _PlanNet(nn.Module):
    """Shared plumbing of the two networks."""

    def _init_runtime(self, precision):
        self.precision = precision or DEFAULT_PRECISION
        self._runtime = None

    @property
    def runtime(self):
        if self._runtime is None or self._runtime.precision != self.precision:
            self._runtime = Runtime(self, self.precision)
        return self._runtime

    def set_precision(self, precision):
        assert precision in ("bf16", "fp32")
        self.precision = precision
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts both MONAI namings of the BatchNorm / PReLU keys (``remap_monai_keys``)."""
        return super().load_state_dict(remap_monai_keys(state_dict, "legacy"), strict=strict, **kw)

    def _call(self, x, want_acts=False):
        """Route a user-level call through the autograd bridge (tape recorded only when a backward can follow)."""
        if not x.is_cuda:
            raise RuntimeError("mpgan networks need CUDA tensors on a B200: there is no CPU fallback")
        anchor = torch.zeros((), device=x.device)
        anchor.requires_grad_(any(p.requires_grad for p in self.parameters()))
        save = torch.is_grad_enabled() and (x.requires_grad or anchor.requires_grad)
        return _NetFunction.apply(x, anchor, self, want_acts, save)

    def _to_cl(self, x):
        """logical NC[D]HW (C == 1) fp32 -> channels-last view."""
        if x.dim() != 2 + self.dims or x.shape[1] != 1:
            raise RuntimeError(f"expected (N, 1, {'D, ' if self.dims == 3 else ''}H, W), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("mpgan networks need CUDA tensors on a B200: there is no CPU fallback")
        return x.detach().contiguous().reshape((x.shape[0],) + tuple(x.shape[2:]) + (1,))


# -------------------------------------------------------------------------------------------------------------
# Generator
# -------------------------------------------------------------------------------------------------------------
class CasNetGenerator(_PlanNet):
    """GAN_final.py:92-122 (n_unet_blocks=6, channels (16,32,64,128)); test_runs/GAN.py:94-129 passes
    n_unet_blocks=4, channels (32,64,128,256), strides (2,2,2,2).  ``dims`` (2 or 3) defaults to
    ``len(img_shape) - 1``; the reference is always 3-D."""

    def __init__(self, img_shape, n_unet_blocks=6, dims=None, channels=(16, 32, 64, 128), strides=(2, 2, 2),
                 precision=None):
        super().__init__()
        self.img_shape = img_shape
        self.dims = dims if dims is not None else len(img_shape) - 1
        blocks = [UNet(dimensions=self.dims, in_channels=1, out_channels=1, channels=channels, strides=strides,
                       num_res_units=2, norm="batch") for _ in range(n_unet_blocks)]
        blocks.append(nn.Tanh())
        self.model = nn.Sequential(*blocks)
        self._init_runtime(precision)

    def forward(self, x):
        return self._call(x)

    def run_forward(self, x, save, need_wgrad=None):
        rt = self.runtime
        rt.ensure(x.device)
        plan = Plan(rt, self.training, save, rt.requires_grad() if need_wgrad is None else need_wgrad)
        xin = self._to_cl(x.float())
        h = ops.add_copy(xin, None, _new(xin, xin.shape, rt.dtype)) if rt.dtype != torch.float32 else xin
        for m in self.model:
            if isinstance(m, UNet):
                h = m._fwd(h, plan)
        hc = h if h.is_contiguous() else h.contiguous()
        y = ops.tanh_fwd(hc, torch.empty(hc.shape, dtype=torch.float32, device=x.device))
        if save:
            plan.tape.append((y,))
        return y.reshape(x.shape), plan

    def run_backward(self, plan, dy, need_dx=False, act_grads=None):
        rt = plan.rt
        (y,) = plan.tape.pop()
        dyc = dy.detach().float().contiguous().reshape(y.shape)
        d32 = ops.tanh_bwd(dyc, y, torch.empty_like(y))
        dh = ops.add_copy(d32, None, _new(d32, d32.shape, rt.dtype)) if rt.dtype != torch.float32 else d32
        unets = [m for m in self.model if isinstance(m, UNet)]
        for i in range(len(unets) - 1, -1, -1):
            dh = unets[i]._bwd(dh, plan, need_dx=(need_dx or i > 0))
        join_wgrad(plan, dyc.device)
        if not need_dx:
            return None
        d_in = dh if dh.dtype == torch.float32 else ops.add_copy(dh, None, _new(dh, dh.shape, torch.float32))
        return d_in.reshape(dy.shape)


# -------------------------------------------------------------------------------------------------------------
# Discriminators
# -------------------------------------------------------------------------------------------------------------
def _valid_out(size, k, s):
    return (size - k) // s + 1


class _ConvBnLeakyStack(_PlanNet):
    """model_conv (Conv -> BatchNorm -> LeakyReLU(0.2)) x 4 followed by model_linear (Flatten, Linear.., Sigmoid)."""

    def _build(self, layers, linear_widths, dims, spatial, precision):
        self.dims = dims
        mods, s = [], spatial
        for cin, cout, k, st in layers:
            mods += [_CONV[dims](cin, cout, kernel_size=k, stride=st), _BN[dims](cout), nn.LeakyReLU(0.2, inplace=True)]
            s = _valid_out(s, k, st)
        self.model_conv = nn.Sequential(*mods)
        lin, fan = [nn.Flatten()], layers[-1][1] * s ** dims
        for wdt in linear_widths:
            lin.append(nn.Linear(fan, wdt))
            fan = wdt
        lin.append(nn.Sigmoid())
        self.model_linear = nn.Sequential(*lin)
        self._init_runtime(precision)

    def _run(self, x, save, want_acts, need_wgrad=None, logical_acts=True):
        rt = self.runtime
        rt.ensure(x.device)
        plan = Plan(rt, self.training, save, rt.requires_grad() if need_wgrad is None else need_wgrad)
        xin = self._to_cl(x.float())
        h = ops.add_copy(xin, None, _new(xin, xin.shape, rt.dtype)) if rt.dtype != torch.float32 else xin
        acts = []
        convs = [m for m in self.model_conv if isinstance(m, (nn.Conv2d, nn.Conv3d))]
        bns = [m for m in self.model_conv if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d))]
        for conv, bn in zip(convs, bns):
            rec = rt.rec[conv]
            ch = conv.out_channels
            stats = plan.zeros64(2 * ch, x.device) if plan.training else None
            c, fused = conv_apply(rec, h, stats=stats)
            if want_acts:  # test_runs/GAN.py:186-190 clones conv, bn and (in-place) lrelu outputs separately
                bn_out, saved = bn_act_forward(c, bn, ACT_NONE, None, 0.0, None, None, plan, stats, fused)
                y = ops.bn_act_apply(bn_out, None, None, ACT_LEAKY, None, 0.2, None, _new(c, c.shape))
                acts += [("sp", c), ("sp", bn_out), ("sp", y)]
            else:
                bn_out = None
                y, saved = bn_act_forward(c, bn, ACT_LEAKY, None, 0.2, None, None, plan, stats, fused)
            if save:
                plan.tape.append((h, c, saved, plan.training, bn_out))
            h = y
        n = h.shape[0]
        feat_shape = tuple(h.shape)
        hf = h.reshape(n, -1)  # channels-last flatten order (spatial..., C)
        if want_acts:
            acts.append(("flat", h))
        linears = [m for m in self.model_linear if isinstance(m, nn.Linear)]
        z_in, first = hf, True
        lin_tape = []
        for lin in linears:
            j, k = lin.weight.shape
            if first:  # weight columns are in (C, spatial) order: permute to channels-last, in the compute dtype
                c_last, spatial = feat_shape[-1], k // feat_shape[-1]
                wcl = ops.permute_flatten(lin.weight.detach(), torch.empty((j, k), dtype=rt.dtype, device=x.device), j,
                                          c_last, spatial, True)
            else:
                wcl = lin.weight.detach()
            use_gemm = first and ops.linear_tc_ok(z_in, j, k)
            if use_gemm:   # a real GEMM (patch discriminator: 4 096 patches x 32 768 -> 64): tcgen05, not the split-K GEMV
                z = ops.linear_tc_fwd(z_in, wcl, lin.bias.detach())
            else:
                z = torch.zeros((n, j), dtype=torch.float32, device=x.device)
                ops.linear_fwd(z_in, wcl, lin.bias.detach(), z)
            lin_tape.append((z_in, wcl, first, use_gemm))
            if want_acts:
                acts.append(("raw", z))
            z_in, first = z, False
        p = ops.sigmoid_fwd(z_in, torch.empty_like(z_in))
        if want_acts:
            acts.append(("raw", p))
            if logical_acts:
                plan.extra["acts"] = self._acts_to_logical(acts)
            else:   # the fused training step consumes the activations in their internal (channels-last) layout
                plan.extra["acts_raw"] = [a for _, a in acts]
        if save:
            plan.tape.append((lin_tape, p, feat_shape))
        return p, plan

    def _acts_to_logical(self, acts):
        """channels-last internals -> the reference's NC[D]HW fp32 tensors (activation dict values).  This is
        layout glue for the autograd-compatible API only; the fused training step never leaves channels-last."""
        nd = self.dims
        perm = (0, nd + 1) + tuple(range(1, nd + 1))
        out = []
        for kind, a in acts:
            if kind == "sp":
                out.append(a.float().permute(perm).contiguous())
            elif kind == "flat":  # nn.Flatten of the NC[D]HW tensor
                out.append(a.float().permute(perm).reshape(a.shape[0], -1))
            else:
                out.append(a.float().clone())
        return out

    def run_backward(self, plan, dprob, need_dx=True, act_grads=None, internal=False, on_layer_done=None):
        """act_grads: {activation index: gradient} -- logical NC[D]HW fp32 tensors (the autograd bridge) or, with
        ``internal=True``, tensors in the activations' own layout / dtype (fused perceptual step).
        on_layer_done(i, stream): called after conv layer i's backward has been enqueued; ``stream`` is the stream that
        is ordered after ALL of that layer's (and the later layers') gradient kernels -- the weight-gradient side stream
        when the pass forks them, else the current stream (data-parallel bucket overlap, mpgan/ddp.py)."""
        rt = plan.rt
        lin_tape, p, feat_shape = plan.tape.pop()
        linears = [m for m in self.model_linear if isinstance(m, nn.Linear)]
        n_conv = 4
        ag = dict(act_grads or {})

        def dense(g):      # small (linear / probability) activations: a lazily defined L1 gradient is materialised
            return g.materialize() if isinstance(g, ops.LazyL1) else g

        def inject(target, g):   # target += gradient injected at an activation, in the activation's own layout
            if isinstance(g, ops.LazyL1):
                return g.add_to(target)
            return ops.add_copy(target, g, target)

        dp = dprob.detach().float().contiguous().reshape(p.shape)
        if 3 * n_conv + 1 + len(linears) in ag:
            dp = dp + dense(ag[3 * n_conv + 1 + len(linears)]).reshape(p.shape)
        dz = ops.sigmoid_bwd(dp, p, torch.empty_like(p))
        dfeat = None
        for li in range(len(linears) - 1, -1, -1):
            lin = linears[li]
            z_in, wcl, first, use_gemm = lin_tape[li]
            if (3 * n_conv + 1 + li) in ag:
                dz = dz + dense(ag[3 * n_conv + 1 + li]).reshape(dz.shape)
            j, k = lin.weight.shape
            dx = torch.empty_like(z_in)
            if use_gemm:
                dwcl = torch.zeros((j, k), dtype=torch.float32, device=dz.device) if plan.need_wgrad else None
                ops.linear_tc_bwd(z_in, wcl, dz, dx, dwcl, lin.bias.grad if plan.need_wgrad else None)
                if plan.need_wgrad:
                    ops.permute_flatten(dwcl, lin.weight.grad, j, feat_shape[-1], k // feat_shape[-1], False, True)
            elif plan.need_wgrad:
                if first:
                    dwcl = torch.zeros((j, k), dtype=torch.float32, device=dz.device)
                    ops.linear_bwd(z_in, wcl, dz, dx, dwcl, lin.bias.grad)
                    ops.permute_flatten(dwcl, lin.weight.grad, j, feat_shape[-1], k // feat_shape[-1], False, True)
                else:
                    ops.linear_bwd(z_in, wcl, dz, dx, lin.weight.grad, lin.bias.grad)
            else:
                ops.linear_bwd(z_in, wcl, dz, dx, None, None)
            dz = dx
        dfeat = dz  # (N, prod(feat)) in the compute dtype, channels-last order
        if 3 * n_conv in ag and internal:
            g = ag[3 * n_conv]
            dfeat = inject(dfeat.reshape(feat_shape), g if isinstance(g, ops.LazyL1) else g.reshape(feat_shape)).reshape(dfeat.shape)
        elif 3 * n_conv in ag:  # gradient w.r.t. the Flatten output (NCHW order)
            g = ag[3 * n_conv].reshape((feat_shape[0], feat_shape[-1]) + tuple(feat_shape[1:-1]))
            nd = self.dims
            g = g.permute((0,) + tuple(range(2, nd + 2)) + (1,)).reshape(dfeat.shape)
            dfeat = dfeat + g.to(dfeat.dtype)
        dh = dfeat.reshape(feat_shape)
        convs = [m for m in self.model_conv if isinstance(m, (nn.Conv2d, nn.Conv3d))]
        bns = [m for m in self.model_conv if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d))]
        nd = self.dims
        to_cl = (0,) + tuple(range(2, nd + 2)) + (1,)
        for i in range(len(convs) - 1, -1, -1):
            h_in, c, saved, trained, bn_out = plan.tape.pop()
            if (3 * i + 2) in ag:
                if internal:
                    dh = inject(dh.contiguous(), ag[3 * i + 2])
                else:
                    dh = dh + ag[3 * i + 2].permute(to_cl).to(dh.dtype)
            if bn_out is not None:
                # activations were exposed: LeakyReLU backward on the saved BN output, add the gradient injected
                # at the BN output, then the BatchNorm backward proper
                dbn = ops.act_bwd(dh.contiguous(), bn_out, ACT_LEAKY, 0.2, _new(bn_out, bn_out.shape))
                if (3 * i + 1) in ag:
                    dbn = inject(dbn, ag[3 * i + 1]) if internal else dbn + ag[3 * i + 1].permute(to_cl).to(dbn.dtype)
                dc = bn_act_backward(dbn, c, saved, bns[i], ACT_NONE, None, 0.0, plan, trained)
                fused_bias = False
            else:
                fused_bias = (3 * i) not in ag
                dc = bn_act_backward(dh, c, saved, bns[i], ACT_LEAKY, None, 0.2, plan, trained,
                                     conv_db=rt.rec[convs[i]].db if fused_bias else None)
            if (3 * i) in ag:
                dc = inject(dc, ag[3 * i]) if internal else dc + ag[3 * i].permute(to_cl).to(dc.dtype)
            dh = conv_backward(rt.rec[convs[i]], h_in, dc, plan, need_dx=(need_dx or i > 0), bias_done=fused_bias)
            if on_layer_done is not None:
                forked = plan.wgrad_forked and N_WGRAD_STREAMS == 1
                if plan.wgrad_forked and not forked:   # several weight-gradient streams: order them all behind the current one
                    join_wgrad(plan, dp.device)
                on_layer_done(i, side_stream(dp.device, "wgrad0") if forked else torch.cuda.current_stream())
        join_wgrad(plan, dp.device)
        if not need_dx:
            return None
        d_in = dh if dh.dtype == torch.float32 else ops.add_copy(dh, None, _new(dh, dh.shape, torch.float32))
        return d_in.reshape((dh.shape[0], 1) + tuple(dh.shape[1:-1]))


class Discriminator(_ConvBnLeakyStack):
    """GAN_final.py:159-209.  ``dims``/``spatial`` default to the reference's literal 3-D 128^3 (Linear fan-in
    256*29^3); the 2-D twin of the BASELINE configs is ``dims=2, spatial=256`` (fan-in 256*61*61)."""

    LAYERS = ((1, 64, 3, 1), (64, 128, 3, 1), (128, 256, 4, 2), (256, 256, 4, 2))

    def __init__(self, img_shape, use_perceptual=True, dims=None, spatial=None, precision=None):
        super().__init__()
        self.use_perceptual = use_perceptual
        dims = dims if dims is not None else len(img_shape) - 1
        spatial = spatial if spatial is not None else (128 if dims == 3 else img_shape[-1])
        self._build(self.LAYERS, (1,), dims, spatial, precision)

    def forward(self, img):
        return self._call(img)

    def run_forward(self, x, save, need_wgrad=None):
        return self._run(x, save, False, need_wgrad)


class PatchDiscriminator(_ConvBnLeakyStack):
    """test_runs/GAN.py:136-198: four k3 s1 valid convs to 512 channels, Linear(512*8^d, 64), Linear(64, 1); returns
    ``(validity, {0..15: activation})`` when ``use_perceptual``."""

    LAYERS = ((1, 64, 3, 1), (64, 128, 3, 1), (128, 256, 3, 1), (256, 512, 3, 1))

    def __init__(self, img_shape, use_perceptual=True, dims=None, spatial=16, precision=None):
        super().__init__()
        self.use_perceptual = use_perceptual
        dims = dims if dims is not None else len(img_shape) - 1
        self._build(self.LAYERS, (64, 1), dims, spatial, precision)

    def forward(self, x):
        if not self.use_perceptual:
            return self._call(x), {}
        outs = self._call(x, True)
        return outs[0], {i: a for i, a in enumerate(outs[1:])}

    def run_forward(self, x, save, need_wgrad=None, want_acts=None, logical_acts=True):
        return self._run(x, save, self.use_perceptual if want_acts is None else want_acts, need_wgrad, logical_acts)
