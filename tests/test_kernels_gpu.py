"""Per-kernel parity on the GPU, through the C ABI, against plain torch fp32 references of the same op."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

from mpgan import ops  # noqa: E402
from mpgan._lib import ACT_LEAKY, ACT_NONE, ACT_PRELU  # noqa: E402

DEV = "cuda"


def cl(x, dtype=torch.float32):
    """NC[D]HW -> channels-last contiguous in `dtype`."""
    nd = x.dim() - 2
    return x.permute((0,) + tuple(range(2, nd + 2)) + (1,)).contiguous().to(dtype)


def uncl(x):
    nd = x.dim() - 2
    return x.float().permute((0, nd + 1) + tuple(range(1, nd + 1))).contiguous()


def oti(w, dtype=torch.float32):
    nd = w.dim() - 2
    return w.permute((0,) + tuple(range(2, nd + 2)) + (1,)).contiguous().to(dtype)


def rnd(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(shape, generator=g) * 2 - 1).to(DEV)


CONV_CASES = [
    # rank, n, cin, cout, size, k, s, p
    (2, 3, 1, 1, 33, 3, 1, 1),
    (2, 2, 1, 64, 41, 3, 1, 0),
    (2, 2, 1, 32, 30, 3, 2, 1),
    (2, 2, 1, 16, 20, 3, 2, 1),
    (2, 2, 1, 64, 19, 3, 1, 0),
    (2, 2, 1, 64, 40, 3, 1, 0),     # width % 8 == 0: the tcgen05 im2col kernel (conv_c1mma.cuh) takes bf16 fprop
    (2, 2, 1, 16, 48, 3, 2, 1),
    (2, 3, 1, 32, 24, 3, 1, 1),
    (2, 2, 1, 1, 32, 3, 1, 1),      # rows of 8k pixels: the run-based kernels with 16-byte window loads (conv_c1_fast.cu)
    (2, 2, 1, 16, 64, 3, 2, 1),
    (2, 3, 1, 32, 32, 3, 2, 1),
    (2, 3, 16, 16, 17, 3, 1, 1),
    (2, 2, 32, 1, 16, 3, 1, 1),
    (2, 2, 24, 40, 13, 4, 2, 0),
    (2, 2, 64, 128, 8, 1, 1, 0),
    (3, 2, 1, 8, 9, 3, 2, 1),
    (3, 1, 6, 10, 8, 3, 1, 0),
    (3, 2, 8, 8, 7, 4, 2, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("use_c1", [False, True])
def test_conv_generic_fwd_dgrad_wgrad(case, dtype, use_c1):
    """use_c1=False: the generic implicit-GEMM kernels only; True: the one-channel direct kernels where they apply."""
    rank, n, cin, cout, size, k, s, p = case
    conv = F.conv2d if rank == 2 else F.conv3d
    x = rnd(n, cin, *([size] * rank), seed=1)
    w = rnd(cout, cin, *([k] * rank), seed=2) * 0.2
    b = rnd(cout, seed=3)
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    x.requires_grad_(True), w.requires_grad_(True)
    y = conv(x, w, b, stride=s, padding=p)
    dy = rnd(*y.shape, seed=4)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y.backward(dy)
    spec = ops.ConvSpec(rank, cin, cout, k, s, p)
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV) if use_c1 else None
    yk, fused = ops.conv_fprop(spec, cl(x.detach(), dtype), oti(w.detach(), dtype), b, use_tc=False, use_c1=use_c1,
                               stats=stats)
    assert rel_l2(uncl(yk), y) <= tol
    if fused:  # BatchNorm statistics of the stored values, reduced by the same launch (one-channel fast path)
        ykd = yk.double().reshape(-1, cout)
        assert torch.allclose(stats[:cout], ykd.sum(0), rtol=1e-6, atol=1e-6)
        assert torch.allclose(stats[cout:], (ykd * ykd).sum(0), rtol=1e-6, atol=1e-6)
    dxk, _ = ops.conv_bprop(spec, cl(dy, dtype), oti(w.detach(), dtype), None, None, xs=(size,) * rank, use_tc=False,
                            use_c1=use_c1)
    assert rel_l2(uncl(dxk), x.grad) <= tol
    dw = torch.zeros(cout, k ** rank, cin, device=DEV)
    ops.conv_wgrad(spec, cl(x.detach(), dtype), cl(dy, dtype), dw, use_tc=False, use_c1=use_c1)
    assert rel_l2(dw, oti(w.grad)) <= (1e-4 if dtype == torch.float32 else 6e-3)
    db = torch.zeros(cout, device=DEV)
    ops.colsum(cl(dy, dtype), db)
    assert rel_l2(db, dy.sum(dim=[0] + list(range(2, rank + 2)))) <= 1e-4


@pytest.mark.parametrize("n,cout,size,s,p", [(2, 64, 40, 1, 0), (3, 16, 48, 2, 1), (2, 32, 30, 2, 1), (1, 64, 19, 1, 0)])
def test_c1_wgrad_through_tensor_cores(n, cout, size, s, p):
    """One-input-channel weight gradient = im2col + tcgen05 1x1 weight gradient + fold (conv_c1col.cu); accumulates."""
    x = rnd(n, 1, size, size, seed=1).bfloat16().float()
    w = (rnd(cout, 1, 3, 3, seed=2) * 0.2).requires_grad_(True)
    y = F.conv2d(x, w, None, stride=s, padding=p)
    dy = rnd(*y.shape, seed=4).bfloat16().float()
    y.backward(dy)
    spec = ops.ConvSpec(2, 1, cout, 3, s, p)
    dw = torch.ones(cout, 9, 1, device=DEV)                    # pre-existing content must be accumulated onto
    calls0 = ops._lib.ABI_CALLS
    ops.conv_wgrad(spec, cl(x, torch.bfloat16), cl(dy, torch.bfloat16), dw, force_c1col=True)
    assert ops._lib.ABI_CALLS - calls0 == 3                    # im2col, tcgen05 wgrad, fold: the composite path ran
    assert rel_l2(dw - 1.0, oti(w.grad)) <= 6e-3


@pytest.mark.parametrize("n,cout,size,p", [(2, 64, 40, 0), (3, 16, 33, 1), (1, 128, 24, 0)])
def test_c1_dgrad_through_tensor_cores(n, cout, size, p):
    """Data gradient of a one-input-channel stride-1 layer = halo tcgen05 kernel with weights padded to 16 columns."""
    x = rnd(n, 1, size, size, seed=1).requires_grad_(True)
    w = (rnd(cout, 1, 3, 3, seed=2) * 0.2).bfloat16().float()
    y = F.conv2d(x, w, None, stride=1, padding=p)
    dy = rnd(*y.shape, seed=4).bfloat16().float()
    y.backward(dy)
    spec = ops.ConvSpec(2, 1, cout, 3, 1, p)
    wo = oti(w, torch.bfloat16).reshape(cout, 9, 1)
    wt = None   # one-input-channel layers keep no transposed shadow
    r = rnd(n, size, size, 1, seed=5).bfloat16()
    dx, _ = ops.conv_bprop(spec, cl(dy, torch.bfloat16), wo, wt, None, xs=(size, size), force_c1out=True)
    assert rel_l2(uncl(dx), x.grad) <= 6e-3
    dx2, _ = ops.conv_bprop(spec, cl(dy, torch.bfloat16), wo, wt, None, xs=(size, size), force_c1out=True, res=r)
    assert rel_l2(dx2.float() - r.float(), dx.float()) <= 6e-3


@pytest.mark.parametrize("n,cout,size,s", [(2, 16, 64, 2), (3, 32, 32, 2), (2, 16, 24, 1), (2, 16, 30, 2)])
def test_c1_conv_act_inference_fusion(n, cout, size, s):
    """One-input-channel layer in evaluation mode: conv + folded bias + PReLU in one launch (mpgan_c1_conv_act); layers the
    run-based kernel does not take (width not a multiple of 8 at stride 2) report "not covered" and conv_act returns None."""
    x = rnd(n, 1, size, size, seed=1).bfloat16()
    w = (rnd(cout, 1, 3, 3, seed=2) * 0.3).bfloat16()
    b = rnd(cout, seed=3)
    slope = torch.tensor([0.2], device=DEV)
    ref = F.prelu(F.conv2d(x.float(), w.float(), b, stride=s, padding=1), slope)
    spec = ops.ConvSpec(2, 1, cout, 3, s, 1)
    y = ops.conv_act(spec, cl(x, torch.bfloat16), oti(w, torch.bfloat16), b, slope)
    if s == 2 and size % 8 != 0:
        assert y is None
        return
    assert y is not None and y.shape == (n, ref.shape[2], ref.shape[3], cout)
    assert rel_l2(uncl(y), ref) <= 6e-3


@pytest.mark.parametrize("n,cin,size", [(2, 32, 24), (3, 16, 40), (1, 64, 16)])
def test_convt_into_one_channel_through_tensor_cores(n, cin, size):
    """ConvTranspose2d(cin -> 1, k3 s2 p1 op1) forward (== data gradient of a 1 -> cin stride-2 conv) through the halo
    kernel's pixel-shuffle mode, with the fused one-channel BatchNorm statistics."""
    ct = torch.nn.ConvTranspose2d(cin, 1, 3, stride=2, padding=1, output_padding=1).to(DEV)
    with torch.no_grad():
        ct.weight.copy_(ct.weight.bfloat16().float())
    x = rnd(n, cin, size, size, seed=3).bfloat16().float()
    ref = ct(x)
    spec = ops.ConvSpec.from_module(ct)
    assert spec.transposed and spec.cx == 1 and spec.cy == cin
    w = ct.weight.detach().permute(0, 2, 3, 1).reshape(cin, 9, 1).contiguous().bfloat16()    # [cy][taps][cx]
    stats = torch.zeros(2, dtype=torch.float64, device=DEV)
    calls0 = ops._lib.ABI_CALLS
    y, fused = ops.conv_bprop(spec, cl(x, torch.bfloat16), w, None, ct.bias.detach(), stats=stats)
    assert ops._lib.ABI_CALLS - calls0 == 1 and fused
    assert y.shape == (n, 2 * size, 2 * size, 1)
    assert rel_l2(uncl(y), ref) <= 6e-3
    yd = y.double().reshape(-1)
    assert torch.allclose(stats, torch.stack([yd.sum(), (yd * yd).sum()]), rtol=1e-6, atol=1e-6)


def test_fused_unet_tail_forward_and_backward():
    """c1_tail_fwd / c1_tail_bwd (one-channel UNet tail: BatchNorm(1) -> PReLU -> conv3x3(1->1) + residual) against
    torch autograd on the same bf16-rounded tensors: outputs, running statistics and every gradient."""
    n, H, W = 3, 40, 24
    torch.manual_seed(1234)   # the 1 -> 1 convolution's default initialisation must not depend on the test order
    c = (rnd(n, 1, H, W, seed=1) * 3 + 0.7).bfloat16().float().requires_grad_(True)
    bn = torch.nn.BatchNorm2d(1).to(DEV)
    with torch.no_grad():
        bn.weight.fill_(1.3), bn.bias.fill_(-0.2)
    act = torch.nn.PReLU().to(DEV)
    conv = torch.nn.Conv2d(1, 1, 3, padding=1).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.bfloat16().float())
    bn2 = copy.deepcopy(bn)
    h_ref = act(bn(c))
    y_ref = conv(h_ref) + h_ref
    dy = rnd(n, 1, H, W, seed=2).bfloat16().float()
    y_ref.backward(dy)
    # ---- forward
    ccl = cl(c.detach(), torch.bfloat16)
    stats = torch.zeros(2, dtype=torch.float64, device=DEV)
    ops.bn_stats(ccl, stats)
    saved = torch.empty((4, 1), dtype=torch.float32, device=DEV)
    w9 = conv.weight.detach().reshape(9).bfloat16().contiguous()
    hk, yk = ops.c1_tail_fwd(ccl, stats, bn2, saved, act.weight.detach(), w9, conv.bias.detach(), True)
    assert rel_l2(uncl(hk), h_ref) <= 6e-3 and rel_l2(uncl(yk), y_ref) <= 6e-3
    assert torch.allclose(bn2.running_mean, bn.running_mean, rtol=1e-4, atol=1e-6)
    assert torch.allclose(bn2.running_var, bn.running_var, rtol=1e-4, atol=1e-6)
    assert int(bn2.num_batches_tracked) == 1
    # ---- backward down to dc (gradient of the BatchNorm input)
    sums = torch.zeros(3, dtype=torch.float64, device=DEV)
    dgamma, dbeta, dalpha, dbias = (torch.zeros(1, device=DEV) for _ in range(4))
    dc = ops.c1_tail_bwd(cl(dy, torch.bfloat16), ccl, (saved[0], saved[1], saved[2], saved[3]), bn2, act.weight.detach(), w9,
                         sums, dgamma, dbeta, dalpha, dbias)
    assert rel_l2(uncl(dc), c.grad) <= 1.5e-2          # bf16-rounded dh feeds the BatchNorm backward
    assert abs(float(dgamma) - float(bn.weight.grad)) <= 1e-2 * abs(float(bn.weight.grad)) + 1e-3
    assert abs(float(dbeta) - float(bn.bias.grad)) <= 1e-2 * abs(float(bn.bias.grad)) + 1e-3
    # the slope gradient is a sum over the negative half only, of bf16-rounded dh * bf16-rounded c: 2-3e-2 with some weights
    assert abs(float(dalpha) - float(act.weight.grad)) <= 3e-2 * abs(float(act.weight.grad)) + 1e-3


def test_conv_transpose_is_bprop():
    """ConvTranspose2d(k3,s2,p1,op1) forward == bprop of the underlying conv with the same OTI weight."""
    ct = torch.nn.ConvTranspose2d(24, 8, 3, stride=2, padding=1, output_padding=1).to(DEV)
    x = rnd(2, 24, 9, 9, seed=5)
    y = ct(x)
    spec = ops.ConvSpec.from_module(ct)
    assert spec.transposed and spec.cx == 8 and spec.cy == 24
    yk, _ = ops.conv_bprop(spec, cl(x), oti(ct.weight.detach()), None, ct.bias.detach(), use_tc=False)
    assert yk.shape == (2, 18, 18, 8)
    assert rel_l2(uncl(yk), y) <= 1e-5


def test_conv_channel_slices_of_concat_buffer():
    x = rnd(2, 16, 12, 12, seed=6)
    w = rnd(8, 16, 3, 3, seed=7) * 0.2
    y = F.conv2d(x, w, None, padding=1)
    big_in = torch.zeros(2, 12, 12, 40, device=DEV)
    big_in[..., 8:24] = cl(x)
    big_out = torch.full((2, 12, 12, 24), 7.0, device=DEV)
    spec = ops.ConvSpec(2, 16, 8, 3, 1, 1)
    ops.conv_fprop(spec, big_in[..., 8:24], oti(w), None, out=big_out[..., 16:24], use_tc=False)
    assert rel_l2(uncl(big_out[..., 16:24]), y) <= 1e-5
    assert torch.all(big_out[..., :16] == 7.0)


# ------------------------------------------------------------------ tcgen05 path
TC_CASES = [
    # n, cin, cout, h, w, k, s, p
    (2, 64, 128, 30, 30, 3, 1, 0),      # D layer 2 family (valid padding, odd extents)
    (2, 128, 256, 30, 30, 4, 2, 0),     # D layer 3 family (k4 s2: parity maps)
    (3, 256, 256, 29, 29, 4, 2, 0),     # D layer 4 family (odd input)
    (2, 16, 16, 32, 32, 3, 1, 1),       # G: 32-byte swizzle, padding via OOB fill
    (2, 16, 32, 32, 32, 3, 2, 1),       # G down, stride 2 with padding
    (2, 32, 64, 16, 16, 3, 2, 1),
    (2, 64, 128, 8, 8, 1, 1, 0),        # 1x1 residual conv == plain GEMM
    (2, 128, 128, 8, 8, 3, 1, 1),
    (5, 256, 512, 10, 10, 3, 1, 0),     # patch-D layer 4: two N tiles, tiny images
    (1, 192, 32, 6, 6, 3, 1, 1),        # three 64-channel chunks
    # stride-1 3x3 layers covered by the halo-resident kernel (conv_halo.cu): ragged 8x16 tiles, every C / N width
    (2, 32, 32, 20, 24, 3, 1, 1),
    (3, 64, 64, 17, 9, 3, 1, 1),
    (2, 32, 64, 19, 21, 3, 1, 0),
    (1, 64, 16, 40, 40, 3, 1, 1),
    (2, 128, 64, 12, 12, 3, 1, 1),
    (2, 16, 128, 33, 7, 3, 1, 1),
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_conv_fprop(case):
    n, cin, cout, h, w_, k, s, p = case
    x = rnd(n, cin, h, w_, seed=11).bfloat16()
    w = (rnd(cout, cin, k, k, seed=12) * 0.1).bfloat16()
    b = rnd(cout, seed=13)
    ref = F.conv2d(x.float(), w.float(), b, stride=s, padding=p)
    spec = ops.ConvSpec(2, cin, cout, k, s, p)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    y, fused = ops.conv_fprop(spec, cl(x, torch.bfloat16), oti(w, torch.bfloat16), b, stats=stats)
    assert fused, "tcgen05 path was not taken"
    torch.cuda.synchronize()
    assert rel_l2(uncl(y), ref) <= 6e-3
    yr = uncl(y).double()
    assert rel_l2(stats[:cout], yr.sum(dim=(0, 2, 3))) <= 1e-4
    assert rel_l2(stats[cout:], (yr * yr).sum(dim=(0, 2, 3))) <= 1e-4


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_conv_bprop(case):
    n, cin, cout, h, w_, k, s, p = case
    oh, ow = (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1
    dy = rnd(n, cout, oh, ow, seed=14).bfloat16()
    w = (rnd(cout, cin, k, k, seed=15) * 0.1).bfloat16()
    ref = torch.nn.grad.conv2d_input((n, cin, h, w_), w.float(), dy.float(), stride=s, padding=p)
    spec = ops.ConvSpec(2, cin, cout, k, s, p)
    wt = torch.empty(w.numel(), dtype=torch.bfloat16, device=DEV)
    ops.weight_transpose(oti(w, torch.bfloat16), wt, cout, k * k, cin)
    dx, _ = ops.conv_bprop(spec, cl(dy, torch.bfloat16), oti(w, torch.bfloat16), wt, None, xs=(h, w_))
    torch.cuda.synchronize()
    assert rel_l2(uncl(dx), ref) <= 6e-3


@pytest.mark.parametrize("case", [c for c in TC_CASES if c[1] in (16, 32, 64, 128, 256)])
def test_tc_conv_wgrad(case):
    n, cin, cout, h, w_, k, s, p = case
    oh, ow = (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1
    x = rnd(n, cin, h, w_, seed=16).bfloat16()
    dy = rnd(n, cout, oh, ow, seed=17).bfloat16()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, k, k), dy.float(), stride=s, padding=p)
    spec = ops.ConvSpec(2, cin, cout, k, s, p)
    dw = torch.zeros(cout, k * k, cin, device=DEV)
    ops.conv_wgrad(spec, cl(x, torch.bfloat16), cl(dy, torch.bfloat16), dw)
    torch.cuda.synchronize()
    assert rel_l2(dw, oti(ref)) <= 2e-3
    ops.conv_wgrad(spec, cl(x, torch.bfloat16), cl(dy, torch.bfloat16), dw)  # accumulates
    assert rel_l2(dw, 2 * oti(ref)) <= 2e-3


# ------------------------------------------------------------------ stride-2 3x3 halo engines (conv_s2.cuh)
S2_CASES = [
    # n, cx (fine-grid channels), cy (coarse-grid channels), (xh, xw)
    (2, 16, 32, (40, 24)),      # G level-2 down conv / its data gradient; ragged 16 x 8 tiles
    (3, 32, 64, (36, 20)),      # G level-3 down conv
    (1, 16, 64, (64, 48)),      # ConvTranspose(64 -> 16) and its data gradient
    (2, 32, 192, (24, 40)),     # ConvTranspose(192 -> 32): three 64-channel planes; N = 192 data gradient
    (2, 32, 128, (20, 20)),
]


@pytest.mark.parametrize("case", S2_CASES)
def test_s2_halo_fprop(case):
    """Conv k3 s2 p1 (forward of the down convolutions == data gradient of a ConvTranspose) through the space-to-depth
    halo engine: input and output are channel slices of wider buffers, bias, fused statistics where the engine has them."""
    n, cx, cy, (xh, xw) = case
    x = rnd(n, cx, xh, xw, seed=21).bfloat16()
    w = (rnd(cy, cx, 3, 3, seed=22) * 0.1).bfloat16()
    b = rnd(cy, seed=23)
    ref = F.conv2d(x.float(), w.float(), b, stride=2, padding=1)
    spec = ops.ConvSpec(2, cx, cy, 3, 2, 1)
    big_in = torch.zeros(n, xh, xw, cx + 16, dtype=torch.bfloat16, device=DEV)
    big_in[..., 8:8 + cx] = cl(x, torch.bfloat16)
    big_out = torch.full((n, xh // 2, xw // 2, cy + 32), 3.0, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2 * cy, dtype=torch.float64, device=DEV) if cy <= 64 else None
    y, fused = ops.conv_fprop(spec, big_in[..., 8:8 + cx], oti(w, torch.bfloat16), b, out=big_out[..., 16:16 + cy], stats=stats)
    torch.cuda.synchronize()
    assert rel_l2(uncl(y), ref) <= 6e-3
    assert torch.all(big_out[..., :16] == 3.0) and torch.all(big_out[..., 16 + cy:] == 3.0)
    if stats is not None:
        assert fused
        yr = uncl(y).double()
        assert rel_l2(stats[:cy], yr.sum(dim=(0, 2, 3))) <= 1e-4
        assert rel_l2(stats[cy:], (yr * yr).sum(dim=(0, 2, 3))) <= 1e-4
    # inference fusion: PReLU slope + residual in the epilogue
    slope = torch.tensor([0.25], device=DEV)
    r = rnd(n, xh // 2, xw // 2, cy, seed=24).bfloat16()
    ya = ops.conv_act(spec, cl(x, torch.bfloat16), oti(w, torch.bfloat16), b, slope, res=r)
    assert ya is not None
    assert rel_l2(uncl(ya), F.prelu(ref, slope) + uncl(r)) <= 6e-3


@pytest.mark.parametrize("case", S2_CASES)
def test_s2_halo_bprop(case):
    """ConvTranspose k3 s2 p1 op1 forward == data gradient of a stride-2 convolution, through the pixel-shuffle halo
    engine: bias + fused statistics (the ConvTranspose forward of a training step), and the residual-add form."""
    n, cx, cy, (xh, xw) = case
    yh, yw = xh // 2, xw // 2
    dy = rnd(n, cy, yh, yw, seed=25).bfloat16()
    w = (rnd(cy, cx, 3, 3, seed=26) * 0.1).bfloat16()
    b = rnd(cx, seed=27)
    ref = torch.nn.grad.conv2d_input((n, cx, xh, xw), w.float(), dy.float(), stride=2, padding=1)
    spec = ops.ConvSpec(2, cx, cy, 3, 2, 1)
    wt = torch.empty(w.numel(), dtype=torch.bfloat16, device=DEV)
    ops.weight_transpose(oti(w, torch.bfloat16), wt, cy, 9, cx)
    big_in = torch.zeros(n, yh, yw, cy + 8, dtype=torch.bfloat16, device=DEV)
    big_in[..., 8:] = cl(dy, torch.bfloat16)
    stats = torch.zeros(2 * cx, dtype=torch.float64, device=DEV)
    dx, fused = ops.conv_bprop(spec, big_in[..., 8:], oti(w, torch.bfloat16), wt, b, xs=(xh, xw), stats=stats)
    torch.cuda.synchronize()
    assert fused
    assert rel_l2(uncl(dx), ref + b.view(1, -1, 1, 1)) <= 6e-3
    xr = uncl(dx).double()
    assert rel_l2(stats[:cx], xr.sum(dim=(0, 2, 3))) <= 1e-4
    assert rel_l2(stats[cx:], (xr * xr).sum(dim=(0, 2, 3))) <= 1e-4
    r = rnd(n, xh, xw, cx, seed=28).bfloat16()
    dx2, _ = ops.conv_bprop(spec, cl(dy, torch.bfloat16), oti(w, torch.bfloat16), wt, None, xs=(xh, xw), res=r)
    assert rel_l2(uncl(dx2), ref + uncl(r)) <= 6e-3
    slope = torch.tensor([0.25], device=DEV)
    spec_t = ops.ConvSpec(2, cx, cy, 3, 2, 1, transposed=True, output_padding=1)
    ya = ops.conv_act(spec_t, cl(dy, torch.bfloat16), wt, b, slope, res=r)
    assert ya is not None
    assert rel_l2(uncl(ya), F.prelu(ref + b.view(1, -1, 1, 1), slope) + uncl(r)) <= 6e-3


WGH_CASES = [
    # n, cx, cy, (xh, xw), stride, pad       -- conv_wgrad_halo.cuh: one X halo box per tile, taps as operand atoms
    (2, 16, 16, (40, 24), 1, 1), (3, 32, 32, (19, 21), 1, 1), (2, 16, 32, (23, 17), 1, 0), (1, 32, 64, (33, 40), 1, 1),
    (2, 16, 64, (16, 8), 1, 1),
    (2, 16, 32, (40, 24), 2, 1), (3, 32, 64, (36, 20), 2, 1), (1, 16, 64, (64, 48), 2, 1), (2, 32, 32, (20, 20), 2, 1),
]


@pytest.mark.parametrize("case", WGH_CASES)
def test_wgrad_halo(case):
    n, cx, cy, (xh, xw), st, p = case
    yh, yw = (xh + 2 * p - 3) // st + 1, (xw + 2 * p - 3) // st + 1
    x = rnd(n, cx, xh, xw, seed=31).bfloat16()
    dy = rnd(n, cy, yh, yw, seed=32).bfloat16()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cy, cx, 3, 3), dy.float(), stride=st, padding=p)
    spec = ops.ConvSpec(2, cx, cy, 3, st, p)
    big_x = torch.zeros(n, xh, xw, cx + 16, dtype=torch.bfloat16, device=DEV)     # channel slices of wider buffers
    big_x[..., 16:] = cl(x, torch.bfloat16)
    big_y = torch.zeros(n, yh, yw, cy + 8, dtype=torch.bfloat16, device=DEV)
    big_y[..., :cy] = cl(dy, torch.bfloat16)
    dw = torch.ones(cy, 9, cx, device=DEV)
    ops.conv_wgrad(spec, big_x[..., 16:], big_y[..., :cy], dw)
    torch.cuda.synchronize()
    assert rel_l2(dw - 1.0, oti(ref)) <= 2e-3
    ops.conv_wgrad(spec, cl(x, torch.bfloat16), cl(dy, torch.bfloat16), dw)     # accumulates
    assert rel_l2(dw - 1.0, 2 * oti(ref)) <= 2e-3


# ------------------------------------------------------------------ tcgen05 path, rank 3 (NDHWC, 5-D TMA boxes)
TC3_CASES = [
    # n, cin, cout, (d, h, w), k, s, p
    (2, 64, 128, (12, 11, 13), 3, 1, 0),     # D layer 2 family at 3-D (GAN_final.py:173-176): 27 taps, valid padding
    (2, 128, 256, (14, 14, 14), 4, 2, 0),    # D layer 3 family: 64 taps, 8 input-parity maps / 8 output-parity classes
    (1, 256, 256, (13, 11, 9), 4, 2, 0),     # D layer 4 family, odd extents
    (2, 16, 16, (8, 8, 8), 3, 1, 1),         # G unit: padding through OOB zero fill in all three dimensions
    (2, 16, 32, (8, 10, 12), 3, 2, 1),       # G down layer: stride 2 with padding
    (2, 32, 64, (6, 6, 6), 3, 2, 1),
    (3, 64, 128, (4, 4, 4), 1, 1, 0),        # 1x1x1 residual conv
    (1, 128, 128, (5, 6, 7), 3, 1, 1),
    (5, 256, 512, (10, 10, 10), 3, 1, 0),    # patch-D layer 4 at 3-D (test_runs/GAN.py:160-165): two N tiles
    (1, 192, 32, (5, 5, 5), 3, 1, 1),        # three 64-channel chunks
]


def _conv3_ref(case, seed):
    n, cin, cout, sp, k, s, p = case
    x = rnd(n, cin, *sp, seed=seed).bfloat16()
    w = (rnd(cout, cin, k, k, k, seed=seed + 1) * 0.1).bfloat16()
    return n, cin, cout, sp, k, s, p, x, w


@pytest.mark.parametrize("case", TC3_CASES)
def test_tc3_conv_fprop(case):
    n, cin, cout, sp, k, s, p, x, w = _conv3_ref(case, 41)
    b = rnd(cout, seed=43)
    ref = F.conv3d(x.float(), w.float(), b, stride=s, padding=p)
    spec = ops.ConvSpec(3, cin, cout, k, s, p)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    y, fused = ops.conv_fprop(spec, cl(x, torch.bfloat16), oti(w, torch.bfloat16), b, stats=stats)
    assert fused, "tcgen05 path was not taken"
    torch.cuda.synchronize()
    assert rel_l2(uncl(y), ref) <= 6e-3
    yr = uncl(y).double()
    assert rel_l2(stats[:cout], yr.sum(dim=(0, 2, 3, 4))) <= 1e-4
    assert rel_l2(stats[cout:], (yr * yr).sum(dim=(0, 2, 3, 4))) <= 1e-4


@pytest.mark.parametrize("case", TC3_CASES)
def test_tc3_conv_bprop(case):
    n, cin, cout, sp, k, s, p, _, w = _conv3_ref(case, 44)
    osp = tuple((v + 2 * p - k) // s + 1 for v in sp)
    dy = rnd(n, cout, *osp, seed=46).bfloat16()
    ref = torch.nn.grad.conv3d_input((n, cin) + sp, w.float(), dy.float(), stride=s, padding=p)
    spec = ops.ConvSpec(3, cin, cout, k, s, p)
    wt = torch.empty(w.numel(), dtype=torch.bfloat16, device=DEV)
    ops.weight_transpose(oti(w, torch.bfloat16), wt, cout, k ** 3, cin)
    res = rnd(n, *sp, cin, seed=47).bfloat16()
    dx, _ = ops.conv_bprop(spec, cl(dy, torch.bfloat16), oti(w, torch.bfloat16), wt, None, xs=sp)
    torch.cuda.synchronize()
    assert rel_l2(uncl(dx), ref) <= 6e-3
    dx2, _ = ops.conv_bprop(spec, cl(dy, torch.bfloat16), oti(w, torch.bfloat16), wt, None, xs=sp, res=res)   # fused branch sum
    assert rel_l2(uncl(dx2), ref + uncl(res)) <= 6e-3


@pytest.mark.parametrize("case", [c for c in TC3_CASES if c[1] in (16, 32, 64, 128, 256)])
def test_tc3_conv_wgrad(case):
    n, cin, cout, sp, k, s, p, x, _ = _conv3_ref(case, 48)
    osp = tuple((v + 2 * p - k) // s + 1 for v in sp)
    dy = rnd(n, cout, *osp, seed=50).bfloat16()
    ref = torch.nn.grad.conv3d_weight(x.float(), (cout, cin, k, k, k), dy.float(), stride=s, padding=p)
    spec = ops.ConvSpec(3, cin, cout, k, s, p)
    dw = torch.zeros(cout, k ** 3, cin, device=DEV)
    ops.conv_wgrad(spec, cl(x, torch.bfloat16), cl(dy, torch.bfloat16), dw)
    torch.cuda.synchronize()
    assert rel_l2(dw, oti(ref)) <= 2e-3
    ops.conv_wgrad(spec, cl(x, torch.bfloat16), cl(dy, torch.bfloat16), dw)  # accumulates
    assert rel_l2(dw, 2 * oti(ref)) <= 2e-3


TC3_C1_CASES = [
    # n, cout, (d, h, w), s, p  -- rank-3 layers with ONE input channel (conv_c1vol.cu)
    (2, 64, (10, 9, 11), 1, 0),      # D layer 1 at 3-D: Conv3d(1, 64, 3), valid padding (GAN_final.py:167-169)
    (2, 16, (12, 12, 12), 2, 1),     # UNet entry convolution 1 -> 16, stride 2
    (1, 32, (8, 10, 6), 2, 1),       # geometry of ConvTranspose3d(32 -> 1) (its forward = this layer's data gradient)
    (2, 16, (7, 7, 7), 1, 1),
    (3, 1, (9, 8, 10), 1, 1),        # UNet tail 1 -> 1: direct 27-point stencils
]


@pytest.mark.parametrize("case", TC3_C1_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_tc3_one_channel_layers(case, dtype):
    n, cout, sp, s, p = case
    if dtype == torch.float32 and cout != 1:
        pytest.skip("fp32 mode keeps the generic kernels for these layers (covered by test_conv_generic_fwd_dgrad_wgrad)")
    x = rnd(n, 1, *sp, seed=61).to(dtype).float().requires_grad_(True)
    w = (rnd(cout, 1, 3, 3, 3, seed=62) * 0.2).to(dtype).float().requires_grad_(True)
    b = rnd(cout, seed=63)
    y = F.conv3d(x, w, b, stride=s, padding=p)
    dy = rnd(*y.shape, seed=64).to(dtype).float()
    y.backward(dy)
    osp = tuple(y.shape[2:])
    spec = ops.ConvSpec(3, 1, cout, 3, s, p)
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-5
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV) if cout > 1 else None
    calls0 = ops._lib.ABI_CALLS
    yk, fused = ops.conv_fprop(spec, cl(x.detach(), dtype), oti(w.detach(), dtype), b, stats=stats)
    assert ops._lib.ABI_CALLS - calls0 == (1 if cout == 1 else 2), "the one-channel volume path was not taken"
    assert rel_l2(uncl(yk), y) <= tol
    if cout > 1:
        assert fused
        ykd = yk.double().reshape(-1, cout)
        assert rel_l2(stats[:cout], ykd.sum(0)) <= 1e-4 and rel_l2(stats[cout:], (ykd * ykd).sum(0)) <= 1e-4
    res = rnd(n, *sp, 1, seed=65).to(dtype)
    dxk, _ = ops.conv_bprop(spec, cl(dy, dtype), oti(w.detach(), dtype), None, None, xs=sp, res=res)
    assert rel_l2(uncl(dxk), x.grad + uncl(res)) <= tol
    dw = torch.ones(cout, 27, 1, device=DEV)                       # accumulates onto existing content
    ops.conv_wgrad(spec, cl(x.detach(), dtype), cl(dy, dtype), dw)
    assert rel_l2(dw - 1.0, oti(w.grad)) <= (6e-3 if dtype == torch.bfloat16 else 1e-4)


def test_tc3_inference_fused_layer():
    """Eval-mode fused layer (folded BatchNorm bias + PReLU + residual in the epilogue) on the rank-3 path."""
    n, cin, cout, sp = 2, 32, 32, (6, 7, 8)
    x = rnd(n, cin, *sp, seed=51).bfloat16()
    w = (rnd(cout, cin, 3, 3, 3, seed=52) * 0.1).bfloat16()
    b, slope = rnd(cout, seed=53), torch.full((1,), 0.25, device=DEV)
    res = rnd(n, *sp, cout, seed=54).bfloat16()
    ref = F.prelu(F.conv3d(x.float(), w.float(), b, padding=1), slope) + uncl(res)
    y = ops.conv_act(ops.ConvSpec(3, cin, cout, 3, 1, 1), cl(x, torch.bfloat16), oti(w, torch.bfloat16), b, slope, res=res)
    assert y is not None and rel_l2(uncl(y), ref) <= 6e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [8 * 1000, 12345])
def test_l1_forward_backward_in_one_pass(dtype, n):
    """F.l1_loss forward + backward fused (vector body + scalar tail), plain and accumulating onto an existing gradient."""
    a, b = rnd(n, seed=81).to(dtype), rnd(n, seed=82).to(dtype)
    b[:7] = a[:7]                                                    # exact ties: zero sub-gradient like torch
    ar = a.float().clone().requires_grad_(True)
    ref = F.l1_loss(ar, b.float()) * 0.37
    ref.backward()
    loss = torch.zeros(1, device=DEV)
    da = ops.l1_fwd_bwd(a, b, 0.37, loss, torch.empty_like(a), False)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(da, ar.grad) <= (1e-6 if dtype == torch.float32 else 4e-3)
    base = rnd(n, seed=83).to(dtype)
    acc = ops.l1_fwd_bwd(a, b, 0.37, None, base.clone(), True)
    assert rel_l2(acc, base.float() + ar.grad) <= (1e-6 if dtype == torch.float32 else 4e-3)
    loss64 = torch.full((1,), 1e3, dtype=torch.float64, device=DEV)   # a large running total must not swallow the term
    ops.l1_fwd_bwd(a, b, 0.37, loss64, None, False)
    assert abs(float(loss64) - 1e3 - float(ref)) <= 2e-6 * abs(float(ref))
    lz = ops.LazyL1(a, b, 0.37, loss)
    assert torch.equal(lz.materialize(), da) and abs(float(loss) - 2 * float(ref)) <= 2e-5 * abs(float(ref))


def test_linear_small_weight_gradient():
    """Linear(64, 1) over 4 096 rows (patch discriminator tail): batch-split weight / bias gradient."""
    x, w = rnd(4096, 64, seed=84).float(), rnd(1, 64, seed=85).float()
    dz = rnd(4096, 1, seed=86).float()
    dx, dw, db = torch.empty_like(x), torch.ones(1, 64, device=DEV), torch.ones(1, device=DEV)
    ops.linear_bwd(x, w, dz, dx, dw, db)
    assert rel_l2(dx, dz @ w) <= 1e-5 and rel_l2(dw - 1.0, dz.t() @ x) <= 1e-4 and rel_l2(db - 1.0, dz.sum(0)) <= 1e-4


@pytest.mark.parametrize("b,k,j", [(256, 512, 64), (4096, 32768, 64), (512, 2048, 16), (1024, 1024, 128)])
def test_linear_as_tcgen05_gemm(b, k, j):
    """nn.Linear that is a real GEMM (patch discriminator: Linear(512 * 8 * 8, 64) over 4 096 patches,
    test_runs/GAN.py:176-181): forward, data gradient (N = K > 512 output columns, no bias), weight gradient (blocks of 256
    X channels) and bias gradient on the tcgen05 kernels against torch fp32 on the same bf16-rounded operands."""
    assert ops.linear_tc_ok(torch.empty((b, k), dtype=torch.bfloat16, device=DEV), j, k)
    x = rnd(b, k, seed=71).bfloat16()
    w = (rnd(j, k, seed=72) * (1.0 / k ** 0.5)).bfloat16()
    bias = rnd(j, seed=73)
    z = ops.linear_tc_fwd(x, w, bias)
    ref = x.float() @ w.float().t() + bias
    assert z.dtype == torch.float32 and rel_l2(z, ref) <= 4e-3           # (bf16 rounding of the stored result)
    dz = rnd(b, j, seed=74, ).float()
    dx = torch.empty_like(x)
    dw = torch.ones(j, k, device=DEV)
    db = torch.ones(j, device=DEV)
    ops.linear_tc_bwd(x, w, dz, dx, dw, db)
    dzr = dz.bfloat16().float()
    assert rel_l2(dx, dzr @ w.float()) <= 4e-3
    assert rel_l2(dw - 1.0, dzr.t() @ x.float()) <= 2e-3
    assert rel_l2(db - 1.0, dzr.sum(0)) <= 1e-3


def test_tc_conv_full_size_linearity():
    """BASELINE-size D layer 3 (32 x 252^2 x 128 -> 125^2 x 256): checked through linearity and a sampled window."""
    n, cin, cout, h = 8, 128, 256, 252
    x1, x2 = rnd(n, h, h, cin, seed=21).bfloat16(), rnd(n, h, h, cin, seed=22).bfloat16()
    w = (rnd(cout, 16, cin, seed=23) * 0.05).bfloat16()
    spec = ops.ConvSpec(2, cin, cout, 4, 2, 0)
    y1, _ = ops.conv_fprop(spec, x1, w, None)
    y2, _ = ops.conv_fprop(spec, x2, w, None)
    y12, _ = ops.conv_fprop(spec, (x1.float() + x2.float()).bfloat16(), w, None)
    assert y1.shape == (n, 125, 125, cout)
    assert rel_l2(y12.float(), y1.float() + y2.float()) <= 2e-2
    ref = F.conv2d(uncl(x1)[:1, :, :40, :40], w.view(cout, 4, 4, cin).permute(0, 3, 1, 2).float(), stride=2)
    assert rel_l2(uncl(y1)[:1, :, :19, :19], ref) <= 6e-3


# ------------------------------------------------------------------ batch norm / activations
@pytest.mark.parametrize("c,act", [(1, ACT_PRELU), (16, ACT_PRELU), (64, ACT_LEAKY), (24, ACT_NONE), (512, ACT_LEAKY)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fused", [False, True])
def test_bn_act_forward_backward(c, act, dtype, fused):
    """fused: mpgan_bn_train_apply (finalize + apply in one launch, the training path of the networks) instead of
    mpgan_bn_finalize followed by mpgan_bn_act_apply."""
    n, h, w_ = 3, 9, 11
    x = (rnd(n, c, h, w_, seed=31) * 2 + 0.3)
    res = rnd(n, c, h, w_, seed=32)
    if dtype == torch.bfloat16:
        x, res = x.bfloat16().float(), res.bfloat16().float()
    x.requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(c, seed=33) + 1.5), bn.bias.copy_(rnd(c, seed=34))
    bn_k = torch.nn.BatchNorm2d(c).to(DEV)
    bn_k.load_state_dict(bn.state_dict())
    alpha = torch.nn.Parameter(torch.tensor([0.25], device=DEV))
    z = bn(x)
    yr = F.prelu(z, alpha) if act == ACT_PRELU else (F.leaky_relu(z, 0.2) if act == ACT_LEAKY else z)
    yr = yr + res
    dy = rnd(n, c, h, w_, seed=35)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    yr.backward(dy)
    # kernels
    xc = cl(x.detach(), dtype)
    stats = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    ops.bn_stats(xc, stats)
    buf = torch.empty(4, c, device=DEV)
    a = alpha.detach() if act == ACT_PRELU else None
    if fused:
        y = ops.bn_train_apply(xc, stats, bn_k, buf, act, a, 0.2, cl(res, dtype), torch.empty_like(xc))
    else:
        ops.bn_finalize(stats, n * h * w_, bn_k, True, buf[0], buf[1], buf[2], buf[3])
        y = ops.bn_act_apply(xc, buf[2], buf[3], act, a, 0.2, cl(res, dtype), torch.empty_like(xc))
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    assert rel_l2(uncl(y), yr) <= tol
    assert rel_l2(bn_k.running_mean, bn.running_mean) <= 1e-5 and rel_l2(bn_k.running_var, bn.running_var) <= 1e-5
    assert int(bn_k.num_batches_tracked) == 1
    sums = torch.zeros(2 * c + 1, dtype=torch.float64, device=DEV)
    dg, db, da = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV), torch.zeros(1, device=DEV)
    dbias = torch.zeros(c, device=DEV)
    dx = ops.bn_act_bwd(cl(dy, dtype), xc, buf[0], buf[1], buf[2], buf[3], act, a, 0.2, sums, dg, db, da,
                        torch.empty_like(xc), dbias=dbias)
    assert torch.allclose(dbias, dx.float().sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-3 * float(dx.float().abs().max()))
    btol = 1e-4 if dtype == torch.float32 else 1e-2
    assert rel_l2(uncl(dx), x.grad) <= btol
    assert rel_l2(dg, bn.weight.grad) <= 1e-4 and rel_l2(db, bn.bias.grad) <= 1e-4
    if act == ACT_PRELU:
        assert rel_l2(da, alpha.grad) <= 1e-4


@pytest.mark.parametrize("n,c,h,w_,act,with_res", [(3, 64, 384, 384, ACT_LEAKY, False), (2, 256, 256, 256, ACT_LEAKY, False),
                                                   (1, 128, 451, 443, ACT_LEAKY, False), (4, 16, 128, 128, ACT_PRELU, True),
                                                   (9, 32, 67, 64, ACT_PRELU, False), (2, 512, 64, 64, ACT_NONE, True)])
def test_bn_streaming_kernels_bf16(n, c, h, w_, act, with_res):
    """Contiguous bf16 tensors >= 2 MB take the bulk-copy streaming kernels of bn_stream.cu (LeakyReLU / PReLU with its
    slope gradient / none; residual input; output written into a channel slice of a wider buffer; partial last chunk)."""
    x = (rnd(n, c, h, w_, seed=51) * 2 + 0.3).bfloat16().float().requires_grad_(True)
    dy = rnd(n, c, h, w_, seed=52).bfloat16().float()
    res = rnd(n, c, h, w_, seed=55).bfloat16().float() if with_res else None
    bn = torch.nn.BatchNorm2d(c).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(c, seed=53) + 1.5), bn.bias.copy_(rnd(c, seed=54))
    bn_k = torch.nn.BatchNorm2d(c).to(DEV)
    bn_k.load_state_dict(bn.state_dict())
    alpha = torch.nn.Parameter(torch.tensor([0.25], device=DEV))
    z = bn(x)
    yr = F.prelu(z, alpha) if act == ACT_PRELU else (F.leaky_relu(z, 0.2) if act == ACT_LEAKY else z)
    if with_res:
        yr = yr + res
    yr.backward(dy)
    xc = cl(x.detach(), torch.bfloat16)
    assert xc.numel() * 2 >= 2 << 20
    stats = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    ops.bn_stats(xc, stats)
    buf = torch.empty(4, c, device=DEV)
    a = alpha.detach() if act == ACT_PRELU else None
    resc = cl(res, torch.bfloat16) if with_res else None
    wide = torch.zeros(n, h, w_, c + 16, dtype=torch.bfloat16, device=DEV)   # concat buffer: write into [..., 8:8+c]
    y = ops.bn_train_apply(xc, stats, bn_k, buf, act, a, 0.2, resc, wide[..., 8:8 + c])
    assert rel_l2(uncl(y), yr) <= 8e-3
    assert float(wide[..., :8].abs().max()) == 0 and float(wide[..., 8 + c:].abs().max()) == 0
    assert rel_l2(bn_k.running_mean, bn.running_mean) <= 1e-5 and rel_l2(bn_k.running_var, bn.running_var) <= 1e-5
    y2 = ops.bn_act_apply(xc, buf[2], buf[3], act, a, 0.2, resc, torch.empty_like(xc))
    assert torch.equal(y.contiguous(), y2)
    sums = torch.zeros(2 * c + 1, dtype=torch.float64, device=DEV)
    dg, db, dbias = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    da = torch.zeros(1, device=DEV)
    dx = ops.bn_act_bwd(cl(dy, torch.bfloat16), xc, buf[0], buf[1], buf[2], buf[3], act, a, 0.2, sums, dg, db,
                        da if act == ACT_PRELU else None, torch.empty_like(xc), dbias=dbias)
    assert rel_l2(uncl(dx), x.grad) <= 1e-2
    assert rel_l2(dg, bn.weight.grad) <= 1e-4 and rel_l2(db, bn.bias.grad) <= 1e-4
    if act == ACT_PRELU:
        assert rel_l2(da, alpha.grad) <= 1e-4
    assert torch.allclose(dbias, dx.float().sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-3 * float(dx.float().abs().max()))


def test_bn_eval_mode():
    c = 16
    bn = torch.nn.BatchNorm2d(c).to(DEV).eval()
    with torch.no_grad():
        bn.running_mean.copy_(rnd(c, seed=41)), bn.running_var.copy_(rnd(c, seed=42).abs() + 0.5)
    x = rnd(2, c, 6, 6, seed=43)
    buf = torch.empty(4, c, device=DEV)
    ops.bn_finalize(None, 72, bn, False, buf[0], buf[1], buf[2], buf[3])
    y = ops.bn_act_apply(cl(x), buf[2], buf[3], ACT_NONE, None, 0.0, None, torch.empty_like(cl(x)))
    assert rel_l2(uncl(y), bn(x)) <= 1e-5
    assert int(bn.num_batches_tracked) == 0


# ------------------------------------------------------------------ linear / losses / adam / patches
@pytest.mark.parametrize("b,c,s,j", [(4, 8, 100, 1), (3, 16, 9, 64), (2, 256, 61 * 61, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_after_flatten(b, c, s, j, dtype):
    k = c * s
    lin = torch.nn.Linear(k, j).to(DEV)
    h = rnd(b, c, s, seed=51)           # NCHW-flatten order (c, spatial)
    if dtype == torch.bfloat16:
        h = h.bfloat16().float()
        with torch.no_grad():
            lin.weight.copy_(lin.weight.bfloat16().float())
    h.requires_grad_(True)
    z = lin(h.reshape(b, -1))
    dz = rnd(b, j, seed=52)
    z.backward(dz)
    hcl = h.detach().permute(0, 2, 1).contiguous().reshape(b, -1).to(dtype)   # (spatial, c) order
    wcl = ops.permute_flatten(lin.weight.detach(), torch.empty(j, k, dtype=dtype, device=DEV), j, c, s, True)
    zk = ops.linear_fwd(hcl, wcl, lin.bias.detach(), torch.zeros(b, j, device=DEV))
    tol = 2e-5 if dtype == torch.float32 else 2e-3
    assert rel_l2(zk, z) <= tol
    dx, dwcl, dbias = torch.empty_like(hcl), torch.zeros(j, k, device=DEV), torch.zeros(j, device=DEV)
    ops.linear_bwd(hcl, wcl, dz, dx, dwcl, dbias)
    dw = ops.permute_flatten(dwcl, torch.zeros(j, k, device=DEV), j, c, s, False, True)
    assert rel_l2(dw, lin.weight.grad) <= tol and rel_l2(dbias, lin.bias.grad) <= 1e-5
    assert rel_l2(dx.float().reshape(b, s, c).permute(0, 2, 1), h.grad) <= (tol if dtype == torch.float32 else 8e-3)


def test_sigmoid_bce_matches_torch_including_clamp():
    z = torch.tensor([[-200.0], [-20.0], [-1.0], [0.0], [0.3], [30.0], [200.0]], device=DEV, requires_grad=True)
    for target in (1.0, 0.9, 0.0):
        t = torch.full_like(z, target)
        p = torch.sigmoid(z)
        ref = F.binary_cross_entropy(p, t)
        (gz,) = torch.autograd.grad(ref, z)
        pk = ops.sigmoid_fwd(z.detach().contiguous(), torch.empty_like(z))
        loss = torch.zeros(1, device=DEV)
        ops.bce_fwd(pk, t, 1.0, loss)
        assert torch.equal(pk, p.detach())
        assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
        dz = ops.sigmoid_bwd(ops.bce_bwd(pk, t, 1.0, None, torch.empty_like(pk)), pk, torch.empty_like(pk))
        assert torch.allclose(dz, gz, rtol=1e-5, atol=1e-7)
    # saturated discriminator: clamped loss is exactly 100 and the gradient vanishes
    zs = torch.full((4, 1), -500.0, device=DEV)
    ps = ops.sigmoid_fwd(zs, torch.empty_like(zs))
    loss = torch.zeros(1, device=DEV)
    ops.bce_fwd(ps, torch.ones_like(ps), 1.0, loss)
    assert float(loss) == 100.0
    dz = ops.sigmoid_bwd(ops.bce_bwd(ps, torch.ones_like(ps), 1.0, None, torch.empty_like(ps)), ps, torch.empty_like(ps))
    assert float(dz.abs().max()) == 0.0


def test_l1_and_tanh():
    a, b = rnd(3, 1, 17, 19, seed=61).requires_grad_(True), rnd(3, 1, 17, 19, seed=62)
    ref = F.l1_loss(a, b)
    (ga,) = torch.autograd.grad(ref, a)
    loss = torch.zeros(1, device=DEV)
    ops.l1_fwd(a.detach(), b, 1.0, loss)
    assert abs(float(loss) - float(ref)) <= 1e-6
    da = ops.l1_bwd(a.detach(), b, 1.0, None, torch.zeros_like(b), False)
    assert torch.allclose(da, ga)
    x = rnd(1000, seed=63) * 3
    y = ops.tanh_fwd(x, torch.empty_like(x))
    assert torch.allclose(y, torch.tanh(x), atol=1e-6)
    dy = rnd(1000, seed=64)
    assert torch.allclose(ops.tanh_bwd(dy, y, torch.empty_like(x)), dy * (1 - torch.tanh(x) ** 2), atol=1e-6)


def test_adam_matches_torch_over_steps():
    n = 10007
    p = torch.nn.Parameter(rnd(n, seed=71))
    opt = torch.optim.Adam([p], lr=5e-4, betas=(0.5, 0.999))
    pk, m, v = p.detach().clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(4, device=DEV)
    shadow = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    for step in range(5):
        g = rnd(n, seed=80 + step) * (10.0 ** (step - 2))
        p.grad = g.clone()
        opt.step()
        ops.adam_step(pk, g, m, v, 5e-4, 0.5, 0.999, 1e-8, state, shadow)
        assert rel_l2(pk, p.detach()) <= 1e-6
    assert torch.equal(shadow, pk.bfloat16())
    assert int(state[0].view(torch.int32)) == 5


@pytest.mark.parametrize("rank", [2, 3])
def test_patch_gather_bit_exact_and_scatter_add(rank):
    b, ns, roi = 3, 7, 16
    sp = (40, 37) if rank == 2 else (20, 18, 21)
    vol = rnd(b, *sp, 1, seed=91)
    rng = np.random.RandomState(2)
    o = np.stack([rng.randint(0, s - roi + 1, size=(b, ns)) for s in sp], axis=-1)
    o_dev = torch.as_tensor(o.reshape(-1, rank), dtype=torch.int32, device=DEV)
    out = ops.patch_gather(vol, o_dev, ns, roi)
    for bi in range(b):
        for si in range(ns):
            sl = tuple(slice(int(v), int(v) + roi) for v in o[bi, si])
            assert torch.equal(out[bi * ns + si], vol[bi][sl])   # bit exact
    g = rnd(*out.shape, seed=92)
    dvol = ops.patch_scatter_add(g, o_dev, ns, roi, torch.zeros_like(vol))
    ref = torch.zeros_like(vol)
    for bi in range(b):
        for si in range(ns):
            sl = tuple(slice(int(v), int(v) + roi) for v in o[bi, si])
            ref[bi][sl] += g[bi * ns + si]
    assert torch.allclose(dvol, ref, atol=1e-5)
    again = ops.patch_scatter_add(g, o_dev, ns, roi, torch.zeros_like(vol))
    assert torch.equal(dvol, again)   # deterministic


def test_empty_and_bad_inputs_raise():
    with pytest.raises(RuntimeError):
        ops.cast(torch.empty(0, device=DEV), torch.empty(0, dtype=torch.bfloat16, device=DEV))
    spec = ops.ConvSpec(2, 4, 4, 3, 1, 0)
    with pytest.raises(RuntimeError):
        ops.conv_fprop(spec, torch.zeros(1, 5, 5, 4), torch.zeros(4, 9, 4, device=DEV), None)  # CPU tensor
