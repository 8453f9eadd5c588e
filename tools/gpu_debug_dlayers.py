"""Layer-by-layer comparison of the discriminator (fp32 mode) against an fp64 oracle: forward tensors and gradients."""
import copy
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-modality-minipig-gan_b200"))
from mpgan import Discriminator, ops  # noqa: E402
from mpgan.nets import conv_backward, bn_act_backward  # noqa: E402
from mpgan._lib import ACT_LEAKY  # noqa: E402
from oracle.nets import Discriminator as ODis  # noqa: E402
from oracle.gan import synthetic_batch  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def to_nchw(t):
    return t.float().permute(0, 3, 1, 2)


size, batch = 64, 3
torch.manual_seed(0)
ref = ODis((1, size, size), dims=2, spatial=size)
ref64 = copy.deepcopy(ref).double()
mine = Discriminator((1, size, size), spatial=size, precision="fp32")
mine.load_state_dict(ref.state_dict())
x = synthetic_batch(batch, 2, size, seed=1)["t1w"]
dp = torch.linspace(-1, 1, batch).reshape(batch, 1)

# oracle fp64 with per-module outputs and their gradients
outs, grads = {}, {}
h = x.double()
for i, m in enumerate(ref64.model_conv):
    if isinstance(m, nn.LeakyReLU):
        h = nn.functional.leaky_relu(h, 0.2)       # out-of-place so every tensor can be inspected
    else:
        h = m(h)
    h.retain_grad()
    outs[i] = h
p64 = ref64.model_linear(h)
p64.backward(dp.double())

p, plan = mine.run_forward(x.cuda(), save=True)
print("prob rel", rel(p, p64))
lin_tape, pp, feat_shape = plan.tape[-1]
for li in range(4):
    h_in, c, saved, trained, _ = plan.tape[li]
    print(f"layer {li}: conv out rel {rel(to_nchw(c), outs[3 * li]):.3e}")
    mean, invstd, scale, shift = saved
    bn = ref64.model_conv[3 * li + 1]
    # batch statistics
    co = outs[3 * li]
    m64 = co.mean(dim=(0, 2, 3)); v64 = co.var(dim=(0, 2, 3), unbiased=False)
    print(f"          mean rel {rel(mean, m64):.3e} invstd rel {rel(invstd, 1 / torch.sqrt(v64 + 1e-5)):.3e}")
# manual backward mirroring run_backward, capturing per-layer gradients
rt = plan.rt
plan.tape.pop()
lin = mine.model_linear[1]
dz = ops.sigmoid_bwd(dp.cuda().float().contiguous(), pp, torch.empty_like(pp))
z_in, wcl, first = lin_tape[0]
dx = torch.empty_like(z_in)
ops.linear_bwd(z_in, wcl, dz, dx, None, None)
dh = dx.reshape(feat_shape)
convs = [m for m in mine.model_conv if isinstance(m, nn.Conv2d)]
bns = [m for m in mine.model_conv if isinstance(m, nn.BatchNorm2d)]
for i in range(3, -1, -1):
    h_in, c, saved, trained, _ = plan.tape.pop()
    print(f"layer {i}: d(act out) rel {rel(to_nchw(dh), outs[3 * i + 2].grad):.3e}")
    dc = bn_act_backward(dh, c, saved, bns[i], ACT_LEAKY, None, 0.2, plan, trained)
    print(f"          d(conv out) rel {rel(to_nchw(dc), outs[3 * i].grad):.3e}")
    rec = rt.rec[convs[i]]
    rec.dw.zero_()
    dh = conv_backward(rec, h_in, dc, plan, need_dx=(i > 0))
    g64 = ref64.model_conv[3 * i].weight.grad
    print(f"          dW rel {rel(convs[i].weight.grad, g64):.3e}   dgamma rel {rel(bns[i].weight.grad, ref64.model_conv[3 * i + 1].weight.grad):.3e}")
    # isolate the kernels: feed the ORACLE's tensors through my wgrad / BN-backward
    if i > 0:
        xin = outs[3 * i - 1].detach().float().permute(0, 2, 3, 1).contiguous().cuda()
    else:
        xin = x.permute(0, 2, 3, 1).contiguous().cuda()
    dy_or = outs[3 * i].grad.float().permute(0, 2, 3, 1).contiguous().cuda()
    dw = torch.zeros_like(rec.dw)
    ops.conv_wgrad(rec.spec, xin, dy_or, dw)
    want = g64.permute(0, 2, 3, 1).reshape(-1)
    print(f"          wgrad kernel on oracle tensors: rel {rel(dw, want):.3e}")
    dact_or = outs[3 * i + 2].grad.float().permute(0, 2, 3, 1).contiguous().cuda()
    c_or = outs[3 * i].detach().float().permute(0, 2, 3, 1).contiguous().cuda()
    dc2 = bn_act_backward(dact_or, c_or, saved, bns[i], ACT_LEAKY, None, 0.2, plan, trained)
    print(f"          BN+LeakyReLU backward kernel on oracle tensors: rel {rel(to_nchw(dc2), outs[3 * i].grad):.3e}")
