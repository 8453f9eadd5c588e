"""Per-parameter gradient error table (GPU vs CPU oracle) for a single UNet / the generator / discriminator."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-modality-minipig-gan_b200"))
from mpgan import CasNetGenerator  # noqa: E402
from oracle.nets import CasNetGenerator as OGen  # noqa: E402
from oracle.gan import synthetic_batch  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def table(nblocks, dims, size, batch, precision):
    shape = (1,) + (size,) * dims
    torch.manual_seed(0)
    ref = OGen(shape, nblocks, dims)
    mine = CasNetGenerator(shape, n_unet_blocks=nblocks, precision=precision)
    mine.load_state_dict(ref.state_dict())
    x = synthetic_batch(batch, dims, size, seed=1)["t1w"]
    dy = synthetic_batch(batch, dims, size, seed=5)["t2w"]
    y_ref = ref(x)
    y_ref.backward(dy)
    y = mine(x.cuda())
    print(f"== {nblocks} unet(s) dims={dims} size={size} {precision}: fwd rel {rel(y, y_ref):.3e}")
    y.backward(dy.cuda())
    scale = max(float(p.grad.norm()) for p in ref.parameters())
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        e = rel(p1.grad, p2.grad)
        flag = "  <<<<" if e > 1e-3 and float(p2.grad.norm()) > 1e-4 * scale else ""
        print(f"{n1:60s} {tuple(p2.shape)!s:22s} |ref| {float(p2.grad.norm()):.3e} rel {e:.3e}{flag}")


def flagged(nblocks, dims, size, batch, precision):
    import copy
    shape = (1,) + (size,) * dims
    torch.manual_seed(0)
    ref = OGen(shape, nblocks, dims)
    ref64 = copy.deepcopy(ref).double()
    x = synthetic_batch(batch, dims, size, seed=1)["t1w"]
    dy = synthetic_batch(batch, dims, size, seed=5)["t2w"]
    ref(x).backward(dy)
    ref64(x.double()).backward(dy.double())
    runs = []
    for r in range(2):
        mine = CasNetGenerator(shape, n_unet_blocks=nblocks, precision=precision)
        mine.load_state_dict({k: v.float() if v.dtype.is_floating_point else v for k, v in ref.state_dict().items()})
        for b1, b2 in zip(mine.buffers(), ref.buffers()):
            pass
        y = mine(x.cuda())
        y.backward(dy.cuda())
        runs.append({n: p.grad.detach().double().cpu().clone() for n, p in mine.named_parameters()})
    scale = max(float(p.grad.norm()) for p in ref64.parameters())
    print(f"== flagged {nblocks} unets size {size} batch {batch} {precision}; scale {scale:.3e}")
    for (n, p32), (_, p64) in zip(ref.named_parameters(), ref64.named_parameters()):
        t = p64.grad
        norm = max(float(t.norm()), 1e-3 * scale)
        e_m = float((runs[0][n] - t).norm()) / norm
        e_o = float((p32.grad.double() - t).norm()) / norm
        e_rr = float((runs[0][n] - runs[1][n]).norm()) / norm
        if e_m > max(1e-4, 20 * e_o):
            print(f"{n:70s} {tuple(t.shape)!s:18s} |g|/scale {float(t.norm())/scale:.2e} mine {e_m:.2e} o32 {e_o:.2e} run-to-run {e_rr:.2e}")


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    flagged(6, 2, 64, 3, "fp32")
    flagged(1, 2, 64, 3, "fp32")
    flagged(1, 2, 64, 2, "fp32")
