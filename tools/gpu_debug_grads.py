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


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    table(1, 2, 32, 2, "fp32")
    table(2, 2, 32, 2, "fp32")
