"""Diagnostic probe for the first GPU runs: exercises each tcgen05 configuration in isolation and prints error
metrics instead of asserting, so one gpurun call tells which descriptor/layout variants are right.
Writes gpurun_out/probe.json."""
import json
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-modality-minipig-gan_b200"))
from mpgan import ops  # noqa: E402

DEV = "cuda"
out = {}


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(shape, generator=g) * 2 - 1).to(DEV)


def cl(x):
    return x.permute(0, 2, 3, 1).contiguous()


def uncl(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def oti(w):
    return w.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def run(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        out[name] = r
        print(name, r, flush=True)
    except Exception as e:  # noqa: BLE001
        out[name] = "EXC " + repr(e)[:300]
        print(name, "EXC", repr(e)[:300], flush=True)
        traceback.print_exc()


def fprop_case(n, cin, cout, h, w_, k, s, p):
    def f():
        x = rnd(n, cin, h, w_, seed=11).bfloat16()
        w = (rnd(cout, cin, k, k, seed=12) * 0.1).bfloat16()
        ref = F.conv2d(x.float(), w.float(), None, stride=s, padding=p)
        spec = ops.ConvSpec(2, cin, cout, k, s, p)
        y, fused = ops.conv_fprop(spec, cl(x), oti(w), None, stats=None)
        torch.cuda.synchronize()
        return {"rel": rel(uncl(y), ref), "nan": bool(torch.isnan(y.float()).any())}
    return f


def bprop_case(n, cin, cout, h, w_, k, s, p):
    def f():
        oh, ow = (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1
        dy = rnd(n, cout, oh, ow, seed=14).bfloat16()
        w = (rnd(cout, cin, k, k, seed=15) * 0.1).bfloat16()
        ref = torch.nn.grad.conv2d_input((n, cin, h, w_), w.float(), dy.float(), stride=s, padding=p)
        spec = ops.ConvSpec(2, cin, cout, k, s, p)
        wt = torch.empty(w.numel(), dtype=torch.bfloat16, device=DEV)
        ops.weight_transpose(oti(w), wt, cout, k * k, cin)
        dx, _ = ops.conv_bprop(spec, cl(dy), oti(w), wt, None, xs=(h, w_))
        torch.cuda.synchronize()
        return {"rel": rel(uncl(dx), ref)}
    return f


def wgrad_case(n, cin, cout, h, w_, k, s, p):
    def f():
        oh, ow = (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1
        x = rnd(n, cin, h, w_, seed=16).bfloat16()
        dy = rnd(n, cout, oh, ow, seed=17).bfloat16()
        ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, k, k), dy.float(), stride=s, padding=p)
        spec = ops.ConvSpec(2, cin, cout, k, s, p)
        dw = torch.zeros(cout, k * k, cin, device=DEV)
        ops.conv_wgrad(spec, cl(x), cl(dy), dw)
        torch.cuda.synchronize()
        return {"rel": rel(dw, oti(ref))}
    return f


GROUPS = {
    "gemm": [("gemm_sw128_N128", fprop_case(1, 64, 128, 16, 16, 1, 1, 0)),
             ("gemm_sw128_K256_N256", fprop_case(2, 256, 256, 16, 16, 1, 1, 0)),
             ("gemm_sw64_N64", fprop_case(1, 32, 64, 16, 16, 1, 1, 0)),
             ("gemm_sw32_N16", fprop_case(1, 16, 16, 16, 16, 1, 1, 0)),
             ("gemm_sw32_N32", fprop_case(1, 16, 32, 16, 16, 1, 1, 0))],
    "conv": [("conv3_s1_p1", fprop_case(2, 64, 64, 16, 16, 3, 1, 1)),
             ("conv3_s1_valid_odd", fprop_case(2, 64, 128, 30, 30, 3, 1, 0)),
             ("conv4_s2", fprop_case(2, 128, 256, 30, 30, 4, 2, 0)),
             ("conv4_s2_odd", fprop_case(3, 256, 256, 29, 29, 4, 2, 0)),
             ("conv3_s2_p1_c16", fprop_case(2, 16, 32, 32, 32, 3, 2, 1))],
    "bprop": [("bprop_s1", bprop_case(2, 64, 128, 30, 30, 3, 1, 0)),
              ("bprop_s2_k4", bprop_case(2, 128, 256, 30, 30, 4, 2, 0)),
              ("bprop_s2_k3_p1", bprop_case(2, 16, 32, 32, 32, 3, 2, 1))],
    "wgrad": [("wgrad_gemm_64", wgrad_case(1, 64, 128, 16, 16, 1, 1, 0)),
              ("wgrad_gemm_256", wgrad_case(2, 256, 256, 16, 16, 1, 1, 0)),
              ("wgrad_k3", wgrad_case(2, 64, 128, 30, 30, 3, 1, 0)),
              ("wgrad_k4s2", wgrad_case(2, 128, 256, 30, 30, 4, 2, 0))],
}

if __name__ == "__main__":
    group = sys.argv[1]
    print(group, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    for name, fn in GROUPS[group]:
        run(name, fn)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"probe_{group}.json"), "w"), indent=1)
