#!/usr/bin/env python
"""Measured parity of the CUDA path against the CPU oracle (GPU box; oracle = test infrastructure, the checker).

    python tools/parity_report.py [--full] > gpurun_out/parity_r2.md

For 64x64 / batch 2 (the smoke configuration) and -- with --full -- BASELINE.json configs[1] at its full size
(batch 32 of 256x256), in fp32 mode (CUDA-core kernels) and bf16 mode (tcgen05 path): relative L2 of the generator
output (train-mode BatchNorm, 1 UNet and the 6-UNet cascade; eval mode), the discriminator probabilities (teacher-forced
on the oracle's generator output), every loss term, the global and worst-tensor parameter gradients of both optimizer
passes from identical states, and the BatchNorm running buffers -- against the plain fp32 oracle, against the
rounding-matched oracle (oracle/rounding.py; forward quantities only) and beside torch's own CPU autocast(bf16) run of
the oracle as an outside yardstick (small size only: CPU bf16 convolutions are slow).
"""
import argparse
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-modality-minipig-gan_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import mpgan  # noqa: E402
from oracle.gan import GANOracle, synthetic_batch  # noqa: E402
from oracle.nets import CasNetGenerator as OGen  # noqa: E402
from oracle.rounding import bf16_matched  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def grads_of(net):
    return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def grel(a, b):
    num = sum(float((a[k].double().cpu() - b[k].double().cpu()).norm()) ** 2 for k in b)
    den = sum(float(b[k].double().norm()) ** 2 for k in b)
    return (num / den) ** 0.5


def worst(a, b):
    scale = max(float(v.double().norm()) for v in b.values())
    w = max(((float((a[k].double().cpu() - b[k].double().cpu()).norm()) / max(float(b[k].double().norm()), 0.1 * scale), k)
             for k in b))
    return w


def oracle_pass(ora, batch, opt_idx, autocast=False, fp64=False):
    ora = copy.deepcopy(ora)
    if fp64:
        ora = ora.double()
        batch = {k: v.double() for k, v in batch.items()}
    nets = (ora.generator, ora.discriminator)
    for i, net in enumerate(nets):
        for p in net.parameters():
            p.requires_grad_(i == opt_idx)
            p.grad = None
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = ora.training_step(batch, 0, opt_idx)
        loss = loss.float()
    else:
        loss = ora.training_step(batch, 0, opt_idx)
    loss.backward()
    return float(loss), grads_of(nets[opt_idx]), ora


def my_pass(mine, dbatch, opt_idx):
    nets = (mine.generator, mine.discriminator)
    for i, net in enumerate(nets):
        net.runtime.zero_grad()
        for p in net.parameters():
            p.requires_grad_(i == opt_idx)
    loss = mine.training_step(dbatch, 0, opt_idx)
    loss.backward()
    g = grads_of(nets[opt_idx])
    for p in mine.parameters():
        p.requires_grad_(True)
    return float(loss), g


def buffers_rel(mine, ora):
    w = 0.0
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ora.named_buffers()):
        if b2.dtype.is_floating_point:
            w = max(w, rel(b1, b2))
    return w


def row(name, *vals):
    vals = list(vals) + [None] * (6 - len(vals))
    print("| " + name + " | " + " | ".join("-" if v is None else (v if isinstance(v, str) else f"{v:.2e}") for v in vals) + " |",
          flush=True)


def report(S, B, with_autocast):
    batch = synthetic_batch(B, 2, S, seed=1)
    dbatch = {k: v.to(DEV) for k, v in batch.items()}
    torch.manual_seed(0)
    ora = GANOracle("final", dims=2, spatial=S)
    state = copy.deepcopy(ora.state_dict())
    t0 = time.time()
    # ---- oracle forward quantities (train mode, from the initial state)
    o_fwd = copy.deepcopy(ora)
    with torch.no_grad():
        gen_ref = o_fwd.generator(batch["t1w"])
        p_ref = o_fwd.discriminator(gen_ref)
        adv_ref = float(o_fwd.adversarial_loss(p_ref, torch.ones(B, 1)))
        rec_ref = float(o_fwd.reconstruction_loss(gen_ref, batch["t2w"]))
    m_fwd = bf16_matched(ora)
    with torch.no_grad():
        gen_m = m_fwd.generator(batch["t1w"])
        p_m = m_fwd.discriminator(gen_m)
        adv_m = float(m_fwd.adversarial_loss(p_m, torch.ones(B, 1)))
        rec_m = float(m_fwd.reconstruction_loss(gen_m, batch["t2w"]))
        # teacher-forced D: the matched D on the fp32 oracle's generator output
        m_d = bf16_matched(ora.discriminator)
        p_m_tf = m_d(gen_ref)
    # single UNet
    torch.manual_seed(0)
    u_ref = OGen((1, S, S), 1, 2)
    with torch.no_grad():
        u_out = copy.deepcopy(u_ref)(batch["t1w"])
        u_out_m = bf16_matched(u_ref)(batch["t1w"])
    # eval mode (running statistics moved off their init by two forwards)
    e_ref = copy.deepcopy(ora.generator)
    with torch.no_grad():
        for _ in range(2):
            e_ref(synthetic_batch(2, 2, S, seed=7)["t1w"])
    e_state = copy.deepcopy(e_ref.state_dict())
    e_ref.eval()
    with torch.no_grad():
        e_out = e_ref(batch["t1w"])
        e_m = bf16_matched(e_ref)
        e_m.eval()
        e_out_m = e_m(batch["t1w"])
    rl, rg, ora_g = {}, {}, {}
    for idx in (0, 1):
        rl[idx], rg[idx], ora_g[idx] = oracle_pass(ora, batch, idx)
    # the fp32 oracle against its own fp64 evaluation: how well conditioned each quantity is in fp32 at all
    g64 = {}
    for idx in (0, 1):
        _, g64[idx], _ = oracle_pass(ora, batch, idx, fp64=True)
    # discriminator pass with both images teacher-forced (t2 and the oracle's generator output)
    dtf = copy.deepcopy(ora.discriminator)
    for p in dtf.parameters():
        p.grad = None
    ltf = (ora.adversarial_loss(dtf(batch["t2w"]), torch.ones(B, 1) * 0.9) + ora.adversarial_loss(dtf(gen_ref), torch.zeros(B, 1))) / 2
    ltf.backward()
    gtf, ltf = grads_of(dtf), float(ltf)
    al, ag = {}, {}
    if with_autocast:
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            gen_a = copy.deepcopy(ora).generator(batch["t1w"]).float()
            u_a = copy.deepcopy(u_ref)(batch["t1w"]).float()
        for idx in (0, 1):
            al[idx], ag[idx], _ = oracle_pass(ora, batch, idx, autocast=True)
    print(f"\n### {S}x{S}, batch {B}  (oracle side: {time.time() - t0:.0f} s CPU)\n")
    print("| quantity | fp32 mode vs fp32 oracle | bf16 mode vs fp32 oracle | bf16 mode vs rounding-matched oracle | "
          "rounding-matched oracle vs fp32 oracle (the floor of this rounding scheme) | torch CPU autocast(bf16) vs fp32 oracle | "
          "fp32 oracle vs its own fp64 evaluation |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    res = {}
    for prec in ("fp32", "bf16"):
        r = {}
        mine = mpgan.GAN(1, S, S, precision=prec)
        mine.load_state_dict(state)
        with torch.no_grad():
            gen = mine.generator(dbatch["t1w"])
            r["buf_g"] = buffers_rel(mine.generator, o_fwd.generator)
            r["buf_g_m"] = buffers_rel(mine.generator, m_fwd.generator)
            r["gen"], r["gen_m"] = rel(gen, gen_ref), rel(gen, gen_m)
            mine.load_state_dict(state)
            p_tf = mine.discriminator(gen_ref.to(DEV))
            r["p_tf"], r["p_tf_m"] = rel(p_tf, p_ref), rel(p_tf, p_m_tf)
            mine.load_state_dict(state)
            p = mine.discriminator(mine.generator(dbatch["t1w"]))
            r["p"], r["p_m"] = rel(p, p_ref), rel(p, p_m)
            u = mpgan.CasNetGenerator((1, S, S), n_unet_blocks=1, precision=prec)
            u.load_state_dict(u_ref.state_dict())
            uo = u(dbatch["t1w"])
            r["unet"], r["unet_m"] = rel(uo, u_out), rel(uo, u_out_m)
            e = mpgan.CasNetGenerator((1, S, S), precision=prec)
            e.load_state_dict(e_state)
            e.eval()
            eo = e(dbatch["t1w"])
            r["eval"], r["eval_m"] = rel(eo, e_out), rel(eo, e_out_m)
        mine.load_state_dict(state)
        logs = mine.fused_step(dbatch).tolist()
        torch.cuda.synchronize()
        r["adv"], r["adv_m"] = abs(logs[0] - adv_ref) / abs(adv_ref), abs(logs[0] - adv_m) / abs(adv_m)
        r["rec"], r["rec_m"] = abs(logs[1] - rec_ref) / abs(rec_ref), abs(logs[1] - rec_m) / abs(rec_m)
        for idx in (0, 1):
            mine.load_state_dict(state)
            ml, mg = my_pass(mine, dbatch, idx)
            r[f"loss{idx}"] = abs(ml - rl[idx]) / abs(rl[idx])
            r[f"g{idx}"] = grel(mg, rg[idx])
            r[f"w{idx}"] = worst(mg, rg[idx])
            nets_o = (ora_g[idx].generator, ora_g[idx].discriminator)
            r[f"buf{idx}"] = max(buffers_rel(mine.generator, nets_o[0]), buffers_rel(mine.discriminator, nets_o[1]))
        mine.load_state_dict(state)
        D = mine.discriminator
        D.runtime.zero_grad()
        loss = (mine.adversarial_loss(D(dbatch["t2w"]), torch.ones(B, 1, device=DEV) * 0.9)
                + mine.adversarial_loss(D(gen_ref.to(DEV)), torch.zeros(B, 1, device=DEV))) / 2
        loss.backward()
        r["ltf"] = abs(float(loss) - ltf) / abs(ltf)
        r["gtf"] = grel(grads_of(D), gtf)
        res[prec] = r
        del mine
        torch.cuda.empty_cache()
    f, b = res["fp32"], res["bf16"]
    A = with_autocast
    row("G output, 6 UNets + tanh, train-mode BN", f["gen"], b["gen"], b["gen_m"], rel(gen_m, gen_ref), rel(gen_a, gen_ref) if A else None)
    row("G output, 1 UNet + tanh, train-mode BN", f["unet"], b["unet"], b["unet_m"], rel(u_out_m, u_out), rel(u_a, u_out) if A else None)
    row("G output, 6 UNets, eval-mode BN", f["eval"], b["eval"], b["eval_m"], rel(e_out_m, e_out), None)
    row("D probabilities on the oracle's G output (teacher-forced)", f["p_tf"], b["p_tf"], b["p_tf_m"], rel(p_m_tf, p_ref), None)
    row("D probabilities on own G output", f["p"], b["p"], b["p_m"], rel(p_m, p_ref), None)
    row("g_adv loss (relative)", f["adv"], b["adv"], b["adv_m"], abs(adv_m - adv_ref) / abs(adv_ref), None)
    row("g_recon loss (relative)", f["rec"], b["rec"], b["rec_m"], abs(rec_m - rec_ref) / abs(rec_ref), None)
    row("G BatchNorm running buffers after one forward (worst)", f["buf_g"], b["buf_g"], b["buf_g_m"], None, None)
    for idx, name in ((0, "generator pass"), (1, "discriminator pass")):
        row(f"{name}: loss (relative)", f[f"loss{idx}"], b[f"loss{idx}"], None, None,
            abs(al[idx] - rl[idx]) / abs(rl[idx]) if A else None)
        row(f"{name}: global parameter-gradient rel-L2", f[f"g{idx}"], b[f"g{idx}"], None, None, grel(ag[idx], rg[idx]) if A else None,
            grel(rg[idx], g64[idx]))
        row(f"{name}: worst tensor (err / max(|g|, 10 % of largest))", f"{f[f'w{idx}'][0]:.2e} ({f[f'w{idx}'][1]})",
            f"{b[f'w{idx}'][0]:.2e} ({b[f'w{idx}'][1]})", None, None, None)
        row(f"{name}: BN running buffers (worst)", f[f"buf{idx}"], b[f"buf{idx}"], None, None, None)
    row("discriminator pass, both images teacher-forced: loss (relative)", f["ltf"], b["ltf"], None, None, None)
    row("discriminator pass, both images teacher-forced: global parameter-gradient rel-L2", f["gtf"], b["gtf"], None, None, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    print("# Measured parity, round 2 (`python tools/parity_report.py --full` on one B200; oracle on the box's "
          f"{os.cpu_count()} host cores)\n")
    print("Same seeded weights (seed 0) and inputs (seed 1) on both sides.  Relative L2 unless stated.  north_star: "
          "<= 1e-4 fp32 mode, <= 1e-2 bf16 mode.")
    report(64, 2, True)
    if args.full:
        report(256, 32, False)


if __name__ == "__main__":
    main()
