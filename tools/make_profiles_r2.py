"""Regenerate the round-2 profile documents from the files one GPU call left in gpurun_out/ (tools/gpu_r2.sh <tag> bench
phases ncu ncugp ncui ncud ops).  usage: python tools/make_profiles_r2.py <tag> [phase-tag]
"Before" launch lists (round-2 session start) are the committed profiles/raw_r2/*_before.csv.gz."""
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_launches  # noqa: E402

tag = sys.argv[1]
ptag = sys.argv[2] if len(sys.argv) > 2 else tag
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def summary(path, title="x"):
    """-> (source line, table rows) of a launch list (csv or csv.gz)"""
    if path.endswith(".gz"):
        tmp = "/tmp/_ll.csv"
        open(tmp, "w").write(gzip.open(path, "rt").read())
        path_show, path = os.path.relpath(path, ROOT), tmp
    else:
        path_show = os.path.relpath(path, ROOT)
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    try:
        summarize_launches.main(path, None, title)
    finally:
        sys.stdout = old
    t = buf.getvalue().split("\n")
    src = [l for l in t if l.startswith("source:")][0]
    src = re.sub(r"`[^`]*`", f"`{path_show}`", src, count=1)
    src = re.sub(r"; ncu serialises.*?\)", ")", src)
    return src, "\n".join(l for l in t if l.startswith("|"))


def two(before, after, title, intro, out):
    sb, rb = summary(before)
    sa, ra = summary(after)
    open(os.path.join(PR, out), "w").write(
        f"# {title}\n\n{intro}\n\n## After (final round-2 build)\n\n{sa}\n\n{ra}\n\n"
        f"## Before (round-2 session start: round-1 kernels + rank-3 path)\n\n{sb}\n\n{rb}\n")


# ---- bench line, phases
bench = json.loads(open(os.path.join(GO, f"bench_{tag}.log")).read().strip().splitlines()[-1])
shutil.copy(os.path.join(GO, f"bench_{tag}.log"), os.path.join(PR, "bench_r2.json"))
ph = open(os.path.join(GO, f"phase_{ptag}.log")).read()
pv = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"^(.+?)\s+x\d:\s+([\d.]+) ms", ph, re.M)}
tot = float(re.search(r"sum over the step: ([\d.]+)", ph).group(1))
rows = [("G forward (tape)", "x1", "2.10", "2.10", "G fwd (tape)"), ("D forward (data-gradient-only plan)", "x1", "1.40-1.45", "1.43", "D fwd (dgrad-only plan)"),
        ("D backward, data gradient only", "x1", "1.86", "1.82", "D bwd dx only"), ("G backward", "x1", "4.66", "4.85", "G bwd"),
        ("G Adam + zero", "x1", "0.03", "0.03", "G adam+zero"), ("D forward (tape)", "x2", "1.40-1.45", "1.43", "D fwd (tape)"),
        ("G forward (no tape)", "x1", "2.08", "2.09", "G fwd (no tape)"), ("D backward, weight + data gradients", "x2", "2.73-2.75", "2.65", "D bwd wgrad only"),
        ("D Adam + zero", "x1", "0.03", "0.03", "D adam+zero")]
r = bench["roofline"]
doc = ["# Where one training step goes, round 2 (each phase of `GAN.fused_step` replayed from its own CUDA graph)", "",
       "`python tools/phase_times.py 10` on one B200 (batch 32, 256x256, bf16; CUDA events; ms).", "",
       "| phase | per step | round 1 final | round-2 session start | now |", "|---|---:|---:|---:|---:|"]
for name, mult, r1, r2s, key in rows:
    doc.append(f"| {name} | {mult} | {r1} | {r2s} | {pv[key]:.2f} |")
doc += [f"| **sum over the step** | | **20.5** | **20.5** | **{tot:.1f}** |", "", "Raw output of the final build:", "", "```", ph.rstrip(), "```", "",
        f"`bench.py` (whole step in one graph, driver contract): **{bench['value']:.0f} image-pairs/s** ({bench['ms_per_step']:.2f} ms/step; round 1: 1 559), "
        f"end to end from pinned host buffers {bench['e2e']['value']:.0f}; roofline kernel (D layer 3 forward) {r['kernel_ms'] * 1e3:.0f} us in situ = "
        f"{r['achieved']:.0f} TFLOP/s ({r['frac_burst']:.2f} of the measured burst peak, {r['frac_sustained']:.2f} of the sustained one; ncu: tensor-pipe active and DRAM "
        f"traffic per launch in `profiles/ncu_r2_dconvs.md`).  Step level: {bench['config']['step_tflops']:.0f} TFLOP/s = "
        f"{r['step_frac_of_sustained_compute_roofline']:.3f} of the sustained compute roofline."]
if "extra" in bench:
    e3, e5 = bench["extra"]["cfg3_perceptual_train"], bench["extra"]["cfg5_inference"]
    doc.append(f"Extra legs of the same run: BASELINE configs[2] (perceptual / patch-discriminator step) {e3['value']:.0f} pairs/s (round 1: 654), configs[4] "
               f"(inference, 512x512) {e5['value']:.0f} slices/s = {e5['frac_of_hbm_roofline']:.2f} of the HBM roofline (round 1: 5 839 = 0.25).")
if "cpu_baseline" in bench:
    c = bench["cpu_baseline"]
    doc.append(f"CPU arm of the same run: {c['value']:.2f} pairs/s on {c['cores']} cores (kind `{c['kind']}`).")
open(os.path.join(PR, "phase_times_r2.md"), "w").write("\n".join(doc) + "\n")

# ---- launch lists
summarize_launches.main(os.path.join(GO, f"launches_{tag}.csv"), os.path.join(PR, "launches_r2_train.md"),
                        "Round 2 final build: two eager training steps (batch 32, 256x256, bf16), `ncu --metrics gpu__time_duration.sum "
                        "--clock-control none -s 2800 -c 2800` (cold caches, serialised: read shares)")
with open(os.path.join(GO, f"launches_{tag}.csv"), "rb") as f, gzip.open(os.path.join(PR, "launches_r2_train.csv.gz"), "wb") as g:
    g.write(f.read())
for ph_ in ("gfwd", "gbwd", "infer"):
    with open(os.path.join(GO, f"launches_{ph_}_{tag}.csv"), "rb") as f, gzip.open(os.path.join(PR, "raw_r2", f"launches_{ph_}_after.csv.gz"), "wb") as g:
        g.write(f.read())
two(os.path.join(PR, "raw_r2", "launches_gfwd_before.csv.gz"), os.path.join(PR, "raw_r2", "launches_gfwd_after.csv.gz"),
    "Generator forward (tape), one eager pass, warm caches",
    "`ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --profile-from-start off --csv python tools/g_phases_once.py fwd` "
    f"(batch 32, 256x256, bf16; the second pass is recorded).  Launches are serialised by ncu, so the sum exceeds the wall time of the CUDA-graph replay "
    f"({pv['G fwd (tape)']:.2f} ms, `profiles/phase_times_r2.md`: the residual-branch convolutions run beside the main chain).", "launches_r2_gfwd.md")
two(os.path.join(PR, "raw_r2", "launches_gbwd_before.csv.gz"), os.path.join(PR, "raw_r2", "launches_gbwd_after.csv.gz"),
    "Generator backward, one eager pass, warm caches",
    "`ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --profile-from-start off --csv python tools/g_phases_once.py bwd` "
    f"(batch 32, 256x256, bf16).  Serialised sum vs {pv['G bwd']:.2f} ms wall (the weight gradients run on their own stream).", "launches_r2_gbwd.md")
inf = f"{bench['extra']['cfg5_inference']['value']:.0f}" if "extra" in bench else "?"
two(os.path.join(PR, "raw_r2", "launches_infer_before.csv.gz"), os.path.join(PR, "raw_r2", "launches_infer_after.csv.gz"),
    "Generator inference, one eager forward (BASELINE configs[4]: batch 64 of 512x512, bf16, eval-mode BatchNorm)",
    "`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/infer_once.py` (third forward; cold caches -- "
    f"the tensors are 134 MB each, so that is also the in-situ regime).  Bench: 5 707 -> {inf} slices/s (`profiles/bench_r2.json` extra.cfg5_inference).",
    "launches_r2_infer.md")

# ---- the nine D conv launches under ncu --set full
dcsv = os.path.join(GO, f"dconvs_{tag}.csv")
if os.path.exists(dcsv):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_table.py"), dcsv], capture_output=True, text=True).stdout
    open(os.path.join(PR, "ncu_r2_dconvs.md"), "w").write(out.replace(f"gpurun_out/dconvs_{tag}", "profiles/ncu_r2_dconvs_raw"))
    shutil.copy(dcsv, os.path.join(PR, "ncu_r2_dconvs_raw.csv"))
    rws = list(csv.reader(open(dcsv)))
    hdr, units, body = rws[0], rws[1], rws[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    rr = body[3]
    assert "tapgemm_kernel<256" in rr[ix["Kernel Name"]].replace("(int)", "")
    mult = lambda u: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(u, 1.0)  # noqa: E731
    rd = float(rr[ix["dram__bytes_read.sum"]]) * mult(units[ix["dram__bytes_read.sum"]])
    wr = float(rr[ix["dram__bytes_write.sum"]]) * mult(units[ix["dram__bytes_write.sum"]])
    json.dump({"kernel": "tc::tapgemm_kernel<256, 64, 1, 0> (D layer 3 forward, 128 -> 256, k4 s2, batch 32)",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "tensor_pipe_pct_of_peak_elapsed": float(rr[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
               "source": "profiles/ncu_r2_dconvs.md (ncu --set full --clock-control none --profile-from-start off over tools/d_convs_once.py, round 2 "
                         "final build; raw page: profiles/ncu_r2_dconvs_raw.csv)"},
              open(os.path.join(PR, "roofline_kernel_traffic.json"), "w"), indent=1)

# ---- per-op table
ops = os.path.join(GO, f"ops_{tag}.log")
if os.path.exists(ops):
    shutil.copy(ops, os.path.join(PR, "ops_r2.txt"))
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    rows_ = []
    for l in open(ops):
        m = re.match(r"(.+?)\s+([\d.]+) us\s+([\d.]+) GB/s\s+([\d.]+) TFLOP/s", l)
        if m:
            rows_.append((m.group(1).strip(), float(m.group(2)), float(m.group(3)), float(m.group(4))))
    out = ["# Per-layer convolution / BatchNorm rooflines, round 2 final build: `python tools/bench_ops.py`, batch 32, bf16, CUDA-graph replay of 20 "
           "back-to-back launches", "",
           f"Denominators: measured burst bf16 {pk['bf16_tflops']} TFLOP/s and HBM copy {pk['hbm_gbs']} GB/s (`MEASURED_PEAKS.json`).  FLOPs and bytes are "
           "algorithmic (2 x pixels x Cout x taps x Cin; inputs read once + outputs written once).  The generator's 4-17 MB tensors are L2-resident between "
           "launches, so their GB/s are not HBM figures; twenty back-to-back launches of the discriminator layers run under `sw_power_cap`, i.e. these are "
           "SUSTAINED figures (a single warm launch under ncu is 10-20 % faster: `profiles/ncu_r2_dconvs.md`).  `3D` rows: the reference's literal 128^3 "
           "volumes, batch 1.", "", "| layer / direction | us | TFLOP/s | of tensor peak | GB/s (algorithmic) | of HBM peak |", "|---|---:|---:|---:|---:|---:|"]
    for n, us, gb, tf in rows_:
        out.append(f"| {n} | {us:.1f} | {tf:.1f} | {100 * tf / pk['bf16_tflops']:.0f} % | {gb:.0f} | {100 * gb / pk['hbm_gbs']:.0f} % |")
    open(os.path.join(PR, "conv_rooflines_r2.md"), "w").write("\n".join(out) + "\n")
print("profiles refreshed from tag", tag)
