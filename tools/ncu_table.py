"""Table of the nine discriminator conv launches from the `ncu --set full` report of tools/d_convs_once.py.
usage: python tools/ncu_table.py gpurun_out/dconvs_<tag>.{ncu-rep|csv} > profiles/ncu_r2_dconvs.md
(the .csv is `ncu -i <rep> --page raw --csv`, made on the GPU box: the report itself is too large to travel)"""
import csv
import io
import json
import os
import subprocess
import sys

rep = sys.argv[1]
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, body = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
B = 32
LAYERS = [("D2", 64, 128, 3, 254, 252), ("D3", 128, 256, 4, 252, 125), ("D4", 256, 256, 4, 125, 61)]
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
print("# ncu `--set full` over the nine FLOP-dominant launches (D layers 2-4 x fprop / dgrad / wgrad), batch 32, bf16\n")
print(f"command (one B200, after the plain run of the same command had exited 0): `ncu --set full --clock-control none "
      f"--profile-from-start off -k regex:\"tapgemm|halo3x3|wgrad\" python tools/d_convs_once.py`; the third launch of "
      "each kind is captured (two warm-ups before it, outside the cudaProfilerStart / Stop window).  Times under a profiler are not bench values (`bench.py` / "
      "`tools/bench_ops.py` hold those); the tensor-pipe and DRAM columns are what this table is for.  fprop launches carry the fused "
      "BatchNorm statistics, data gradients do not (as in the training step).\n")
print("| launch | kernel | us (ncu) | algorithmic TFLOP/s | `sm__pipe_tensor_cycles_active` % of peak (elapsed / active) | DRAM read + write MB | "
      "algorithmic MB | `lts__throughput` % | regs | grid |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
out = {}
for li, (name, cin, cout, k, xs, ys) in enumerate(LAYERS):
    flop = 2.0 * B * ys * ys * cout * k * k * cin
    for di, d in enumerate(("fprop", "dgrad", "wgrad")):
        r = body[li * 3 + di]
        us = float(r[ix["gpu__time_duration.sum"]])
        rd = float(r[ix["dram__bytes_read.sum"]])
        wr = float(r[ix["dram__bytes_write.sum"]])
        ru, wu = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
        rd *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3}.get(ru, 1.0)
        wr *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3}.get(wu, 1.0)
        xb, yb, wb = B * xs * xs * cin * 2 / 1e6, B * ys * ys * cout * 2 / 1e6, cout * k * k * cin * 2 / 1e6
        alg = {"fprop": xb + yb + wb, "dgrad": xb + yb + wb, "wgrad": xb + yb + 2 * wb}[d]
        kn = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        te = float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])
        ta = float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
        print(f"| {name} {d} {cin}->{cout} k{k} | `{kn}` | {us:.1f} | {flop / us / 1e6:.0f} | **{te:.1f}** / {ta:.1f} | {rd:.0f} + {wr:.0f} | {alg:.0f} | "
              f"{float(r[ix['lts__throughput.avg.pct_of_peak_sustained_elapsed']]):.0f} | {r[ix['launch__registers_per_thread']]} | {r[ix['launch__grid_size']]} |")
        out[f"{name}_{d}"] = {"us": us, "tensor_pct": te, "dram_mb": rd + wr}
print(f"\nDenominators for the TFLOP/s column: measured bf16 burst {pk['bf16_tflops']} / sustained {pk['bf16_tflops_sustained']} TFLOP/s "
      "(`MEASURED_PEAKS.json`).")
json.dump(out, open("/tmp/ncu_dconvs.json", "w"))
