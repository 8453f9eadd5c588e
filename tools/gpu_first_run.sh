#!/bin/bash
# First-contact GPU run: tcgen05 probes (each group under its own timeout), then the gpu test-suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for g in gemm conv bprop wgrad; do
  timeout 240 python tools/gpu_probe.py $g > gpurun_out/probe_$g.log 2>&1
  echo "probe $g exit $?" >> gpurun_out/probe_status.txt
done
timeout 1500 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k "not tc_" > gpurun_out/pytest_kernels_generic.log 2>&1
echo "kernels generic exit $?" >> gpurun_out/probe_status.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -k "tc_" > gpurun_out/pytest_kernels_tc.log 2>&1
echo "kernels tc exit $?" >> gpurun_out/probe_status.txt
timeout 1500 python -m pytest tests/test_nets_gpu.py -m gpu -q --timeout 600 -k "fp32" > gpurun_out/pytest_nets_fp32.log 2>&1
echo "nets fp32 exit $?" >> gpurun_out/probe_status.txt
cat gpurun_out/probe_status.txt
tail -5 gpurun_out/probe_*.log
tail -30 gpurun_out/pytest_kernels_generic.log
tail -30 gpurun_out/pytest_kernels_tc.log
tail -30 gpurun_out/pytest_nets_fp32.log
