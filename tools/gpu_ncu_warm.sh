#!/bin/bash
# launch list of one eager step with caches NOT flushed between kernels (closer to the in-graph behaviour of the
# L2-resident generator tensors than the default cold-cache list)
TAG=${1:-w}
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plainw_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s ${NCU_SKIP:-1100} -c ${NCU_COUNT:-1100} --csv \
    --log-file gpurun_out/launchesw_$TAG.csv $CMD > gpurun_out/ncuw_$TAG.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launchesw_$TAG.csv
