#!/bin/bash
# One GPU round: gpu test-suite, graph bench, and (optionally) an ncu launch list of one eager step.
#   tools/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-x}; TESTS=${2:-tests}; NCU=${3:-ncu}
mkdir -p gpurun_out
if [ "$TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest exit $?"; tail -15 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
if [ "$NCU" = "ncu" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2800} -c ${NCU_COUNT:-2800} --csv \
      --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
  echo "ncu exit $?"; wc -l gpurun_out/launches_$TAG.csv
fi
