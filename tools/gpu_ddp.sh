#!/bin/bash
# N-GPU bench (torchrun), overlapped and blocking gradient exchange.   tools/gpu_ddp.sh <tag> <N> [ab]
TAG=${1:-d}; N=${2:-2}; AB=${3:-}
mkdir -p gpurun_out
run() {  # name, extra env
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_$1.log 2> gpurun_out/bench_${TAG}_$1.err
  echo "$1 exit $?"; cat gpurun_out/bench_${TAG}_$1.log; grep -c "NCCL INFO" gpurun_out/bench_${TAG}_$1.err; grep -m3 "nranks\|NVLS\|Connected all" gpurun_out/bench_${TAG}_$1.err
}
run overlap "X=1"
if [ "$AB" = "ab" ]; then run blocking "MPGAN_NO_COMM_OVERLAP=1"; fi
