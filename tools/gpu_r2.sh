#!/bin/bash
# Round-2 GPU call: new parity tests, the measured parity report, the whole GPU suite, the bench line.
#   tools/gpu_r2.sh <tag> [what...]     what: r2tests parity alltests bench ncu
TAG=${1:-a}; shift
WHAT=${@:-r2tests parity bench}
mkdir -p gpurun_out
nproc > gpurun_out/nproc_$TAG.txt
for w in $WHAT; do
  case $w in
    r2tests) timeout 1500 python -m pytest tests/test_parity_r2_gpu.py -m gpu -q --timeout 900 -x --durations=15 > gpurun_out/pytest_r2_$TAG.log 2>&1
             echo "r2tests exit $?"; tail -40 gpurun_out/pytest_r2_$TAG.log;;
    r2tests_all) timeout 1500 python -m pytest tests/test_parity_r2_gpu.py -m gpu -q --timeout 900 --durations=15 > gpurun_out/pytest_r2_$TAG.log 2>&1
             echo "r2tests exit $?"; tail -60 gpurun_out/pytest_r2_$TAG.log;;
    parity)  timeout 1500 python tools/parity_report.py --full > gpurun_out/parity_$TAG.md 2> gpurun_out/parity_$TAG.err
             echo "parity exit $?"; cat gpurun_out/parity_$TAG.md; tail -5 gpurun_out/parity_$TAG.err;;
    alltests) timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --durations=10 > gpurun_out/pytest_all_$TAG.log 2>&1
             echo "alltests exit $?"; tail -30 gpurun_out/pytest_all_$TAG.log;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
             echo "smoke exit $?"; tail -12 gpurun_out/smoke_$TAG.log;;
    bench)   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
             echo "bench exit $?"; cat gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err;;
    benchq)  timeout 900 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
             echo "bench exit $?"; cat gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err;;
    ref)     timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/benchref_$TAG.log 2> gpurun_out/benchref_$TAG.err
             echo "ref exit $?"; cat gpurun_out/benchref_$TAG.log;;
    phases)  timeout 600 python tools/phase_times.py 10 > gpurun_out/phase_$TAG.log 2>&1; echo "phases exit $?"; cat gpurun_out/phase_$TAG.log;;
    ops)     timeout 900 python tools/bench_ops.py ${OPS_FILTER:-} > gpurun_out/ops_$TAG.log 2>&1; echo "ops exit $?"; tail -70 gpurun_out/ops_$TAG.log;;
    k3d)     timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "tc3" > gpurun_out/pytest_k3d_$TAG.log 2>&1
             echo "k3d exit $?"; tail -30 gpurun_out/pytest_k3d_$TAG.log;;
    n3d)     timeout 900 python -m pytest tests/test_nets_gpu.py -m gpu -q --timeout 600 -k "3" > gpurun_out/pytest_n3d_$TAG.log 2>&1
             echo "n3d exit $?"; tail -30 gpurun_out/pytest_n3d_$TAG.log;;
    b3d)     timeout 900 python tools/bench_3d.py ${B3D_ARGS:-1 3 128} > gpurun_out/bench3d_$TAG.log 2> gpurun_out/bench3d_$TAG.err
             echo "b3d exit $?"; cat gpurun_out/bench3d_$TAG.log; tail -5 gpurun_out/bench3d_$TAG.err;;
    ncud)    timeout 600 python tools/d_convs_once.py > gpurun_out/plain_dconv_$TAG.log 2>&1 &&
             timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"tapgemm|halo3x3|wgrad_kernel" \
                 -o gpurun_out/dconvs_$TAG -f python tools/d_convs_once.py > gpurun_out/ncu_dconv_$TAG.log 2>&1
             echo "ncud exit $?"; ls -la gpurun_out/dconvs_$TAG.ncu-rep;;
    ncu)     CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extra"
             timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2800} -c ${NCU_COUNT:-2800} --csv \
                 --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
             echo "ncu exit $?"; wc -l gpurun_out/launches_$TAG.csv;;
  esac
done
