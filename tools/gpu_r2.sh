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
    ktest)   timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "${KFILTER:-s2_halo or tc_conv}" > gpurun_out/pytest_k_$TAG.log 2>&1
             echo "ktest exit $?"; tail -30 gpurun_out/pytest_k_$TAG.log;;
    opsab)   # A/B of the stride-2 halo engines against the tap-GEMM path, training and inference sizes
             for sz in "32 1" "64 2"; do set -- $sz
               echo "== batch $1 scale $2: halo_s2"; OPS_BATCH=$1 OPS_SCALE=$2 timeout 600 python tools/bench_ops.py ${OPS_FILTER:-g_16_32 g_32_64 gT_ g_16_16} 2>&1 | tee -a gpurun_out/opsab_$TAG.log
               echo "== batch $1 scale $2: MPGAN_NO_HALO_S2=1"; MPGAN_NO_HALO_S2=1 OPS_BATCH=$1 OPS_SCALE=$2 timeout 600 python tools/bench_ops.py ${OPS_FILTER:-g_16_32 g_32_64 gT_} 2>&1 | tee -a gpurun_out/opsab_$TAG.log
             done;;
    opsw)    # A/B of the halo weight-gradient kernel
             for e in 0 1; do echo "== MPGAN_NO_WGRAD_HALO=$e"; MPGAN_NO_WGRAD_HALO=$e timeout 600 python tools/bench_ops.py g_16_16 g_16_32 g_32_32 g_32_64 gT_16 2>&1 | grep wgrad | tee -a gpurun_out/opsw_$TAG.log; done;;
    opsc1)   # A/B of the run-based one-channel kernels with vector window loads
             for e in 0 1; do echo "== MPGAN_NO_C1RUN=$e"; MPGAN_NO_C1RUN=$e timeout 600 python tools/bench_ops.py c1_g1 c1_gT c1_11 2>&1 | tee -a gpurun_out/opsc1_$TAG.log; done
             echo "== inference size"; OPS_BATCH=64 OPS_SCALE=2 timeout 600 python tools/bench_ops.py c1_g1 c1_gT c1_11 2>&1 | tee -a gpurun_out/opsc1_$TAG.log;;
    k3d)     timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600 -k "tc3" > gpurun_out/pytest_k3d_$TAG.log 2>&1
             echo "k3d exit $?"; tail -30 gpurun_out/pytest_k3d_$TAG.log;;
    n3d)     timeout 900 python -m pytest tests/test_nets_gpu.py -m gpu -q --timeout 600 -k "3" > gpurun_out/pytest_n3d_$TAG.log 2>&1
             echo "n3d exit $?"; tail -30 gpurun_out/pytest_n3d_$TAG.log;;
    b3d)     timeout 900 python tools/bench_3d.py ${B3D_ARGS:-1 3 128} > gpurun_out/bench3d_$TAG.log 2> gpurun_out/bench3d_$TAG.err
             echo "b3d exit $?"; cat gpurun_out/bench3d_$TAG.log; tail -5 gpurun_out/bench3d_$TAG.err;;
    ncud)    timeout 600 python tools/d_convs_once.py > gpurun_out/plain_dconv_$TAG.log 2>&1 &&
             timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:"tapgemm|halo3x3|wgrad" \
                 -o /tmp/dconvs_$TAG -f python tools/d_convs_once.py > gpurun_out/ncu_dconv_$TAG.log 2>&1
             echo "ncud exit $?"; ls -la /tmp/dconvs_$TAG.ncu-rep
             ncu -i /tmp/dconvs_$TAG.ncu-rep --page raw --csv > gpurun_out/dconvs_$TAG.csv 2> gpurun_out/dconvs_csv_$TAG.err; wc -c gpurun_out/dconvs_$TAG.csv;;
    ncug)    # ncu --set full over the first UNet of one inference forward (batch 64, 512x512): raw + source pages as csv
             timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -c ${NCUG_COUNT:-24} \
                 -o /tmp/ginf_$TAG -f python tools/infer_once.py > gpurun_out/ncu_ginf_$TAG.log 2>&1
             echo "ncug exit $?"; ls -la /tmp/ginf_$TAG.ncu-rep
             ncu -i /tmp/ginf_$TAG.ncu-rep --page raw --csv > gpurun_out/ginf_raw_$TAG.csv 2> /dev/null
             ncu -i /tmp/ginf_$TAG.ncu-rep --page source --csv 2> /dev/null | gzip > gpurun_out/ginf_source_$TAG.csv.gz
             ls -la gpurun_out/ginf_*;;
    ncugp)   # warm-cache launch lists of the generator's forward and backward (one eager pass each)
             for ph in fwd bwd; do
               timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --profile-from-start off --csv \
                   --log-file gpurun_out/launches_g${ph}_$TAG.csv python tools/g_phases_once.py $ph > gpurun_out/ncu_g${ph}_$TAG.log 2>&1
               echo "ncugp $ph exit $?"; wc -l gpurun_out/launches_g${ph}_$TAG.csv
             done;;
    ncugt)   # ncu --set full (warm caches) over the first UNet of the generator forward and the last UNet's backward, training size
             timeout 900 ncu --set full --cache-control none --clock-control none --profile-from-start off -c 32 \
                 -o /tmp/gtf_$TAG -f python tools/g_phases_once.py fwd > gpurun_out/ncu_gtf_$TAG.log 2>&1
             timeout 1200 ncu --set full --cache-control none --clock-control none --profile-from-start off -c 76 \
                 -o /tmp/gtb_$TAG -f python tools/g_phases_once.py bwd > gpurun_out/ncu_gtb_$TAG.log 2>&1
             ncu -i /tmp/gtf_$TAG.ncu-rep --page raw --csv > gpurun_out/gtrain_fwd_raw_$TAG.csv 2> /dev/null
             ncu -i /tmp/gtb_$TAG.ncu-rep --page raw --csv > gpurun_out/gtrain_bwd_raw_$TAG.csv 2> /dev/null
             ls -la gpurun_out/gtrain_*;;
    ncui)    CMD="python tools/infer_once.py"
             timeout 600 $CMD > gpurun_out/plain_infer_$TAG.log 2>&1 &&
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
                 --log-file gpurun_out/launches_infer_$TAG.csv $CMD > gpurun_out/ncu_infer_$TAG.log 2>&1
             echo "ncui exit $?"; wc -l gpurun_out/launches_infer_$TAG.csv;;
    ncu)     CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extra"
             timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2800} -c ${NCU_COUNT:-2800} --csv \
                 --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
             echo "ncu exit $?"; wc -l gpurun_out/launches_$TAG.csv;;
  esac
done
