#!/bin/bash
# Dev tool: the round's evidence batch on one B200 (run under gpurun from the repo root): default bench line, launch lists of the
# bf16 and fp32 modes (graph off so that every launch is listed), `ncu --set full` of the hot kernels.  Outputs under gpurun_out/.
set -u
TAG=${1:-r02z}
OUT=gpurun_out
timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
for PREC in bf16 fp32; do
  PM_CUDA_GRAPH=0 PM_OVERLAP=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv \
    --log-file $OUT/launches_${TAG}_${PREC}.csv python bench.py --precision $PREC --steps 1 --warmup 1 --no-e2e --no-extras \
    --no-gpu-baseline --no-cpu-baseline > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:encoder_fwd_tc<' -s 4 -c 1 -f -o $OUT/prof_fwd_bf16_${TAG} \
  python scripts/enc_timing.py bf16 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc_kernel' -s 6 -c 1 -f -o $OUT/prof_gemm_tc_${TAG} \
  python scripts/enc_timing.py fp32 > /dev/null 2>&1
ls -la $OUT/*${TAG}*
