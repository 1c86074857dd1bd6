set -u
OUT=gpurun_out; TAG=r02zz
for PREC in bf16 fp32; do
  PM_CUDA_GRAPH=0 PM_OVERLAP=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv \
    --log-file $OUT/launches_${TAG}_${PREC}.csv python bench.py --precision $PREC --steps 1 --warmup 1 --no-e2e --no-extras \
    --no-gpu-baseline --no-cpu-baseline > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k gemm_tc_kernel -s 6 -c 1 -f -o $OUT/prof_gemm_tc_${TAG} \
  python scripts/enc_timing.py fp32 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k open_drawer_post_kernel -s 3 -c 1 -f -o $OUT/prof_env_post_${TAG} \
  python scripts/env_step_timing.py 4096 > /dev/null 2>&1
ls -la $OUT/*${TAG}*
