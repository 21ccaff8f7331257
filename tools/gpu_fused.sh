set -x
mkdir -p gpurun_out
for lag in default 3 6; do
  for ring in default; do
    if [ $lag = default ]; then unset SPFFT_B200_XY_LAG; else export SPFFT_B200_XY_LAG=$lag; export SPFFT_B200_XY_RING=$((2*lag+2)); fi
    SPFFT_B200_TUNE=5 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_fused_$lag.json 2>> gpurun_out/bench_fused.err
    cat gpurun_out/bench_fused_$lag.json
  done
done
unset SPFFT_B200_XY_LAG SPFFT_B200_XY_RING
SPFFT_B200_TUNE=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_xy_' -s 2 -c 2 -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log
