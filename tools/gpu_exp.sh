mkdir -p gpurun_out
V=spfft_b200/lib/variants
echo "=== trace pipe"; SPFFT_B200_LIB=$V/libspfft_b200_trace.so SPFFT_B200_TUNE=9 timeout 200 python tools/xy_trace.py 2>&1 | tail -12
for lag in 3 10; do
echo "=== bench pipe lag $lag"; SPFFT_B200_XY_LAG=$lag SPFFT_B200_XY_RING=$((2*lag+2)) SPFFT_B200_TUNE=9 timeout 300 python bench.py --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['stage_ms'])"
done
SPFFT_B200_TUNE=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_xy_pipe' -s 2 -c 2 -o gpurun_out/prof_pipe python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pipe.log 2>&1
