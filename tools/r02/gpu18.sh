# round 2, GPU session 18: L2 eviction hints (streaming data evict-first, hand-off ring evict-last): time and DRAM bytes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for cfg in "8 18" "10 22" "6 14"; do
set -- $cfg
SPFFT_B200_XY_LAG=$1 SPFFT_B200_XY_RING=$2 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > gpurun_out/bench_wfft_v10_$1.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "lag $1 ring $2"; cut -c1-2800 gpurun_out/bench_wfft_v10_$1.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}'
SPFFT_B200_XY_LAG=$1 SPFFT_B200_XY_RING=$2 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_wxy -c 2 --csv --log-file gpurun_out/ncu_traffic_lag$1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > /dev/null 2>&1
grep -o '"k_wxy_[a-z]*\|dram__bytes[a-z_.]*","[A-Za-z]*","[0-9.,]*\|gpu__time[a-z_.]*","[a-z]*","[0-9.,]*' gpurun_out/ncu_traffic_lag$1.csv | tr '\n' ' '; echo
done
