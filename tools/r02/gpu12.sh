# round 2, GPU session 12: flag duties spread over the warps of a group, relaxed polling, shift decode
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for cfg in "0 0 0" "8 18 0" "8 18 2" "8 18 8" "6 14 4"; do
set -- $cfg
SPFFT_B200_XY_LAG=$1 SPFFT_B200_XY_RING=$2 SPFFT_B200_WGROUP=$3 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v7_$1_$3.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "lag $1 ring $2 wgroup $3"; cut -c1-2800 gpurun_out/bench_wfft_v7_$1_$3.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
done
