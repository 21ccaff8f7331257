# round 2, GPU session 10: lag / ring sweep of the fused xy kernel (are the B parts polled before their plane is published?)
set -x
mkdir -p gpurun_out
for cfg in "6 14" "8 18" "12 26" "16 34"; do
set -- $cfg
SPFFT_B200_XY_LAG=$1 SPFFT_B200_XY_RING=$2 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > gpurun_out/bench_wfft_v6_lag$1.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "lag $1 ring $2"; cut -c1-2600 gpurun_out/bench_wfft_v6_lag$1.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}'
done
