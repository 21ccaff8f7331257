# round 2, GPU session 26: claim result first touched a part later; lag sweep with the dynamic queue and the L2 hints
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for cfg in "8 18" "10 22" "12 26" "14 30"; do
set -- $cfg
SPFFT_B200_XY_LAG=$1 SPFFT_B200_XY_RING=$2 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > gpurun_out/bench_wfft_v14_$1.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "lag $1 ring $2"; cut -c1-2800 gpurun_out/bench_wfft_v14_$1.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
done
