# round 2, GPU session 15: forward x parts fed by bulk copies issued one part ahead
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for w in 0 4; do
SPFFT_B200_WGROUP=$w timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v9_w$w.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "wgroup $w"; cut -c1-2800 gpurun_out/bench_wfft_v9_w$w.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
done
