# round 2, GPU session 8: deferred publication (tensor-store read wait only at the barrier), release reduction,
# asynchronous inverse-map prefetch; backward W = 4 / forward W = 8
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for w in 0 2 8; do
SPFFT_B200_WGROUP=$w timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v6_w$w.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cut -c1-2600 gpurun_out/bench_wfft_v6_w$w.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
done
