# round 2, GPU session 34: conflict-free shared-memory layout of the asynchronously fetched inverse-map entries
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > gpurun_out/bench_wfft_v16_$i.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cut -c1-2800 gpurun_out/bench_wfft_v16_$i.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}'
done
