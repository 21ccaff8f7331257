# round 2, GPU session 6: L2 prefetch of the next item's inputs
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v4.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cat gpurun_out/bench_wfft_v4.json
