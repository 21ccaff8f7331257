# round 2, GPU session 2: first run of the warp-FFT kernels (wfft_xy.cu, wfft_z.cu)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -30 gpurun_out/pytest_wfft.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_wfft.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cat gpurun_out/bench_wfft.json
SPFFT_B200_WFFT=2 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_wfft_zonly.json 2>> gpurun_out/bench_wfft.err; cat gpurun_out/bench_wfft_zonly.json
SPFFT_B200_WFFT=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_wfft_xyonly.json 2>> gpurun_out/bench_wfft.err; cat gpurun_out/bench_wfft_xyonly.json
