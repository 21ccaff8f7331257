# round 2, GPU session 3: warp-FFT kernels after the publication fix: tests + ncu --set full of the four kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -15 gpurun_out/pytest_wfft.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_w -c 4 -o gpurun_out/r02_wfft_v1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > gpurun_out/ncu_wfft.log 2>&1
tail -5 gpurun_out/ncu_wfft.log
ls -la gpurun_out/*.ncu-rep
timeout 300 tools/ubench/wfft_bench2 256 5 > gpurun_out/ubench2.log 2>&1; cat gpurun_out/ubench2.log
