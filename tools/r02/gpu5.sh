# round 2, GPU session 5: xor-form addressing + branch-free signs: tests, bench, ncu source-level profile
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -5 gpurun_out/pytest_wfft.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v3.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cat gpurun_out/bench_wfft_v3.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_w -c 4 -o gpurun_out/r02_wfft_v3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > gpurun_out/ncu_wfft.log 2>&1
tail -3 gpurun_out/ncu_wfft.log
