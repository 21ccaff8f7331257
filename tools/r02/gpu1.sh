# round 2, GPU session 1: sanity of the round-1 kernels under the new bench.py (parity, GPU reference arm)
# + micro-benchmark of the warp-autonomous FFT
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 120 tools/ubench/wfft_bench 512 10 > gpurun_out/ubench_512.log 2>&1; cat gpurun_out/ubench_512.log
timeout 120 tools/ubench/wfft_bench 8 200 > gpurun_out/ubench_8.log 2>&1; cat gpurun_out/ubench_8.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference-gpu --steps 5 > gpurun_out/bench_refgpu.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_refgpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
ls -la gpurun_out
