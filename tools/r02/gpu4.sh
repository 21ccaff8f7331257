# round 2, GPU session 4: lane twiddles from shared memory: micro-benchmark of the tile structures + bench
set -x
mkdir -p gpurun_out
timeout 300 tools/ubench/wfft_bench2 256 5 > gpurun_out/ubench2_v2.log 2>&1; cat gpurun_out/ubench2_v2.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v2.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cat gpurun_out/bench_wfft_v2.json
