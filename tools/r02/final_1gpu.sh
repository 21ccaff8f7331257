# round 2, GPU session 17: full GPU test suite, the bench lines of the round, ncu launch list + counters (1 GPU)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-3500 gpurun_out/r02_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference-gpu --steps 5 > gpurun_out/r02_bench_reference_gpu.json 2>> gpurun_out/bench.err
SPFFT_B200_WFFT=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/r02_bench_r01kernels.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/r02_bench_r01kernels.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_w -c 4 -o gpurun_out/r02_wfft_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
bash tools/gpu_configs.sh > gpurun_out/r02_configs.log 2>&1; cp gpurun_out/configs.jsonl gpurun_out/r02_configs.jsonl; grep "pairs/s\|===" gpurun_out/r02_configs.log
