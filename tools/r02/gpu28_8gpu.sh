# round 2, GPU session 28 (8 GPUs): final multi-GPU code (double-buffered exchange targets): parity + bench lines
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_8gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02_dist_check_8gpu.log
grep -c " ok" gpurun_out/r02_dist_check_8gpu.log; grep "FAIL\|DIST_GPU_CHECK\|exit" gpurun_out/r02_dist_check_8gpu.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err; cut -c1-200 gpurun_out/r02_bench_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --size 256 --type r2c --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_8gpu_r2c256.json 2> gpurun_out/bench8c.err; tail -2 gpurun_out/bench8c.err; cut -c1-200 gpurun_out/r02_bench_8gpu_r2c256.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --bands 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_8gpu_bands4.json 2> gpurun_out/bench8b.err; tail -2 gpurun_out/bench8b.err; cut -c1-200 gpurun_out/r02_bench_8gpu_bands4.json
