# round 2, GPU session 21 (8 GPUs): distributed parity with the final code, bench at N = 8 (single transform, 4 bands, 256^3 R2C)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_8gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02_dist_check_8gpu.log
grep -c " ok" gpurun_out/r02_dist_check_8gpu.log; grep "FAIL\|DIST_GPU_CHECK\|exit" gpurun_out/r02_dist_check_8gpu.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err; cut -c1-400 gpurun_out/r02_bench_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --bands 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_8gpu_bands4.json 2> gpurun_out/bench8b.err; tail -2 gpurun_out/bench8b.err; cut -c1-300 gpurun_out/r02_bench_8gpu_bands4.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --size 256 --type r2c --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_8gpu_r2c256.json 2> gpurun_out/bench8c.err; tail -2 gpurun_out/bench8c.err; cut -c1-300 gpurun_out/r02_bench_8gpu_r2c256.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/bench4.err; tail -2 gpurun_out/bench4.err; cut -c1-300 gpurun_out/r02_bench_4gpu.json
