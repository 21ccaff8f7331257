# round 2, GPU session 16 (2 GPUs): distributed parity (incl. multi-transform, length-512 cases), bench at N = 2
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02_dist_check_2gpu.log
grep -c " ok" gpurun_out/r02_dist_check_2gpu.log; grep "FAIL\|DIST_GPU_CHECK\|Error\|error\|exit" gpurun_out/r02_dist_check_2gpu.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/r02_bench_2gpu.json | cut -c1-3000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --bands 4 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_2gpu_bands4.json 2> gpurun_out/bench2b.err; tail -3 gpurun_out/bench2b.err; cat gpurun_out/r02_bench_2gpu_bands4.json | cut -c1-1500
