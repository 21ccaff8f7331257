# usage: bash tools/sessions/gpu_dist_check_only.sh N   (under gpurun --gpus N)
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist check exit $?" >> gpurun_out/dist_check_$N.log
grep -E "DIST_GPU_CHECK|FAIL|Error|error|exit" gpurun_out/dist_check_$N.log | head -n 20
