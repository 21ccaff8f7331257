# usage: bash tools/gpu_dist.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist check exit $?" >> gpurun_out/dist_check_$N.log
grep -E "DIST_GPU_CHECK|FAIL|Error|error|exit" gpurun_out/dist_check_$N.log | head -20
for mode in 1 0; do
SPFFT_B200_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-e2e > gpurun_out/bench_dist_${N}_p2p$mode.json 2> gpurun_out/bench_dist_${N}_p2p$mode.err
tail -1 gpurun_out/bench_dist_${N}_p2p$mode.json
done
