# round 2, GPU session 30 (2 GPUs): fused xy stage on distributed transforms (y tiles read / write the exchange buffers)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02_dist_check_2gpu.log
grep -c " ok" gpurun_out/r02_dist_check_2gpu.log; grep "FAIL\|DIST_GPU_CHECK\|Error\|exit" gpurun_out/r02_dist_check_2gpu.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline']['stage_ms'], d['parity'])
PY
SPFFT_B200_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/r02_bench_2gpu_nccl.json 2> gpurun_out/bench2n.err; tail -3 gpurun_out/bench2n.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_2gpu_nccl.json').read().strip().splitlines()[-1])
print('nccl', round(d['value'],1), d['roofline']['stage_ms'], d['parity']['ok'])
PY
