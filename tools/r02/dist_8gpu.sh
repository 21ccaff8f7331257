# round 2, GPU session 31 (8 GPUs): distributed fused xy stage: parity + the 512^3 bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_8gpu.log 2>&1; echo "exit $?" >> gpurun_out/r02_dist_check_8gpu.log
grep -c " ok" gpurun_out/r02_dist_check_8gpu.log; grep "FAIL\|DIST_GPU_CHECK\|exit" gpurun_out/r02_dist_check_8gpu.log | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_8gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline']['stage_ms'], d['parity']['ok'], d['nvlink']['frac'])
PY
