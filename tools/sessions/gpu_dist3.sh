# usage: bash tools/sessions/gpu_dist3.sh N  (under gpurun --gpus N): parity incl. float wire format + benches
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist check exit $?" >> gpurun_out/dist_check_$N.log
grep -E "DIST_GPU_CHECK|FAIL|Error|error|exit" gpurun_out/dist_check_$N.log | head -n 20
grep -c "f32-wire" gpurun_out/dist_check_$N.log
for cfg in "--size 512" "--size 512 --exchange float"; do
  tag=$(echo $cfg | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N $cfg --no-e2e > gpurun_out/bench_dist_${N}_$tag.json 2> gpurun_out/bench_dist_${N}_$tag.err
  tail -n 1 gpurun_out/bench_dist_${N}_$tag.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["config"].get("exchange"), round(d["value"],1), "pairs/s", d["roofline"]["stage_ms"], "nvlink", round(d["nvlink"]["frac"],3), round(d["nvlink"]["exchange_ms"],3))'
done
