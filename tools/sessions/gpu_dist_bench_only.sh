# usage: bash tools/sessions/gpu_dist_bench_only.sh N   (under gpurun --gpus N)
N=${1:-8}
mkdir -p gpurun_out
for cfg in "--size 512" "--size 256 --type r2c" "--size 512 --type r2c"; do
  tag=$(echo $cfg | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N $cfg --no-e2e > gpurun_out/bench_dist_${N}_$tag.json 2> gpurun_out/bench_dist_${N}_$tag.err
  tail -n 1 gpurun_out/bench_dist_${N}_$tag.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["config"]["workload"][:40], round(d["value"],1), "pairs/s", d["roofline"]["stage_ms"], "nvlink", round(d["nvlink"]["frac"],3), round(d["nvlink"]["exchange_ms"],3))'
done
