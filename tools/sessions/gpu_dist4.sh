# under gpurun --gpus N: parity + forward tile order A/B
N=${1:-2}
mkdir -p gpurun_out
DIST_CHECK_MODES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist check exit $?" >> gpurun_out/dist_check_$N.log
grep -E "DIST_GPU_CHECK|FAIL|Error|error|exit" gpurun_out/dist_check_$N.log | head -n 20
for ord in 1 0 1 0; do
  SPFFT_B200_FWD_ORDER=$ord timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --size 512 --no-e2e > gpurun_out/bench_dist_${N}_ord$ord.json 2> gpurun_out/bench_dist_${N}_ord$ord.err
  tail -n 1 gpurun_out/bench_dist_${N}_ord$ord.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("order", '$ord', round(d["value"],1), "pairs/s", d["roofline"]["stage_ms"])'
done
