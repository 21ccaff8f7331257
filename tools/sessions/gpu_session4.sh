set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
for cfg in "128 c2c double" "192 c2c double" "192 c2c single" "256 r2c double" "256 c2c double" "384 c2c double" "512 r2c double"; do
  set -- $cfg
  timeout 300 python bench.py --size $1 --type $2 --precision $3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1_$2_$3.json 2>> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1_$2_$3.json"))
print("$cfg", round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), d["roofline"]["stage_ms"])
PY
done
ls -la gpurun_out
