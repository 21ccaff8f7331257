set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "three_times or spherical or golden" > gpurun_out/pytest_fast3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fast3.log
tail -5 gpurun_out/pytest_fast3.log
for cfg in "96 c2c double" "192 c2c double" "192 c2c single" "192 r2c double" "384 c2c double" "384 c2c single" "768 c2c single"; do
  set -- $cfg
  timeout 300 python bench.py --size $1 --type $2 --precision $3 --no-cpu-baseline --no-e2e > gpurun_out/bench3_$1_$2_$3.json 2>> gpurun_out/bench3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench3_$1_$2_$3.json"))
print("$cfg", round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), d["roofline"]["stage_ms"])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(z|y|x)_fast3' -s 6 -c 6 -o gpurun_out/prof_fast3_192 python bench.py --size 192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fast3.log 2>&1
ls -la gpurun_out
