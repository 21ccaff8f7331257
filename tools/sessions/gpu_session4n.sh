mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "all:", s)'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for cfg in "--size 192 --precision single" "--size 192 --precision single --bands 64" "--size 96 --precision single --bands 128" "--size 192 --precision single --type r2c" "--size 192" "--size 512"; do
  echo "=== default $cfg"
  timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
done
tail -n 5 gpurun_out/exp.err
