mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), d["roofline"]["stage_ms"])'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for cfg in "--size 192" "--size 192 --precision single" "--size 384" "--size 96" "--size 256 --type r2c" "--size 512 --type r2c" "--size 192 --type r2c" "--size 256 --type r2c --precision single" "--size 384 --type r2c"; do
  echo "=== default $cfg"
  timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
done
for lib in f3d256 f3f512; do
  for cfg in "--size 192" "--size 192 --precision single" "--size 96"; do
    echo "=== $lib $cfg"
    SPFFT_B200_LIB=$V/libspfft_b200_$lib.so timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
echo "=== bands"
for cfg in "--size 192 --bands 64" "--size 192 --bands 64 --precision single"; do
  echo "=== bands $cfg"
  timeout 600 python bench.py $cfg --no-cpu-baseline --no-e2e --steps 10 2>>gpurun_out/exp.err | python -c "$show"
done
tail -n 5 gpurun_out/exp.err
