mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "x:", s.get("x backward"), s.get("x forward"), "all:", s)'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for lib in default vx1 vx0 vx3; do
  for cfg in "--size 512" "--size 256" "--size 384" "--size 192" "--size 512 --type r2c" "--size 128"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
for lib in default vxf2; do
  for cfg in "--size 512 --precision single" "--size 192 --precision single" "--size 256 --precision single --type r2c"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
tail -n 5 gpurun_out/exp.err
