mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "all:", s)'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for cfg in "--size 512" "--size 256" "--size 384" "--size 192" "--size 512 --type r2c" "--size 256 --type r2c" "--size 128" "--size 512 --precision single" "--size 192 --precision single" "--size 1024 --precision single" "--size 768 --precision single"; do
  echo "=== default $cfg"
  timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
done
echo "=== fused xy with 4-lane tiles"
SPFFT_B200_LIB=$V/libspfft_b200_v4.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_xy" 2>&1 | tail -n 3
for tune in 1 5 9; do
  for cfg in "--size 512" "--size 256"; do
  echo "=== v4 TUNE=$tune $cfg"
  SPFFT_B200_TUNE=$tune SPFFT_B200_LIB=$V/libspfft_b200_v4.so timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
tail -n 5 gpurun_out/exp.err
