mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "all:", s)'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for lib in default f32v16; do
  for cfg in "--size 192 --precision single" "--size 384 --precision single" "--size 192 --precision single --bands 64" "--size 512 --precision single" "--size 192 --precision single --type r2c"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e --steps 10 2>>gpurun_out/exp.err | python -c "$show"
  done
done
SPFFT_B200_LIB=$V/libspfft_b200_f32v16.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "three_times or single or batched" 2>&1 | tail -n 3
tail -n 5 gpurun_out/exp.err
