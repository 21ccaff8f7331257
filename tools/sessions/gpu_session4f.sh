mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), d["roofline"]["stage_ms"])'
for lib in default v4; do
  for cfg in "--size 512" "--size 256" "--size 384" "--size 192" "--size 512 --type r2c" "--size 128"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
SPFFT_B200_LIB=$V/libspfft_b200_v4.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "power_of_two or three_times or spherical or full_size or batched" 2>&1 | tail -n 3
tail -n 5 gpurun_out/exp.err
