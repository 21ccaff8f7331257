mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "all:", s)'
SPFFT_B200_LIB=$V/libspfft_b200_bulk.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "power_of_two or spherical or full_size or golden or reference_shapes" 2>&1 | tail -n 3
for lib in default bulk default bulk; do
  for cfg in "--size 512" "--size 256" "--size 512 --type r2c"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
tail -n 5 gpurun_out/exp.err
