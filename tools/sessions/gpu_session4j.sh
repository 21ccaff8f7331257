mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "all:", s)'
for lib in default memS memCG default; do
  for cfg in "--size 512" "--size 256"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
tail -n 5 gpurun_out/exp.err
