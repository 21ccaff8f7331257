mkdir -p gpurun_out
V=spfft_b200/lib/variants
show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), d["roofline"]["stage_ms"], "e2e", d.get("e2e"))'
for lib in default x3m1 x3m2 f3r168; do
  for cfg in "--size 192" "--size 192 --precision single" "--size 384"; do
    echo "=== $lib $cfg"
    if [ $lib = default ]; then L=""; else L="$V/libspfft_b200_$lib.so"; fi
    SPFFT_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
echo "=== correctness of the variants (3*2^k tests)"
for lib in x3m1 x3m2; do SPFFT_B200_LIB=$V/libspfft_b200_$lib.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "three_times" 2>&1 | tail -n 2; done
echo "=== bands"
for cfg in "--size 192 --bands 8" "--size 192 --bands 64" "--size 192 --bands 64 --precision single" "--size 128 --bands 64" "--size 64 --bands 256"; do
  echo "=== bands $cfg"
  timeout 600 python bench.py $cfg --no-cpu-baseline --steps 10 2>>gpurun_out/exp.err | tee gpurun_out/bench_bands_$(echo $cfg | tr -d ' -').json | python -c "$show"
done
tail -n 5 gpurun_out/exp.err
