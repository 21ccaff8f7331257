mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "launches", d["gpu_launches"], d["roofline"]["stage_ms"])'
true
true
for cfg in "--size 64 --bands 256" "--size 128 --bands 64" "--size 192 --bands 64" "--size 192 --bands 64 --precision single" "--size 256 --bands 32" "--size 96 --bands 128"; do
  for batch in 1 0; do
  echo "=== bands $cfg SPFFT_B200_BATCH=$batch"
  SPFFT_B200_BATCH=$batch timeout 600 python bench.py $cfg --no-cpu-baseline --no-e2e --steps 10 2>>gpurun_out/exp.err | python -c "$show"
  done
done
tail -n 5 gpurun_out/exp.err
