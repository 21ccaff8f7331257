mkdir -p gpurun_out
V=spfft_b200/lib/variants
for cfg in "--size 512 --precision single" "--size 256 --precision single" "--size 192 --precision single"; do
echo "=== bench f32v8 $cfg"; SPFFT_B200_LIB=$V/libspfft_b200_f32v8.so timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'pairs/s pair_frac', round(d['roofline']['pair_frac'],3), d['roofline']['stage_ms'])"
done
echo "=== bench default 256 single"; timeout 300 python bench.py --size 256 --precision single --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'pairs/s pair_frac', round(d['roofline']['pair_frac'],3), d['roofline']['stage_ms'])"
echo "=== pytest f32v8 (float cases)"; SPFFT_B200_LIB=$V/libspfft_b200_f32v8.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "single or float" 2>&1 | tail -3
