mkdir -p gpurun_out
V=spfft_b200/lib/variants
echo "=== pytest fused (TUNE=5)"; SPFFT_B200_TUNE=5 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not distributed" 2>&1 | tail -3
echo "=== trace"; SPFFT_B200_LIB=$V/libspfft_b200_trace.so SPFFT_B200_TUNE=5 timeout 300 python tools/xy_trace.py 2>&1 | tail -12
echo "=== bench fused"; SPFFT_B200_TUNE=5 timeout 300 python bench.py --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['stage_ms'])"
echo "=== bench separate"; timeout 300 python bench.py --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['stage_ms'])"
