# round 2, GPU session 20: second wave of CTAs started half a part late (anti-phase)
set -x
mkdir -p gpurun_out
for st in 0 3000 5000 8000; do
SPFFT_B200_WSTAGGER=$st timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > gpurun_out/bench_wfft_v12_$st.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "stagger $st"; cut -c1-2800 gpurun_out/bench_wfft_v12_$st.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}'
done
