# round 2, GPU session 9: occupancy sensitivity of the fused xy kernel (1 CTA per SM) + ncu of the v6 code
set -x
mkdir -p gpurun_out
SPFFT_B200_WPAD=60 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity > gpurun_out/bench_wfft_v6_occ1.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cut -c1-2600 gpurun_out/bench_wfft_v6_occ1.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_wxy -c 2 -o gpurun_out/r02_wfft_v6 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > gpurun_out/ncu_wfft.log 2>&1
tail -3 gpurun_out/ncu_wfft.log
