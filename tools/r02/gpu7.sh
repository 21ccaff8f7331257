# round 2, GPU session 7: independent warp groups inside the CTA (W = 2 / 4 / 8 columns per group)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for w in 2 4 8; do
SPFFT_B200_WGROUP=$w timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v5_w$w.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; cut -c1-2600 gpurun_out/bench_wfft_v5_w$w.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_wxy -c 2 -o gpurun_out/r02_wfft_v5 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-gpu-reference --no-stage-pass > gpurun_out/ncu_wfft.log 2>&1
tail -3 gpurun_out/ncu_wfft.log
