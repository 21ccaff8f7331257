# round 2, GPU session 29: half-tile z kernels (4 warps, two buffers, three CTAs per SM) against the 8-stick tile kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "warp_fft" > gpurun_out/pytest_wfft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wfft.log
tail -3 gpurun_out/pytest_wfft.log
for z4 in 1 0; do
SPFFT_B200_WZ4=$z4 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-gpu-reference > gpurun_out/bench_wfft_v15_$z4.json 2> gpurun_out/bench_wfft.err; tail -5 gpurun_out/bench_wfft.err; echo "wz4 $z4"; cut -c1-2800 gpurun_out/bench_wfft_v15_$z4.json | grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"stage_ms": {[^}]*}\|"ok": [a-z]*'
done
