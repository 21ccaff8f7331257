# round 2, GPU session 35: host-pointer calls with the space domain moved in slabs on a second stream
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host_pointer or warp_fft or internal_device or in_place" > gpurun_out/pytest_host.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_host.log
tail -4 gpurun_out/pytest_host.log
for sl in 4 1; do
SPFFT_B200_HOST_SLABS=$sl timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference --no-parity --no-stage-pass > gpurun_out/bench_e2e_slabs$sl.json 2> gpurun_out/bench_wfft.err; tail -3 gpurun_out/bench_wfft.err; echo "slabs $sl"; grep -o '"value": [0-9.]*, "unit": "pairs/s"\|"e2e": {[^}]*}' gpurun_out/bench_e2e_slabs$sl.json
done
