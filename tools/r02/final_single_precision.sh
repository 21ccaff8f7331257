# round 2, last GPU call: the whole GPU suite on the final binary (single precision on the warp kernels by default),
# the new single-precision tests on their own (a dead-locked persistent kernel must not take the suite with it),
# and 512^3 single precision timed with the round-1 kernels (SPFFT_B200_WFFT=7) and with the warp kernels
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "not warp_fft_kernels_single_precision" > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_final.log
tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 150 python -m pytest tests -m gpu -q -k "warp_fft_kernels_single_precision" > gpurun_out/r02_pytest_gpu_single.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/r02_pytest_gpu_single.log
tail -25 gpurun_out/r02_pytest_gpu_single.log | cut -c1-300
B="--precision single --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity"
SPFFT_B200_WFFT=7 timeout 120 python bench.py $B > gpurun_out/r02_bench_single_r1kernels.json 2> gpurun_out/bench_s7.err; tail -2 gpurun_out/bench_s7.err
if [ $rc -eq 0 ]; then
  timeout 120 python bench.py $B > gpurun_out/r02_bench_single_warp.json 2> gpurun_out/bench_s15.err; tail -2 gpurun_out/bench_s15.err
fi
python - <<'PY'
import json
for f in ("r02_bench_single_r1kernels", "r02_bench_single_warp"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), d["roofline"].get("stage_ms"), d["roofline"].get("pair_frac"))
    except Exception as e:
        print(f, "no line:", e)
PY
