# round 2: single precision with the round-1 backward z kernel + warp kernels for xy and forward z
set -x
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -k "warp_fft_kernels_single_precision or test_single_precision or smoke" > gpurun_out/r02_pytest_gpu_single2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_single2.log
tail -4 gpurun_out/r02_pytest_gpu_single2.log | cut -c1-300
B="--precision single --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-reference --no-parity"
timeout 60 python bench.py $B > gpurun_out/r02_bench_single_warp2.json 2> gpurun_out/bench_s15b.err; tail -2 gpurun_out/bench_s15b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_single_warp2.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d["roofline"].get("stage_ms"), d["roofline"].get("pair_frac"))
PY
