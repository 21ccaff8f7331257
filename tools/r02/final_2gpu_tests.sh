# round 2: the whole GPU test suite on a 2-GPU box (so that test_distributed_two_gpus runs too) + the default bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_2gpubox.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_2gpubox.log
tail -5 gpurun_out/r02_pytest_gpu_2gpubox.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline']['stage_ms'], d['roofline']['frac'], d['roofline']['pair_frac'], d['roofline']['traffic'], d['parity']['ok'], 'e2e', d['e2e']['value'], 'gpu_ref', d['gpu_reference']['ratio'])
PY
