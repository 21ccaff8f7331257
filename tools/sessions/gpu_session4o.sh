mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.read()); s=d["roofline"]["stage_ms"]; print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "all:", s)'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for pdl in 1 0; do
  for cfg in "--size 64" "--size 128" "--size 192" "--size 512" "--size 64 --bands 256"; do
    echo "=== PDL=$pdl $cfg"
    SPFFT_B200_PDL=$pdl timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e 2>>gpurun_out/exp.err | python -c "$show"
  done
done
# unprofiled loop (no events between the kernels): wall-clock pairs/s of a single small transform
for pdl in 1 0; do
SPFFT_B200_PDL=$pdl python - <<'PY'
import os, time, numpy as np, torch, sys
sys.path.insert(0, os.getcwd())
from spfft_b200 import capi
from bench import spherical_triplets
lib = capi.load()
for n in (64, 128, 192):
    trip = spherical_triplets(n, False)
    t = capi.Transform(lib, transform_type=0, dim_x=n, dim_y=n, dim_z=n, indices=trip)
    t.set_execution_mode(capi.SPFFT_EXEC_ASYNCHRONOUS)
    v = torch.rand(2 * len(trip), dtype=torch.float64, device="cuda")
    s = torch.empty(2 * n ** 3, dtype=torch.float64, device="cuda")
    o = torch.empty_like(v)
    for _ in range(20):
        t.backward_ptr(v, s); t.forward_ptr(s, o, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        t.backward_ptr(v, s); t.forward_ptr(s, o, 0)
    e1.record(); torch.cuda.synchronize()
    print("PDL", os.environ.get("SPFFT_B200_PDL"), n, "unprofiled", round(200 / (e0.elapsed_time(e1) * 1e-3), 1), "pairs/s")
PY
done
tail -n 5 gpurun_out/exp.err
