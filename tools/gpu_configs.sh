# One bench line per BASELINE.json config that fits one GPU, final code (under gpurun, 1 GPU)
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl
run() { echo "=== $*"; timeout 600 python bench.py "$@" --no-cpu-baseline 2>>gpurun_out/configs.err | tee -a gpurun_out/configs.jsonl | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "pairs/s pair_frac", round(d["roofline"]["pair_frac"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1))'; }
run --size 64
run --size 128
run --size 256 --type r2c
run --size 192 --bands 256 --steps 5
run --size 192 --bands 256 --precision single --steps 5
run --size 192 --bands 256 --type r2c --steps 5
run --size 64 --bands 256
tail -n 3 gpurun_out/configs.err
