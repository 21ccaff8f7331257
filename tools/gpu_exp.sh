mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "--size 256 --type r2c" "--size 256" "--size 192" "--size 192 --precision single" "--size 128" "--size 512 --precision single"; do
echo "=== bench $cfg"; timeout 300 python bench.py $cfg --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'pairs/s pair_frac', round(d['roofline']['pair_frac'],3), d['roofline']['stage_ms'])"
done
