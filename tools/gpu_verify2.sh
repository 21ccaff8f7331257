# under gpurun --gpus 2: the whole GPU suite (incl. the 2-GPU distributed parity check), the default bench line
# and its ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.err
python -c 'import json; d=json.load(open("gpurun_out/bench.json")); print(round(d["value"],1), d["roofline"]["pair_frac"], d["roofline"]["stage_ms"], "e2e", d["e2e"]["value"])'
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
