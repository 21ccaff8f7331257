# under gpurun --gpus 2: the whole GPU suite (incl. the 2-GPU distributed parity check), smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -n 3
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.err
python -c 'import json; d=json.load(open("gpurun_out/bench.json")); print(round(d["value"],1), d["roofline"]["pair_frac"], d["roofline"]["stage_ms"], "e2e", d["e2e"]["value"], d["clocks"])'
