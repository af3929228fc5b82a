set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s8_pytest.log; tail -25 gpurun_out/s8_pytest.log
python tools/eval_post_bench.py 32 2>&1 | tail -2
python tools/eval_post_bench.py 256 2>&1 | tail -1
python bench.py --workload meta_interactron --episodes 2 --steps 6 --warmup 3 --cpu-episodes 0 2>&1 | tail -1 | cut -c1-330
python bench.py --episodes 1 --steps 10 --warmup 3 --cpu-episodes 0 2>&1 | tail -1 | cut -c1-330
