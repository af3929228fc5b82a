set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err; cut -c1-200 gpurun_out/s28_bench.json
