set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s14_pytest.log; tail -6 gpurun_out/s14_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/parity_report.py > gpurun_out/s14_parity_report.txt 2>&1; tail -25 gpurun_out/s14_parity_report.txt
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/s14_ref.json 2>/dev/null; cut -c1-200 gpurun_out/s14_ref.json
python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; cut -c1-400 gpurun_out/s14_bench.json; tail -2 gpurun_out/s14_bench.err
