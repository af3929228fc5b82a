set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/s5_pytest.log; tail -25 gpurun_out/s5_pytest.log
python tools/trainer_bench.py interactron > gpurun_out/s5_trainer_bench.json 2> gpurun_out/s5_trainer_bench.err; cat gpurun_out/s5_trainer_bench.json; tail -3 gpurun_out/s5_trainer_bench.err
python tools/trainer_bench.py interactron_random 2>&1 | tail -1
