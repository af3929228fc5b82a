set -x
python -m pytest tests/test_predict_gpu.py tests/test_evaluator_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s9_pytest.log; tail -15 gpurun_out/s9_pytest.log
python tools/eval_post_bench.py 32 2>&1 | tail -2 | tee gpurun_out/s9_eval_post_32.json
python tools/eval_post_bench.py 256 2>&1 | tail -1 | tee gpurun_out/s9_eval_post_256.json
python bench.py --workload rollout --episodes 8 --steps 3 --warmup 3 > gpurun_out/s9_rollout_lock8.json 2> gpurun_out/s9_rollout.err; cut -c1-900 gpurun_out/s9_rollout_lock8.json; tail -3 gpurun_out/s9_rollout.err
python bench.py --workload rollout --episodes 32 --steps 3 --warmup 3 > gpurun_out/s9_rollout_lock32.json 2> gpurun_out/s9_rollout.err; cut -c1-900 gpurun_out/s9_rollout_lock32.json; tail -3 gpurun_out/s9_rollout.err
python bench.py --workload rollout --episodes 8 --sequential --steps 3 --warmup 3 > gpurun_out/s9_rollout_seq.json 2> gpurun_out/s9_rollout.err; cut -c1-900 gpurun_out/s9_rollout_seq.json; tail -3 gpurun_out/s9_rollout.err
