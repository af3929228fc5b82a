set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_e8.log 2>&1
tail -1 gpurun_out/bench_e8.log
timeout 600 python bench.py --steps 10 --warmup 3 --episodes 1 --cpu-episodes 0 > gpurun_out/bench_e1.log 2>&1
tail -1 gpurun_out/bench_e1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e8.csv python tools/profile_step.py 8 interactron_random 2 > gpurun_out/profile_step.log 2>&1
tail -3 gpurun_out/profile_step.log
