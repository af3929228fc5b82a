set -x
python tools/gemm_epi_bench.py 225000 256 64 2>&1 | head -10
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-episodes 0 > gpurun_out/bench_e8.log 2>&1
tail -1 gpurun_out/bench_e8.log | cut -c1-300
timeout 900 python bench.py --steps 6 --warmup 3 --cpu-episodes 0 --episodes 32 > gpurun_out/bench_e32.log 2>&1
tail -1 gpurun_out/bench_e32.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e8.csv python tools/profile_step.py 8 interactron_random 2 > gpurun_out/profile_step.log 2>&1
