set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
python tools/gemm_one.py 16480 2048 512 tf32 128
python tools/gemm_one.py 16480 2048 512 tf32x3 128
python tools/gemm_one.py 16480 2048 512 tf32x3 256
python tools/gemm_one.py 8192 8192 2048 tf32 128
python tools/gemm_one.py 8192 8192 2048 tf32x3 256
python tools/gemm_one.py 361 361 32 tf32x3 128 320
python tools/gemm_one.py 2060 2060 64 tf32x3 128 64
python tools/gemm_one.py 14440 256 256 tf32x3 128
timeout 900 python -m pytest tests/test_predict_gpu.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu2.log 2>&1
tail -4 gpurun_out/pytest_gpu2.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-episodes 0 > gpurun_out/bench_e8.log 2>&1
tail -1 gpurun_out/bench_e8.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e8.csv python tools/profile_step.py 8 interactron_random 2 > gpurun_out/profile_step.log 2>&1
