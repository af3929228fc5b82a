python tools/gemm_trace.py 225000 256 64 tf32x3 | head -1
python tools/gemm_epi_bench.py 225000 256 64 2>&1 | head -10
python tools/gemm_one.py 16480 2048 512 tf32 128
python tools/gemm_one.py 16480 2048 512 tf32x3 256
python tools/gemm_one.py 8192 8192 2048 tf32 128
python tools/gemm_one.py 8192 8192 2048 tf32x3 256
python tools/gemm_one.py 14440 256 256 tf32x3
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-episodes 0 > gpurun_out/bench_e8.log 2>&1
tail -1 gpurun_out/bench_e8.log | cut -c1-300
timeout 900 python bench.py --steps 6 --warmup 3 --cpu-episodes 0 --episodes 32 > gpurun_out/bench_e32.log 2>&1
tail -1 gpurun_out/bench_e32.log | cut -c1-300
