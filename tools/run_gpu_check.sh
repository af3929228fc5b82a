set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/gemm_check.py --group perf > gpurun_out/gemm_perf.log 2>&1
cat gpurun_out/gemm_perf.log | tail -30
timeout 600 python tools/backbone_bench.py > gpurun_out/backbone_bench.log 2>&1
grep -v -i warn gpurun_out/backbone_bench.log
timeout 900 python -m pytest tests/test_predict_gpu.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu2.log 2>&1
tail -4 gpurun_out/pytest_gpu2.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-episodes 0 > gpurun_out/bench_e8.log 2>&1
tail -1 gpurun_out/bench_e8.log
