set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python tools/gemm_check.py > gpurun_out/gemm_check.log 2>&1
tail -5 gpurun_out/gemm_check.log
timeout 600 python -m pytest tests -q -m gpu -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
