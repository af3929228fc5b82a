set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python tools/parity_report.py > gpurun_out/parity_fp32bb.log 2>&1
tail -60 gpurun_out/parity_fp32bb.log
timeout 600 python tools/parity_report.py --tf32-backbone > gpurun_out/parity_tf32bb.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
