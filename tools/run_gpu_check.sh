set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/parity_report.py > gpurun_out/parity_x3_fp32bb.log 2>&1
grep -v -i warn gpurun_out/parity_x3_fp32bb.log | tail -70
timeout 600 python tools/parity_report.py --tf32-backbone > gpurun_out/parity_x3_tf32bb.log 2>&1
