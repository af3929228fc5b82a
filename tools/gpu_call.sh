set -x
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/s30_traffic_e32.csv python tools/profile_step.py 32 interactron_random 2 > gpurun_out/s30_traffic.log 2>&1
python tools/traffic_summary.py gpurun_out/s30_traffic_e32.csv gpurun_out/s30_traffic_e32.json | tee gpurun_out/s30_traffic_summary.txt | head -12
python tools/launch_summary.py gpurun_out/s30_traffic_e32.csv --gemm > gpurun_out/s30_launch_summary.txt 2>&1; head -12 gpurun_out/s30_launch_summary.txt
