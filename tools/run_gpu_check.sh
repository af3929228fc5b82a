timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --cpu-episodes 0 > gpurun_out/bench_e32.log 2>&1; tail -1 gpurun_out/bench_e32.log | cut -c1-250
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tf32 --csv --log-file gpurun_out/gemm_dram_e32.csv python tools/profile_step.py 32 interactron_random 2 > gpurun_out/prof_dram.log 2>&1
tail -2 gpurun_out/prof_dram.log; wc -l gpurun_out/gemm_dram_e32.csv
