set -x
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.log 2>&1
tail -2 gpurun_out/bench_n2.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2_ref.log 2>&1
tail -1 gpurun_out/bench_n2_ref.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 330 -c 8 -o gpurun_out/r01_gemm_full -f python tools/profile_step.py 8 interactron_random 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
