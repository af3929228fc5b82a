set -x
python bench.py --steps 10 --warmup 3 --cpu-episodes 0 > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err; cut -c1-200 gpurun_out/s23_bench.json; tail -3 gpurun_out/s23_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 1 --steps 4 --warmup 3 --cpu-episodes 0 2>/dev/null | tail -1 | cut -c1-200
