set -x
nvidia-smi -L | wc -l
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/s10_bench_n$N.json 2> gpurun_out/s10_bench_n$N.err; tail -1 gpurun_out/s10_bench_n$N.json | cut -c1-1500; tail -3 gpurun_out/s10_bench_n$N.err
