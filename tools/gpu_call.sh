set -x
nvidia-smi -L
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/s6_bench_n2.json 2> gpurun_out/s6_bench_n2.err; tail -1 gpurun_out/s6_bench_n2.json | cut -c1-700; tail -2 gpurun_out/s6_bench_n2.err
$T --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/s6_ref_n2.json 2> gpurun_out/s6_ref_n2.err; tail -1 gpurun_out/s6_ref_n2.json | cut -c1-300; tail -2 gpurun_out/s6_ref_n2.err
$T --master-port 29513 bench.py --gpus 2 --workload meta_interactron --episodes 2 --steps 6 --warmup 3 > gpurun_out/s6_meta_n2.json 2> gpurun_out/s6_meta_n2.err; tail -1 gpurun_out/s6_meta_n2.json | cut -c1-900; tail -2 gpurun_out/s6_meta_n2.err
python bench.py --workload meta_interactron --episodes 2 --steps 6 --warmup 3 --cpu-episodes 0 2>/dev/null | tail -1 | cut -c1-400
