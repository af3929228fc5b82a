set -x
python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; cut -c1-300 gpurun_out/s21_bench.json; tail -2 gpurun_out/s21_bench.err
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | cut -c1-200
