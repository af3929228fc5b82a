set -x
echo "== BK=16"
python tools/gemm_one.py 16480 2048 512 2>&1 | tail -1
python tools/gemm_one.py 8192 8192 2048 2>&1 | tail -1
python tools/gemm_one.py 57760 512 4608 2>&1 | tail -1
python tools/gemm_one.py 57760 256 256 2>&1 | tail -1
python tools/gemm_one.py 1805 256 2048 tf32x3 256 32 2>&1 | tail -1
python tools/gemm_one.py 16480 2048 512 tf32 2>&1 | tail -1
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -6
python tools/gemm_precision.py 2>&1 | tail -9
python bench.py --steps 8 --warmup 3 --cpu-episodes 0 2>&1 | tail -1 | cut -c1-220
