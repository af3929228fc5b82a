set -x
python tools/gemm_scores.py 0 2>&1 | tail -1
python tools/gemm_scores.py 0 50 361 32 160 2>&1 | tail -1
python tools/gemm_scores.py 0 255 1805 64 32 2>&1 | tail -1
python -m pytest tests/test_kernels_gpu.py tests/test_predict_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --cpu-episodes 0 2>/dev/null | tail -1 | cut -c1-250
