set -x
python tools/gemm_trace.py 46208 361 32 tf32x3 0 pad 384 2>&1 | head -3
python tools/gemm_trace.py 46208 361 32 tf32x3 0 pad 2>&1 | head -3
python tools/gemm_trace.py 46208 384 32 tf32x3 0 pad 2>&1 | head -3
python tools/gemm_trace.py 16480 2048 512 2>&1 | head -3
python tools/gemm_trace.py 46208 512 32 2>&1 | head -3
