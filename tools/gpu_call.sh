set -x
python -m pytest tests/test_predict_gpu.py -m gpu -x -q 2>&1 | tail -8
