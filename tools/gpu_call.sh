set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
