set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | cut -c1-330
