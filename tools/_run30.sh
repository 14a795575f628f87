python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/closed_loop_bench.py 2>&1 | tail -7 | cut -c1-330
