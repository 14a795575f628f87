python tools/closed_loop_bench.py 2>&1 | tail -6
