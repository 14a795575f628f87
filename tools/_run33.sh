python -m pytest tests -m gpu -x -q -k "golden or ragged or closed or full" 2>&1 | tail -2
python tools/closed_loop_bench.py 2>&1 | grep gpi | cut -c1-200
for c in c2 c3; do python tools/quick_bench.py --kernel gpi --config $c --mode strict --reps 3 2>&1 | tail -1 | cut -c1-170; done
