python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c2 c3; do for m in strict fast; do python tools/quick_bench.py --kernel gpi --config $c --mode $m --reps 3 2>&1 | tail -1 | cut -c1-230; done; done
python tools/quick_bench.py --kernel gpi --config c3 --mode strict --reps 3 --max_iter 1 2>&1 | tail -1 | cut -c1-200
python tools/quick_bench.py --kernel gpi --config c3 --mode strict --reps 3 --max_iter 5 2>&1 | tail -1 | cut -c1-200
