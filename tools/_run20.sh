python -m pytest tests -m gpu -x -q -k "golden or ragged or closed" 2>&1 | tail -3
for m in strict fast; do python tools/quick_bench.py --kernel gpi --config c3 --mode $m --reps 3 2>&1 | tail -1 | cut -c1-250; done
python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-250
