python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in c2 c3 c4; do for m in strict fast; do
python tools/quick_bench.py --kernel tpi --config $c --mode $m --reps 3 2>&1 | tail -1 | cut -c1-250
done; done
