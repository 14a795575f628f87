python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in c2 c3; do for m in strict fast; do
python tools/quick_bench.py --kernel gpi --config $c --mode $m --reps 3 2>&1 | tail -1 | cut -c1-270
done; done
