python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in c2 c3; do for m in fast strict; do
python tools/quick_bench.py --kernel tpi --config $c --mode $m --reps 3 2>&1 | tail -1 | cut -c1-260
done; done
python tools/quick_bench.py --kernel tpi --config c4 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-260
python tools/quick_bench.py --kernel tpi --config c4 --mode fast --reps 3 2>&1 | tail -1 | cut -c1-260
