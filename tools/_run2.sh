set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
for c in c2 c3; do for m in strict fast; do
python tools/quick_bench.py --kernel gpi --config $c --mode $m 2>&1 | tail -2
done; done
python tools/quick_bench.py --kernel tpi --config c3 --mode fast 2>&1 | tail -2
