python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 0.4 0.5 0.55 0.6 0.65 0.7; do
  echo "== hybrid frac $f"
  TINYMPC_HYBRID_GPI_FRACTION=$f python tools/quick_bench.py --kernel hybrid --config c2 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-270
  TINYMPC_HYBRID_GPI_FRACTION=$f python tools/quick_bench.py --kernel hybrid --config c3 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-270
done
TINYMPC_HYBRID_GPI_FRACTION=0.55 python tools/quick_bench.py --kernel hybrid --config c2 --mode fast --reps 3 2>&1 | tail -1 | cut -c1-270
