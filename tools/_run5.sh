python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for chunk in -1 4736 9472 14208 18944 28416 37888; do
  echo "== TPI chunk $chunk"
  TINYMPC_TPI_CHUNK=$chunk python tools/quick_bench.py --kernel tpi --config c2 --mode fast --reps 3 2>&1 | tail -1 | cut -c1-260
  TINYMPC_TPI_CHUNK=$chunk python tools/quick_bench.py --kernel tpi --config c2 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-260
done
echo "== auto chunk"
python tools/quick_bench.py --kernel tpi --config c2 --mode fast --reps 3 2>&1 | tail -1 | cut -c1-260
python tools/quick_bench.py --kernel tpi --config c3 --mode fast --reps 3 2>&1 | tail -1 | cut -c1-260
python tools/quick_bench.py --kernel tpi --config c3 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-260
python tools/quick_bench.py --kernel tpi --config c4 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-260
echo "== GPI v3"
python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-260
python tools/quick_bench.py --kernel gpi --config c2 --mode fast --reps 3 2>&1 | tail -1 | cut -c1-260
python tools/quick_bench.py --kernel gpi --config c3 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-260
