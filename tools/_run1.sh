set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q -k "tpi or errors" 2>&1 | tail -30
python tools/quick_bench.py --kernel tpi --config c2 2>&1 | tail -3
python tools/quick_bench.py --kernel tpi --config c2 --mode fast 2>&1 | tail -3
python tools/quick_bench.py --kernel tpi --config c3 2>&1 | tail -3
python tools/quick_bench.py --kernel tpi --config c4 2>&1 | tail -3
