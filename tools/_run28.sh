python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/n100_compare.py 2>&1 | tail -16
