#!/bin/bash
# One-stop GPU check used during development (run on a B200 box, e.g. `gpurun -- bash tools/gpu_check.sh`):
# parity suite, drop-in shim, smoke, the contract bench (both arms) and the developer micro-benchmarks.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1
for c in c2 c3; do for k in gpi tpi; do python tools/quick_bench.py --kernel $k --config $c --mode strict --reps 3 2>&1 | tail -1; done; done
python tools/quick_bench.py --kernel tpi --config c4 --mode strict --reps 3 2>&1 | tail -1
