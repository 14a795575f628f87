set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in c2 c3; do for m in strict fast; do
python tools/quick_bench.py --kernel gpi --config $c --mode $m 2>&1 | tail -1
done; done
python bench.py --steps 5 --warmup 3 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gpi_solve -c 1 -o gpurun_out/gpi_c2_strict_v2 python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 0 --max_iter 20 > gpurun_out/ncu_gpi_v2.log 2>&1
