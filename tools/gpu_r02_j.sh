cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6) > gpurun_out/r02j_pytest.txt
for c in c4 c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 3 2>&1 | tail -1; done > gpurun_out/r02j_quick.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.txt 2>&1
tail -2 gpurun_out/r02j_pytest.txt; cut -c1-200 gpurun_out/r02j_quick.txt; cat gpurun_out/r02j_smoke.txt
