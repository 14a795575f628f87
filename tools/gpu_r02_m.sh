cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8) > gpurun_out/r02m_pytest.txt
for c in c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 3 2>&1 | tail -1; done > gpurun_out/r02m_quick.txt
timeout 900 python tools/auto_rule_sweep.py --reps 1 2>&1 | grep "float64" > gpurun_out/r02m_auto_f64.md
tail -3 gpurun_out/r02m_pytest.txt; cut -c1-200 gpurun_out/r02m_quick.txt; cat gpurun_out/r02m_auto_f64.md
