cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in c2 c3; do timeout 300 python tools/quick_bench.py --config $c --kernel gpi --reps 4 2>&1 | tail -1; done > gpurun_out/r02k_quick.txt
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden or ragged or full_size_identical or full_size_tracking" 2>&1 | tail -3) > gpurun_out/r02k_pytest.txt
(timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02k_bench.err | tail -1) > gpurun_out/r02k_bench.json
cut -c1-200 gpurun_out/r02k_quick.txt; tail -2 gpurun_out/r02k_pytest.txt; cut -c1-300 gpurun_out/r02k_bench.json; tail -2 gpurun_out/r02k_bench.err
