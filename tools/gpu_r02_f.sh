# round-2 GPU session F: full parity suite + the contract bench (both arms) + launch list + sweeps, on the build with the TMA-ring GPS kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -12) > gpurun_out/r02f_pytest.txt
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>gpurun_out/r02f_bench_ref.err | tail -1) > gpurun_out/r02f_bench_ref.json
(timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02f_bench.err | tail -1) > gpurun_out/r02f_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
timeout 900 python tools/auto_rule_sweep.py --reps 1 > gpurun_out/r02f_auto_sweep.md 2>&1
timeout 900 python tools/roofline_sweep.py > gpurun_out/r02f_sweep_1gpu.md 2> gpurun_out/r02f_sweep.err
timeout 600 python tools/closed_loop_bench.py > gpurun_out/r02f_closed_loop.txt 2>&1
tail -3 gpurun_out/r02f_pytest.txt; cut -c1-400 gpurun_out/r02f_bench_n1.json; tail -4 gpurun_out/r02f_sweep_1gpu.md; tail -3 gpurun_out/r02f_closed_loop.txt
