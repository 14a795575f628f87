# round-2 GPU session A: full parity suite, contract bench (both arms), launch list, ncu captures of the C4 (gps) and C2 (gpi) kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/r02a_env.txt 2>&1
nproc >> gpurun_out/r02a_env.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02a_env.txt 2>&1
(timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -25) > gpurun_out/r02a_pytest.txt
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>gpurun_out/r02a_bench_ref.err | tail -1) > gpurun_out/r02a_bench_ref.json
(timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02a_bench.err | tail -1) > gpurun_out/r02a_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 1 -c 1 -o gpurun_out/r02a_gps_c4 python tools/quick_bench.py --config c4 --kernel gps --reps 1 > gpurun_out/r02a_ncu_gps_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 1 -c 1 -o gpurun_out/r02a_gpi_c2 python tools/quick_bench.py --config c2 --kernel gpi --reps 1 > gpurun_out/r02a_ncu_gpi_c2.log 2>&1
for c in c2 c3 c4; do timeout 300 python tools/quick_bench.py --config $c --kernel auto --reps 3 2>&1 | tail -1; done > gpurun_out/r02a_quick.txt
tail -3 gpurun_out/r02a_pytest.txt; cat gpurun_out/r02a_quick.txt; cut -c1-600 gpurun_out/r02a_bench_n1.json; tail -2 gpurun_out/r02a_bench.err
