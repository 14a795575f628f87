cd $GRAFT_REPO_ROOT
ls -d build 2>/dev/null && du -sh build
ncu --set full --clock-control none --import-source on -k regex:gps_solve -s 1 -c 1 -o gpurun_out/r02_gps_c4_v1 python tools/quick_bench.py --config c4 --kernel gps --reps 1 > gpurun_out/ncu_gps_c4.log 2>&1
tail -3 gpurun_out/ncu_gps_c4.log
