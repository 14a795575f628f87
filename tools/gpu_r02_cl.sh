cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpi_solve -s 4 -c 1 -f -o gpurun_out/r02_gpi_closed_loop python tools/closed_loop_bench.py > gpurun_out/r02cl_ncu.log 2>&1
tail -3 gpurun_out/r02cl_ncu.log | cut -c1-200
