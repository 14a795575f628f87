# compute-sanitizer passes over small parity cases of the three kernel families
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
(timeout 900 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden and (rocket_soc_N10_f64 or quad_tvlin_f32 or quad_hover_N10_f32 or lti_8_2_f64)" 2>&1 | tail -12) > gpurun_out/r02l_memcheck.txt
(timeout 900 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden and (rocket_soc_N10_f64 or quad_hover_N10_f32) and (gps or gpi)" 2>&1 | tail -12) > gpurun_out/r02l_racecheck.txt
(timeout 600 $S --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden and (rocket_soc_N10_f64) and gps" 2>&1 | tail -8) > gpurun_out/r02l_synccheck.txt
tail -5 gpurun_out/r02l_memcheck.txt; tail -5 gpurun_out/r02l_racecheck.txt; tail -4 gpurun_out/r02l_synccheck.txt
