cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "closed_loop" 2>&1 | tail -12) > gpurun_out/r02t_pytest.txt
cat gpurun_out/r02t_pytest.txt
