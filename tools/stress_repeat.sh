# repeat the parity suite / the full-size streamed cases to catch rare ordering bugs (TMA ring, refill queue)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do (timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -1); done > gpurun_out/r02_stress.txt
for i in 1 2 3 4 5 6 7 8; do (timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "full_size_rocket or full_size_tracking or ragged" 2>&1 | tail -1); done >> gpurun_out/r02_stress.txt
cat gpurun_out/r02_stress.txt
