# multi-GPU bench (the driver's launch line):  gpurun --gpus N -- bash tools/bench_multi_gpu.sh N
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r02_bench_n$N.err | tail -1) > gpurun_out/r02_bench_n$N.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2> gpurun_out/r02_bench_ref_n$N.err | tail -1) > gpurun_out/r02_bench_reference_arm_n$N.json
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
cut -c1-400 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err; cut -c1-300 gpurun_out/r02_bench_reference_arm_n$N.json
