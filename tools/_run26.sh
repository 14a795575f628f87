python -m pytest tests -m gpu -x -q -k "golden or ragged or full_size" 2>&1 | tail -2
for b in 4736 9472 65536; do python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 3 --B $b 2>&1 | tail -1 | cut -c1-170; done
python tools/quick_bench.py --kernel gpi --config c3 --mode strict --reps 3 2>&1 | tail -1 | cut -c1-170
for c in 4736 9472 18944 65536; do echo "== host chunk $c"; TINYMPC_HOST_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['kernel_launches_per_step'])"; done
echo "== default"; python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['kernel_launches_per_step'])"
