for b in 4736 8192 9472 18944 33152; do python tools/quick_bench.py --kernel gpi --config c2 --mode strict --reps 3 --B $b 2>&1 | tail -1 | cut -c1-200; done
