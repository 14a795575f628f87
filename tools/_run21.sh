for c in c2 c3; do for mi in 1 2 5 10; do python tools/quick_bench.py --kernel gpi --config $c --mode strict --reps 3 --max_iter $mi 2>&1 | tail -1 | cut -c1-200; done; done
