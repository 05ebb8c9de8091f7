# last check of the head build: per-iteration times of the four operators at rank 0 and the operator / full-config tests
for op in gaussian_blur inpainting motion_blur super_resolution; do python tools/cg_microbench.py --f64 --batch 1 --pairs 0 --operator $op --iters 200 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$op rank 0 us/it', round(d['us_per_iteration'],2))"; done
python -m pytest tests -m gpu -q -x -k "motion or config" 2>&1 | tail -3
