# experiment harness: per-iteration time of the float64 resident kernel (x4 SR, motion blur)
for cfg in "super_resolution 0" "motion_blur 0" "motion_blur 8" "motion_blur 28"; do
  set -- $cfg
  FH_TRACE=1 python tools/cg_microbench.py --f64 --batch 1 --pairs $2 --operator $1 --iters 200 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 pairs', d['pairs'], 'us/it', round(d['us_per_iteration'],2), d['cg_path'], 'niter', d['niter'], d.get('trace_us',{}).get('6'))"
done
