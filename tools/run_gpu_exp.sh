for p in 0; do FH_TRACE=1 python tools/cg_microbench.py --f64 --batch 1 --pairs $p --operator super_resolution --iters 200 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d.get('trace_us',{}).get('6'); print('SR pairs', d['pairs'], 'us/it', round(d['us_per_iteration'],2), t)"; done
