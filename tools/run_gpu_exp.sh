# experiment harness: per-iteration time of the float64 resident kernel (x4 SR, motion blur) + the SR parity tests
for cfg in "super_resolution 0" "super_resolution 8" "super_resolution 16" ; do
  set -- $cfg
  FH_TRACE=1 python tools/cg_microbench.py --f64 --batch 1 --pairs $2 --operator $1 --iters 200 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 pairs', d['pairs'], 'us/it', round(d['us_per_iteration'],2), d['cg_path'], 'niter', d['niter'], d.get('trace_us',{}).get('6'))"
done
python -m pytest tests -m gpu -q -x -k "super or sr or config" 2>&1 | tail -3
python bench.py --operator super_resolution --solver heun --num-steps 30 --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SR b1 value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['hot_path']['phases_ms'])"
