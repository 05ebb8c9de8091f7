for p in 0 8 28; do FH_TRACE=1 python tools/cg_microbench.py --f64 --batch 1 --pairs $p --operator motion_blur --iters 200 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d.get('trace_us',{}).get('6'); print('motion pairs', d['pairs'], 'us/it', round(d['us_per_iteration'],2), 'gather_at', round(t[12]-t[11],2), 'gather_a', round(t[15]-t[14],2))"; done
python -m pytest tests -m gpu -q -x -k "motion or config" 2>&1 | tail -3
python bench.py --operator motion_blur --solver euler --num-steps 100 --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 > gpurun_out/bench_r02_motion_blur_b1.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_motion_blur_b1.json').read().strip().splitlines()[-1]); print('motion b1 value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['hot_path']['phases_ms'], d['roofline']['us_per_cg_iteration'], d['roofline']['rank0']['us_per_cg_iteration'])"
