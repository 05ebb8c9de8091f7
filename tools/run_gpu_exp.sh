for p in 0 16; do FH_TRACE=1 python tools/cg_microbench.py --f64 --batch 1 --pairs $p --operator super_resolution --iters 200 --reps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d.get('trace_us',{}).get('6'); print('SR pairs', d['pairs'], 'us/it', round(d['us_per_iteration'],2), 'gather_at', round(t[12]-t[11],2), 'gather_a', round(t[15]-t[14],2))"; done
python -m pytest tests -m gpu -q -x -k "super or sr or config" 2>&1 | tail -3
python bench.py --operator super_resolution --solver heun --num-steps 30 --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 > gpurun_out/bench_r02_super_resolution_b1.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_super_resolution_b1.json').read().strip().splitlines()[-1]); print('SR b1 value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['hot_path']['phases_ms'], d['roofline']['us_per_cg_iteration'], d['roofline']['rank0']['us_per_cg_iteration'])"
