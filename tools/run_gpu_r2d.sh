set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8
for pairs in 1 8 16; do
python tools/cg_microbench.py --f64 --batch 1 --pairs $pairs --operator gaussian_blur --iters 200 --reps 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('pairs $pairs', d['us_per_iteration'], d['niter'], d['cg_path'])"
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 > gpurun_out/bench_taps.json 2> gpurun_out/bench_taps.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_taps.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'], 'r32', r['us_per_cg_iteration'], 'r0', r['rank0']['us_per_cg_iteration'])
PY
