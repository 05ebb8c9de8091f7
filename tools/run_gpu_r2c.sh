set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err ) 2>&1 | tail -4
tail -c 400 gpurun_out/bench_head.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_head.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
print(json.dumps(r)[:1500])
print(json.dumps(d.get('hot_path',{}))[:2500])
PY
