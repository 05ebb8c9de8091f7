set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -c 300 gpurun_out/bench_r02.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms',d['ms_per_step'],'frac',r['frac'],'step',r['step']['frac'], d['hot_path']['phases_ms'], d['unet'].get('in_loop_eager'))
PY
