set -x
rm -f gpurun_out/mb64.jsonl gpurun_out/parity_log.jsonl
for p in 0 4 16 28; do
  FH_TRACE=1 timeout 120 python tools/cg_microbench.py --pairs $p --operator motion_blur --mode resident --iters 100 --f64 >> gpurun_out/mb64.jsonl 2>> gpurun_out/mb64.err
done
python - <<'PY'
import json
for l in open('gpurun_out/mb64.jsonl'):
    d=json.loads(l); print(d['operator'],d['pairs'],round(d['us_per_iteration'],2),d['cg_path'],d['niter'],d.get('trace_us',{}).get('6'))
PY
tail -5 gpurun_out/mb64.err
timeout 1500 python -m pytest tests -m gpu -q -k "motion or rank" 2>&1 | tail -5
