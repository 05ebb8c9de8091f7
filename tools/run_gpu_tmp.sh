set -x
rm -f gpurun_out/parity_log.jsonl
timeout 1200 python -m pytest tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -8
python - <<'PY'
import json
for l in open('gpurun_out/parity_log.jsonl'):
    d=json.loads(l)
    if d['test']!='full_config': continue
    print(d['cfg'],'final',d['final'],'first',d['first_call'],'worst',d['worst_call'],d['worst_call_index'],'cg',d['cg_total'],d['ref_cg_total'],'same',d['calls_with_same_niter'],'/',d['calls'],'firstdiff',d['first_call_with_other_niter'])
    print('  errs',d['errs'][:12])
PY
