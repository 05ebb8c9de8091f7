set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/gputests.log 2>&1; tail -6 gpurun_out/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for cfg in "gaussian_blur heun 30 1" "super_resolution heun 30 1" "motion_blur euler 100 1" "inpainting heun 30 1"; do
  set -- $cfg
  python bench.py --operator $1 --solver $2 --num-steps $3 --batch $4 --steps 3 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 > gpurun_out/bq_$1.json 2> gpurun_out/bq_$1.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bq_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    r=d['roofline']
    print(f, 'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms',d['ms_per_step'], 'r32', r.get('us_per_cg_iteration'), 'r0', (r.get('rank0') or {}).get('us_per_cg_iteration'), d['hot_path']['phases_ms'])
    print('   ', d['hot_path']['kernels_ms_event_pairs'])
PY
