set -x
rm -f gpurun_out/parity_log.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -c 400 gpurun_out/bench_r02.err
for cfg in "inpainting heun 30 1" "super_resolution heun 30 1" "motion_blur euler 100 1" "super_resolution heun 30 16"; do
  set -- $cfg
  python bench.py --operator $1 --solver $2 --num-steps $3 --batch $4 --steps 3 --warmup 2 --no-cpu-baseline --no-unet-timing --batched 0 > gpurun_out/bench_r02_$1_b$4.json 2> gpurun_out/bench_r02_$1_b$4.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    if 'sweep' in f or 'gpu.json' in f or 'ref' in f: continue
    r=d.get('roofline',{})
    print(f, 'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms',d['ms_per_step'],'it/img',d['run']['cg_iterations_per_image'] if 'run' in d else None,'us/it',r.get('us_per_cg_iteration'),'frac',r.get('frac'),'step',(r.get('step') or {}).get('frac'), d['hot_path']['phases_ms'])
PY
