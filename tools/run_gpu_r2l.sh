set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/gputests.log 2>&1; tail -4 gpurun_out/gputests.log
python tools/parity_summary.py > gpurun_out/parity_r02.md 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -c 300 gpurun_out/bench_r02.err
for cfg in "inpainting heun 30 1" "super_resolution heun 30 1" "motion_blur euler 100 1" "super_resolution heun 30 16"; do
  set -- $cfg
  python bench.py --operator $1 --solver $2 --num-steps $3 --batch $4 --steps 3 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 > gpurun_out/bench_r02_$1_b$4.json 2> gpurun_out/bench_r02_$1_b$4.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    if 'ref' in f: print(f, d['value'], d['cpu_baseline']['sample'][:200]); continue
    r=d.get('roofline',{})
    print(f, 'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms',d['ms_per_step'],'it/img',d['run']['cg_iterations_per_image'] if 'run' in d else None,'us/it',r.get('us_per_cg_iteration'),'frac',r.get('frac'),'step',(r.get('step') or {}).get('frac'), d['hot_path']['phases_ms'])
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0"
$B > gpurun_out/bench_for_ncu.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ncu_launches_bench_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
