set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -c 300 gpurun_out/bench_r02.err
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
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0"
$B > gpurun_out/bench_for_ncu.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ncu_launches_bench_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
M="python tools/cg_microbench.py --f64 --batch 1 --pairs 16 --operator gaussian_blur --iters 200 --reps 2"
$M > gpurun_out/mb.log 2>&1 && tail -1 gpurun_out/mb.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fh_k_cg_res64 -s 2 -c 1 -f -o gpurun_out/res64_r02_final $M > gpurun_out/ncu_full1.log 2>&1; tail -2 gpurun_out/ncu_full1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fh_k_upd64 -s 28 -c 2 -f -o gpurun_out/upd64_r02_final $M > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out/*.ncu-rep
