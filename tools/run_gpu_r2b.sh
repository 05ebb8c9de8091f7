set -x
rm -f gpurun_out/parity_log.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0"
$B > gpurun_out/bench_zt_auto.json 2> gpurun_out/bench_zt_auto.err; tail -c 300 gpurun_out/bench_zt_auto.err
FH_R64_ZT=0 $B > gpurun_out/bench_zt_off.json 2> gpurun_out/bench_zt_off.err
python - <<'PY'
import json
for n in ("zt_auto","zt_off"):
    try:
        d=json.loads(open(f'gpurun_out/bench_{n}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(n,'value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],d['ms_steps'],d['run']['cg_iterations_steps'],d['hot_path']['phases_ms'], 'r32', r['us_per_cg_iteration'], 'r0', r['rank0']['us_per_cg_iteration'], d['hot_path']['kernels_ms_event_pairs']['fh_k_cg_res64'])
    except Exception as e: print(n,'failed',e)
PY
