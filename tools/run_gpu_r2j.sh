set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -6
for cfg in "gaussian_blur heun 30 1" "super_resolution heun 30 1" "motion_blur euler 100 1"; do
  set -- $cfg
  python bench.py --operator $1 --solver $2 --num-steps $3 --batch $4 --steps 3 --warmup 2 --no-cpu-baseline --no-unet-timing --batched 0 > gpurun_out/bq_$1.json 2> gpurun_out/bq_$1.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bq_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms',d['ms_per_step'], d['hot_path']['phases_ms'])
    print('   ', d['hot_path']['kernels_ms_event_pairs'])
PY
