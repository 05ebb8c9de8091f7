for i in 1 2 3; do
python bench.py --operator inpainting --solver heun --num-steps 30 --batch 1 --steps 10 --warmup 2 --no-cpu-baseline --no-unet-timing --batched 0 > gpurun_out/stall.json 2> gpurun_out/stall_$i.err
python -c "
import json
d=json.loads(open('gpurun_out/stall.json').read().strip().splitlines()[-1]); print('run $i value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), [round(v) for v in d['e2e']['ms_steps']], [round(v) for v in d['ms_steps']])"
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-unet-timing --no-unet-in-loop --batched 0 > gpurun_out/stall.json 2> gpurun_out/stall_g.err
python -c "
import json
d=json.loads(open('gpurun_out/stall.json').read().strip().splitlines()[-1]); print('gauss value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), [round(v) for v in d['e2e']['ms_steps']], [round(v) for v in d['ms_steps']])"
