set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err
tail -c 1500 gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
print('N=$N value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_steps'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --operator motion_blur --solver euler --num-steps 100 --images $((4*N)) --steps 1 --warmup 1 --no-cpu-baseline --no-unet-timing --batched 0 > gpurun_out/bench_r02_sweep_${N}gpu.json 2> gpurun_out/bench_r02_sweep_${N}gpu.err
tail -c 600 gpurun_out/bench_r02_sweep_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_sweep_${N}gpu.json').read().strip().splitlines()[-1])
print('sweep N=$N', d['sweep'])
PY
