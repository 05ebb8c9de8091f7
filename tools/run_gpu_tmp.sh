set -x
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-unet-timing --batched 0 > gpurun_out/plain_bench.log 2>&1 &&
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 3400 --csv --log-file gpurun_out/ncu_launches_bench_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-unet-timing --batched 0 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/ncu_launches_bench_r02.csv
