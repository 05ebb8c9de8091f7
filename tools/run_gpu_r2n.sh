set -x
mkdir -p gpurun_out
for op in motion_blur super_resolution; do
  M="python tools/cg_microbench.py --f64 --batch 1 --pairs 0 --operator $op --iters 200 --reps 2"
  $M > gpurun_out/mb_$op.log 2>&1 && tail -1 gpurun_out/mb_$op.log | cut -c1-200
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:fh_k_cg_res64 -s 2 -c 1 -f -o gpurun_out/res64_r02_$op $M > gpurun_out/ncu_full_$op.log 2>&1; tail -2 gpurun_out/ncu_full_$op.log
done
ls -la gpurun_out/*.ncu-rep
