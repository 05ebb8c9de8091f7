set -x
mkdir -p gpurun_out
for cfg in "gaussian_blur 0" "gaussian_blur 16" "inpainting 0" "inpainting 16" "motion_blur 0" "motion_blur 28" "super_resolution 0" "super_resolution 16"; do
  set -- $cfg
  FH_TRACE=1 python tools/cg_microbench.py --f64 --batch 1 --pairs $2 --operator $1 --iters 200 --reps 3 > gpurun_out/trace_$1_$2.json 2>/dev/null
done
M="python tools/cg_microbench.py --f64 --batch 1 --pairs 16 --operator gaussian_blur --iters 200 --reps 2"
$M > gpurun_out/mb.log 2>&1 && tail -1 gpurun_out/mb.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fh_k_upd64 -s 28 -c 2 -f -o gpurun_out/upd64_r02_final $M > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out/*.ncu-rep
