for op in gaussian_blur; do
for pairs in 1 4 7 16; do
for zt in 0 1; do
echo "== $op pairs $pairs zt $zt"
FH_R64_ZT=$zt python tools/cg_microbench.py --f64 --batch 1 --pairs $pairs --operator $op --iters 200 --reps 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['us_per_iteration'], d['niter'], d['cg_path'])"
done; done; done
