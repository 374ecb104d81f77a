for u in 1 2 4; do for f in 1 0; do
TDVMC_SWEEP_UNROLL=$u TDVMC_SWEEP_FIX=$f python bench.py --steps 3 --warmup 3 --no-exhibits 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels'][0]
print('unroll $u fix $f', round(d['value']/1e6,1), 'M/s  sweep avg ms', round(k['avg_launch_ms'],3))"
done; done
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 1 -f -o gpurun_out/sweep_r01e_fix python bench.py --steps 1 --warmup 3 --no-exhibits > /dev/null 2>&1
ls gpurun_out/*.ncu-rep
