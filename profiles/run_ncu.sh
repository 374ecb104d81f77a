#!/bin/bash
# Profiling pass on the GPU box (run under gpurun from the repo root); outputs land in gpurun_out/.
#   1. launch list of one bench step (cold-cache, serialised: compare SHARES, not absolutes)
#   2. ncu --set full capture of the top kernels (sweep, evaluate, tables, syrk)
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-exhibits > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 1 -f -o gpurun_out/sweep_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-exhibits > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 2 -c 1 -f -o gpurun_out/evaluate_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-exhibits > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tables_kernel|contract_kernel" -c 2 -f -o gpurun_out/tables_${TAG} \
    python bench.py --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 5 -c 1 -f -o gpurun_out/syrk_${TAG} \
    python bench.py --steps 1 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out
