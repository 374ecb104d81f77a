#!/bin/bash
# r02a: parity at the tightened tolerances + fresh ncu captures of the FINAL sweep and solve kernels (VERDICT weak #6)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02a_tests.log
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 1 -f -o gpurun_out/sweep_r02a \
    python bench.py --steps 1 --warmup 3 --no-exhibits > gpurun_out/r02a_ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/solve_r02a \
    python bench.py --steps 1 --warmup 3 --no-exhibits > gpurun_out/r02a_ncu_solve.log 2>&1
cat gpurun_out/r02a_tests.log
