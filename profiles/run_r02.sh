#!/bin/bash
# r02 measurement pass on the GPU box (run under gpurun from the repo root); outputs land in gpurun_out/.
#   1. full GPU test suite (parity file gpurun_out/parity_r02.txt)
#   2. bench lines: our arm (default flags) and the reference arm
#   3. launch list of one bench step (cold-cache, serialised: compare SHARES, not absolutes)
#   4. ncu --set full captures of the kernels of the step and of the exhibits
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-exhibits --walkers-per-gpu 4096 > gpurun_out/r02_bench_under_ncu.log 2>&1   # (explicit W: no side ensembles)
ncu --set full --clock-control none --import-source on -k regex:sweep_queue_kernel -s 6 -c 1 -f -o gpurun_out/sweep_queue_r02 \
    python bench.py --steps 1 --warmup 3 --no-exhibits > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 2 -c 1 -f -o gpurun_out/evaluate_r02 \
    python bench.py --steps 1 --warmup 3 --no-exhibits > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 17 -c 1 -f -o gpurun_out/evaluate_n1728_r02 \
    python profiles/ab_evaluate.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tables_kernel|contract_kernel" -c 2 -f -o gpurun_out/tables_r02 \
    python profiles/ab_tables.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 5 -c 1 -f -o gpurun_out/syrk_r02 \
    python bench.py --steps 1 --warmup 3 > /dev/null 2>&1
cat gpurun_out/r02_tests.log
