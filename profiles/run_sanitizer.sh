#!/bin/bash
# compute-sanitizer over the parity tests (run on the GPU box through gpurun): memcheck on the whole
# small-size set, racecheck + synccheck on the kernels that use warp-synchronous shared-memory updates.
# Output: gpurun_out/sanitizer_{memcheck,racecheck,synccheck}.log
set -u
mkdir -p gpurun_out
SMALL='not statistics and not full_size and not n1728 and not n343 and not two_gpu and not cpp_host and not accumulate_fixed and not time_evolution and not headline_size and not 700 and not 400 and not 416 and not euler_step_at_headline'
timeout 900 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SMALL" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
RACE='(fixed_configuration and (n64_fixture or n216 or hebulk_n64_fixture or hedrop_n6_fixture or he4he4na_fixture)) or (sweep_replays and n216) or sample_and_accumulate or hedrop_chain or mixture_chain or (accumulate_fixed and 33) or (observables_fixed and n64) or cluster_observables_fixed or (edge_sizes and (33 or 128 or 3)) or update_stored or device_solver or boxradial_estimators or (boxradial_fixed and n27_equil) or (boxradial_sweep and n27) or (inhcontact and n3_equil) or mixture4_he4he4na_equil or (low_dimensional and (bosonsbulk2d or bosonsbulk1d)) or boxradial2d or device_qr_solver or (split_sweep and n343) or small_ensemble_forced'
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$RACE" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$RACE" > gpurun_out/sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log
for f in gpurun_out/sanitizer_*.log; do echo "== $f"; tail -n 4 "$f"; done
