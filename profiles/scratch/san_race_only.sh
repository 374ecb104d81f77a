RACE=$(grep "^RACE=" profiles/run_sanitizer.sh | sed "s/^RACE='//; s/'$//")
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$RACE" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -5 gpurun_out/sanitizer_racecheck.log
