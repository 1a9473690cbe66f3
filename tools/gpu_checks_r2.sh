# round-2 checks on one B200: the full -m gpu suite, then compute-sanitizer over the int8 tensor-core kernels
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_full_r2c.log 2>&1
echo "full suite rc=$?"; tail -n 4 gpurun_out/t_full_r2c.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "bit_identical or accumulators_are_drained or (i8_gemm_accuracy and 300) or (int8_range_passes and 6000 and (3 or 1))" ) > gpurun_out/sanitizer_memcheck_i8_r2c.log 2>&1
echo "memcheck rc=$?"; tail -n 5 gpurun_out/sanitizer_memcheck_i8_r2c.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "bit_identical" ) > gpurun_out/sanitizer_racecheck_i8_r2c.log 2>&1
echo "racecheck rc=$?"; tail -n 5 gpurun_out/sanitizer_racecheck_i8_r2c.log
grep -c "Race reported\|hazard" gpurun_out/sanitizer_racecheck_i8_r2c.log
