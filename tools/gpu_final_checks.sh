# round-end style checks on one B200: sanitizer over the newest kernels, then both bench arms
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -x \
    -k "(lupp and not 1500) or cg_reference or lsqr_reference or (lsqr_matches and 3001) or (cgls_matches and 3001) or (conjugate and 50)" ) > gpurun_out/sanitizer_memcheck_new.log 2>&1
echo "memcheck rc=$?"; tail -n 4 gpurun_out/sanitizer_memcheck_new.log
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -x \
    -k "(lupp and 257) or (lupp and ties) or cg_reference" ) > gpurun_out/sanitizer_racecheck_new.log 2>&1
echo "racecheck rc=$?"; tail -n 4 gpurun_out/sanitizer_racecheck_new.log
( time python bench.py --impl reference ) > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
tail -c 600 gpurun_out/bench_ref_final.json; tail -n 4 gpurun_out/bench_ref_final.err
( time python bench.py ) > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
head -c 700 gpurun_out/bench_final2.json; echo; tail -n 4 gpurun_out/bench_final2.err
