mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -k lsqr ) > gpurun_out/t_lsqr.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k degenerate ) > gpurun_out/t_degen.log 2>&1
tail -n 5 gpurun_out/t_lsqr.log gpurun_out/t_degen.log
( time timeout 600 python -m pytest tests/test_cpp_mirror.py -m gpu -q ) > gpurun_out/t_cpp.log 2>&1
( timeout 600 python tools/perf_lsqr.py ) > gpurun_out/perf_lsqr.log 2>&1
tail -n 4 gpurun_out/t_cpp.log gpurun_out/perf_lsqr.log
