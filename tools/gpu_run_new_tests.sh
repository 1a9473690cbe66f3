mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -k "cg or lsqr" ) > gpurun_out/t_cg.log 2>&1
( time timeout 600 python -m pytest tests/test_cpp_mirror.py -m gpu -q ) > gpurun_out/t_cpp.log 2>&1
tail -n 6 gpurun_out/t_cg.log gpurun_out/t_cpp.log
