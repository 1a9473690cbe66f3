mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_next_rows.py -m gpu -x -q -k lsqr ) > gpurun_out/t_lsqr.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k degenerate ) > gpurun_out/t_degen.log 2>&1
tail -5 gpurun_out/t_lsqr.log gpurun_out/t_degen.log
