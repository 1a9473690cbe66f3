mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "evd2 or int8" ) > gpurun_out/t_evd2.log 2>&1; tail -n 15 gpurun_out/t_evd2.log
