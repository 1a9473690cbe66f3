mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/t_full.log 2>&1
tail -n 45 gpurun_out/t_full.log
