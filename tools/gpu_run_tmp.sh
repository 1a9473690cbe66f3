mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_next_rows.py -m gpu -q -k golden ) > gpurun_out/t_golden.log 2>&1; tail -n 3 gpurun_out/t_golden.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"gemv|axpby|lsqr|dot_" --csv --log-file gpurun_out/ncu_lsqr.csv python tools/ncu_target_lsqr.py > gpurun_out/ncu_lsqr.log 2>&1
tail -n 2 gpurun_out/ncu_lsqr.log; wc -l gpurun_out/ncu_lsqr.csv
