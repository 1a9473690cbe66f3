mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x -k "gemv or lsqr or cgls or blendenpik or lsrn or saddle or c4_full or conjugate or golden" ) > gpurun_out/t_gemvt.log 2>&1; tail -n 3 gpurun_out/t_gemvt.log
timeout 300 python tools/perf_lsqr.py > gpurun_out/perf_lsqr2.log 2>&1; head -n 3 gpurun_out/perf_lsqr2.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"gemv" --csv --log-file gpurun_out/ncu_lsqr2.csv python tools/ncu_target_lsqr.py > gpurun_out/ncu_lsqr2.log 2>&1
grep "gemv_t_kernel" gpurun_out/ncu_lsqr2.csv | grep "time_duration" | cut -d, -f13-15
