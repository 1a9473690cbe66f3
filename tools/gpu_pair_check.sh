#!/bin/bash
# CTA-pair int8 sweeps: parity tests first (under a timeout: a barrier mistake would spin forever), then per-sweep timing with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "i8 or int8" > gpurun_out/t_pair_i8.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/t_pair_i8.log
tail -5 gpurun_out/t_pair_i8.log
timeout 300 python tools/i8_pass_time.py 110 7 > gpurun_out/pair_pass_time.log 2>&1; echo "rc=$?" >> gpurun_out/pair_pass_time.log
RNLA_I8_PAIR=0 timeout 300 python tools/i8_pass_time.py 110 7 > gpurun_out/single_pass_time.log 2>&1; echo "rc=$?" >> gpurun_out/single_pass_time.log
cat gpurun_out/pair_pass_time.log gpurun_out/single_pass_time.log
