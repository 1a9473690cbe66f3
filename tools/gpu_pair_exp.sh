#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "i8_gemm or drained" 2>&1 | tail -3
for rep in 1 2; do
echo "== both" >> gpurun_out/pair_exp.log
timeout 300 python tools/i8_pass_time.py 110 7 2>&1 | sed 's/i8_mma//g' >> gpurun_out/pair_exp.log
echo "== separate" >> gpurun_out/pair_exp.log
RNLA_I8_BOTH=0 timeout 300 python tools/i8_pass_time.py 110 7 2>&1 | sed 's/i8_mma//g' >> gpurun_out/pair_exp.log
done
cat gpurun_out/pair_exp.log
