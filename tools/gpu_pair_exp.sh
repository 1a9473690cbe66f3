#!/bin/bash
mkdir -p gpurun_out
RNLA_I8_N16=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "i8_gemm or drained" 2>&1 | tail -3
for nbs in 4 3 6; do
echo "== N16 nbs $nbs" >> gpurun_out/pair_exp.log
RNLA_I8_NBS2=$nbs RNLA_I8_N16=1 timeout 300 python tools/i8_pass_time.py 110 7 2>&1 | sed 's/i8_mma//g' >> gpurun_out/pair_exp.log
done
cat gpurun_out/pair_exp.log
