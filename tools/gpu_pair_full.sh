#!/bin/bash
# full -m gpu suite with the CTA-pair kernels as the product path, then the bench line
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_full_r2b.log 2>&1
echo "full suite rc=$?"; tail -n 4 gpurun_out/t_full_r2b.log
( timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err ); echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['fp64']['ms_per_step'] if 'fp64' in d else None)
print(d['phases_ms']); print(d['roofline']['per_kernel_ms']); print(d['clocks'])
PY
