#!/bin/bash
# one-pass kernel evidence on one B200: full ncu capture of the kernel next to the two streaming kernels, then the default bench
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normal_pass_kernel|gemv_n_kernel|gemv_t_kernel" -c 5 -f -o gpurun_out/r02_normal_pass python tools/ncu_target_onepass.py > gpurun_out/ncu_onepass.log 2>&1; tail -n 3 gpurun_out/ncu_onepass.log
( timeout 900 python bench.py > gpurun_out/bench_r2_onepass.json 2> gpurun_out/bench_r2_onepass.err ); echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r2_onepass.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_onepass.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'fp64', d['fp64']['ms_per_step'] if d.get('fp64') else None)
for k,v in (d.get('secondary') or {}).items():
    if k.startswith('c4'): print(k, {a:b for a,b in v.items() if a not in ('phases_ms','iteration')})
PY
