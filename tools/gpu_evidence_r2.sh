#!/bin/bash
# round-2 evidence on one B200: bench (both arms), ncu DRAM traffic of one call per mode, ncu launch list of the bench command
mkdir -p gpurun_out
( timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err ); echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r2_final.err
( timeout 900 python bench.py --impl reference > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err ); echo "ref rc=$?"; tail -c 400 gpurun_out/bench_r2_ref.json
for mode in auto fp64; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step_$mode.csv python tools/ncu_target_i8.py $mode > gpurun_out/ncu_step_$mode.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --secondary 0 --cpu-rows 0 > gpurun_out/bench_under_ncu_r2.log 2>&1
wc -l gpurun_out/ncu_step_auto.csv gpurun_out/ncu_step_fp64.csv gpurun_out/r02_ncu_launches_bench.csv
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'fp64', d['fp64']['ms_per_step'] if d.get('fp64') else None)
print(d['phases_ms']); print(d['roofline']['frac'], d['roofline']['per_kernel_ms']); print(d['clocks'])
print({k:(v.get('ms'), v.get('hbm_frac')) for k,v in (d.get('secondary') or {}).items() if isinstance(v, dict)})
PY
