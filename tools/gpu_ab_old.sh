#!/bin/bash
# same-box comparison of the round-2-start library (tools/librnla_r2start.so, commit eec2ffe) with the current one
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["phases_ms"].items()}, {k[9:]: round(v, 2) for k, v in d["roofline"]["per_kernel_ms"].items()}, d["clocks"]["sm_mhz"])
PY
}
for rep in 1 2; do
  timeout 600 python tools/bench_old_lib.py tools/librnla_r2start.so --secondary 0 --cpu-rows 0 --e2e-steps 0 --mode auto > gpurun_out/ab_old.json 2> gpurun_out/ab_old.err; echo -n "round-2 start: "; show gpurun_out/ab_old.json || tail -3 gpurun_out/ab_old.err
  timeout 600 python bench.py --secondary 0 --cpu-rows 0 --e2e-steps 0 --mode auto > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err; echo -n "current      : "; show gpurun_out/ab_new.json
done
