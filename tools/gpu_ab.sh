#!/bin/bash
# A/B of an environment switch inside one box: bash tools/gpu_ab.sh VAR=VALUE   (device-resident bench only, no side legs)
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["phases_ms"].items()}, {k[9:]: round(v, 2) for k, v in d["roofline"]["per_kernel_ms"].items()}, d["clocks"]["sm_mhz"])
PY
}
for rep in 1 2; do
  timeout 600 python bench.py --secondary 0 --cpu-rows 0 --e2e-steps 0 --mode auto > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err; echo -n "base   : "; show gpurun_out/ab_base.json
  env "$1" timeout 600 python bench.py --secondary 0 --cpu-rows 0 --e2e-steps 0 --mode auto > gpurun_out/ab_var.json 2> gpurun_out/ab_var.err; echo -n "$1: "; show gpurun_out/ab_var.json
done
