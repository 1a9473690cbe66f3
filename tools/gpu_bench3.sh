#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
  timeout 600 python bench.py --secondary 0 --cpu-rows 0 --e2e-steps 0 > gpurun_out/b3_$rep.json 2> gpurun_out/b3_$rep.err
  python - gpurun_out/b3_$rep.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 2), round(sum(d["phases_ms"].values()), 2), d["clocks"])
PY
done
