mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "int8 or stream or upload or host" ) > gpurun_out/t_i8s.log 2>&1; tail -n 3 gpurun_out/t_i8s.log
python bench.py --secondary 0 --cpu-rows 0 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
"
tail -n 3 gpurun_out/bench_e2e.err
