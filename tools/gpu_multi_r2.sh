# multi-GPU checks of round 2 on N GPUs of one box:  bash tools/gpu_multi_r2.sh N [weak|strong|both] [test]
N=${1:-2}; WHAT=${2:-both}; T=${3:-}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo_${N}gpu.txt
if [ -n "$T" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/t_multi_r2_${N}gpu.log
fi
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "scaling", "n_gpus")}, d["dtype"][:70])
print("   e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"], 1), round(d["e2e"]["value"], 1), d["e2e"]["staging"]))
print("   phases", {k: round(v, 2) for k, v in d["phases_ms"].items()})
print("   fp64", d["fp64"] and (round(d["fp64"]["ms_per_step"], 2), d["fp64"]["e2e"] and round(d["fp64"]["e2e"]["ms_per_step"], 1)))
print("   per_rank", d.get("per_rank"))
print("   upload_only", d["e2e"] and d["e2e"].get("upload_only_ms"), d["e2e"] and d["e2e"].get("upload_only_GBps_all_gpus"))
PY
}
if [ "$WHAT" != "strong" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/bench_r2_weak_n$N.json 2> gpurun_out/bench_r2_weak_n$N.err
  tail -c 300 gpurun_out/bench_r2_weak_n$N.err; show gpurun_out/bench_r2_weak_n$N.json
fi
if [ "$WHAT" != "weak" ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --scaling strong \
      --rows-total ${ROWS_TOTAL:-2000000} --steps 3 --warmup 2 --e2e-steps 0 > gpurun_out/bench_r2_strong_n$N.json 2> gpurun_out/bench_r2_strong_n$N.err
  tail -c 300 gpurun_out/bench_r2_strong_n$N.err; show gpurun_out/bench_r2_strong_n$N.json
fi
