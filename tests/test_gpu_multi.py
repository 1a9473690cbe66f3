"""The row-sharded (one process per GPU, NCCL) drivers against the single-GPU result for the SAME global matrix, under torchrun.
Skipped on a box with fewer than two GPUs (the driver's round-end `-m gpu` run has one); run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` (log under profiles/)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _last_json(text):
    for line in reversed(text.strip().splitlines()):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise AssertionError("no JSON line in:\n" + text[-2000:])


@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_drivers_match_the_single_gpu_run(world):
    """rand_svd (reference src/lora_drivers.rs:49-68 over the passes of src/lora_helpers.rs:89, 95, 41, 21) with A row-sharded over
    `world` ranks -- A S and the integer-tensor-core passes local, A^T Q all-reduced, CholeskyQR Gram matrices all-reduced -- gives the
    singular values of the 1-GPU run to 1e-10 and an orthonormal U; the block sparse-sign sketch of the shards sums to the sketch of
    the whole; blendenpik, lsqr and cgls on the sharded system reproduce the 1-GPU iterates (tools/multi_gpu_check.py)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ",".join(str(i) for i in range(torch.cuda.device_count()))))
    one = subprocess.run([sys.executable, "tools/multi_gpu_check.py"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr[-3000:]
    ref = _last_json(one.stdout)
    assert ref["n_gpus"] == 1 and ref["orth"] < 1e-12
    many = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                           "--master-port", str(_free_port()), "tools/multi_gpu_check.py"], cwd=ROOT, env=env, capture_output=True, text=True,
                          timeout=900)
    assert many.returncode == 0, many.stderr[-3000:]
    out = _last_json(many.stdout)
    assert out["n_gpus"] == world
    assert out["max_rel_sigma_diff_vs_1gpu"] < 1e-10 and out["orth"] < 1e-12
    assert out["saso_block_max_abs_diff_vs_1gpu"] < 1e-12 and out["blendenpik_x_rel_diff_vs_1gpu"] < 1e-8
    assert out["lsqr_x_rel_diff_vs_1gpu"] < 1e-9 and out["cgls_plain_x_rel_diff_vs_1gpu"] < 1e-9
    assert out["solver_iterations_read_A_once"] and ref["solver_iterations_read_A_once"]      # the sharded one-pass dataflow ran
    assert out["pass"], out
