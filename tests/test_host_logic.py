"""Host-side logic of the multi-GPU path, on CPU (-m "not gpu"):
  * the row-shard layout used by bench.py / runtime.shard_rows
  * the row-sharded dataflow of SURVEY.md §8e (local A*S and A^T Q, all-reduce of the n x l partials and of the
    l x l Gram matrices) reproduces the single-process oracle -- run over world_size-2 gloo, with the ORACLE
    doing the arithmetic of each rank (this is a test of the sharding scheme, not of the product kernels).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    from randnla_b200.runtime import shard_rows
    for m in [1, 7, 8, 100, 2000, 200000, 2000001]:
        for w in [1, 2, 4, 8]:
            spans = [shard_rows(m, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(s[1] >= s[0] for s in spans)
            if m >= 4 * w:
                assert all(s[0] % 4 == 0 for s in spans)       # Philox row-quads never straddle ranks


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cholqr2_sharded(X_local, dist, sharded):
    """CholeskyQR2 with an all-reduced Gram matrix: the distributed orthonormalisation of SURVEY.md §8e."""
    import torch
    for _ in range(2):
        G = X_local.T @ X_local
        if sharded:
            t = torch.from_numpy(G); dist.all_reduce(t); G = t.numpy()
        R = np.linalg.cholesky(G).T
        X_local = np.linalg.solve(R.T, X_local.T).T
    return X_local


def _worker(rank, world, port, m, n, k, s, q, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    from randnla_b200.runtime import shard_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((m, 12)) @ rng.standard_normal((12, n)) + 1e-3 * rng.standard_normal((m, n)))
    r0, r1 = shard_rows(m, world, rank)
    Al = np.asfortranarray(A[r0:r1])
    l = k + s

    def allreduce(Z):
        t = torch.from_numpy(np.ascontiguousarray(Z)); dist.all_reduce(t); return t.numpy()

    done = 0
    if q % 2 == 0:
        S = orc.omega_fill(0, n, l, seed=0, stream=1)                         # replicated: regenerated from the seed
    else:
        Om = orc.omega_fill(0, r1 - r0, l, seed=0, stream=2, row_off=r0)      # row shard of Omega (m x l)
        S = allreduce(orc.gemm_tn(Al, Om)); done = 1
        S = _cholqr2_sharded(S, dist, False)
    while q - done >= 2:
        Y = _cholqr2_sharded(orc.gemm_nn(Al, S), dist, True); done += 1
        S = _cholqr2_sharded(allreduce(orc.gemm_tn(Al, Y)), dist, False); done += 1
    Q = _cholqr2_sharded(orc.gemm_nn(Al, S), dist, True)
    Bt = allreduce(orc.gemm_tn(Al, Q))
    sig = np.linalg.svd(Bt, compute_uv=False)[:k]
    if rank == 0:
        np.save(out, sig)
    dist.destroy_process_group()


@pytest.mark.parametrize("q", [2, 3])
def test_row_sharded_dataflow_matches_single_process_oracle(tmp_path, orc, q):
    import torch.multiprocessing as mp
    m, n, k, s = 203, 60, 8, 4
    out = str(tmp_path / "sig.npy")
    mp.spawn(_worker, args=(2, _free_port(), m, n, k, s, q, out), nprocs=2, join=True)
    sig = np.load(out)
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((m, 12)) @ rng.standard_normal((12, n)) + 1e-3 * rng.standard_normal((m, n)))
    _, S, _ = orc.rand_svd(A, k, 0.1, s, orc.make_opts(mode=0, num_passes=q))
    assert (np.abs(sig - np.diag(S)) / np.diag(S)).max() < 1e-10
