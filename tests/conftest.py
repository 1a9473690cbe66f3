import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/): the checker, never the thing under test."""
    from oracle import oracle as o
    o.load()
    return o


@pytest.fixture(scope="session")
def rb():
    """The product: randnla_b200 over librnla.so.  GPU tests only."""
    import randnla_b200 as m
    from randnla_b200 import runtime
    runtime.init(None)
    return m


# ---- seeded restatements of the reference's fixtures (src/test_assist.rs; the reference uses thread_rng()) ----
def rank_k_matrix(m, n, k, seed=0):
    """src/test_assist.rs:7-31: sum of k rank-1 updates u v^T with i.i.d. N(0,1) vectors."""
    rng = np.random.default_rng(seed)
    A = np.zeros((m, n))
    for _ in range(k):
        u = rng.standard_normal(m)
        v = rng.standard_normal(n)
        A += np.outer(u, v)
    return np.asfortranarray(A)


def random_matrix(rows, cols, seed=0):
    """src/test_assist.rs:124-129"""
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((rows, cols)))


def random_hermitian(n, seed=0):
    """src/test_assist.rs:132-140"""
    G = np.random.default_rng(seed).standard_normal((n, n))
    return np.asfortranarray(0.5 * (G + G.T))


def random_psd(n, seed=0):
    """src/test_assist.rs:143-151"""
    G = np.random.default_rng(seed).standard_normal((n, n))
    return np.asfortranarray(G @ G.T)


def lowrank_plus_noise(m, n, seed=0, k=20, gap=1e-2, noise=1e-9):
    """decaying signal with a gap after k, plus small noise."""
    rng = np.random.default_rng(seed)
    r0 = 2 * k
    U, _ = np.linalg.qr(rng.standard_normal((m, r0)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r0)))
    sig = np.concatenate([np.logspace(0, -2, k), np.full(r0 - k, 1e-2 * gap)])
    return np.asfortranarray((U * sig) @ V.T + noise * rng.standard_normal((m, n)) / np.sqrt(m)), sig


def subspace_angle(X, Y):
    """largest principal angle between range(X) and range(Y) (orthonormal columns), from its sine so that
    small angles are resolved below sqrt(eps)."""
    R = Y - X @ (X.T @ Y)
    s = np.linalg.svd(R, compute_uv=False)
    return float(np.arcsin(np.clip(s.max(), 0.0, 1.0)))
