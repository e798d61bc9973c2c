"""CPU check of the mathematics behind the blocked exchange recurrence (csrc/exchange.cu, sections 2b / 3d-3f), in numpy,
against the oracle's V (reference: quadratic_bosonic_exchange.cpp:73-99):

  * the forward recursion is the triangular linear system  W[v+1] = 1/(v+1) sum_{j<=v} c(j,v) W[j];
  * inside a 32-row block  u = G rho + omega_0 h  with  G = (I - diag(mu) T)^-1 diag(mu)  built by recursive halving,
    G = [[G11, 0], [G22 T21 G11, G22]], and h = G kin;
  * with per-(block,row) power-of-two scales every quantity is a plain double and all terms are non-negative.

The CUDA kernels are tested on the GPU (tests/test_gpu_exchange_blocked.py); this file pins the algorithm itself."""
import numpy as np

from pimd_b_b200.config import SimConfig
from tests.helpers import KELVIN, MEV, FEMTOSECOND, Oracle


def block_inverse(T, mu):
    """G = (I - diag(mu) T)^-1 diag(mu) for strictly lower-triangular T >= 0, by recursive halving."""
    n = len(mu)
    if n == 1:
        return np.array([[mu[0]]])
    h = n // 2
    G11 = block_inverse(T[:h, :h], mu[:h])
    G22 = block_inverse(T[h:, h:], mu[h:])
    G = np.zeros((n, n))
    G[:h, :h] = G11
    G[h:, h:] = G22
    G[h:, :h] = G22 @ (T[h:, :h] @ G11)
    return G


def test_blocked_forward_recurrence_reproduces_the_reference_potential():
    N, P = 100, 4
    cfg = SimConfig(nbeads=P, natoms=N, ndim=3, bosonic=True, fixcom=False, pbc=False, temperature=1.0 * KELVIN,
                    mass=1.0, size=2000.0, interaction="free", external="harmonic", ext_omega=3 * MEV,
                    thermostat="none", seed=1, dt=FEMTOSECOND)
    rng = np.random.default_rng(100)
    x = np.repeat(rng.normal(0.0, 60.0, size=(1, N, 3)), P, axis=0) + rng.normal(0.0, 6.0, size=(P, N, 3))
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.update_forces()
    V = orc.exchange("V")
    beta = orc.lib.orc_beta(orc.h) / P                       # exchange beta (bosonic_exchange_base.cpp:16-18)
    k = orc.lib.orc_spring_constant(orc.h)
    x1, xP = x[0], x[P - 1]
    A = np.concatenate([[0.0], np.cumsum(np.sum((x1[1:] - xP[:-1]) ** 2, axis=1))])
    d2 = np.sum((xP[None, :, :] - x1[:, None, :]) ** 2, axis=-1)          # d2[u][v]
    log2c = -(0.5 * beta * k) * (A[None, :] - A[:, None] + d2) / np.log(2.0)   # log2 c(u,v), valid for u <= v

    # W[s] = om[s] * 2^Ex[s]; accumulators of later rows as (mantissa, exponent) pairs via log2 (exact enough here)
    lw2 = np.zeros(N + 1)                                     # log2 W, filled block by block from the blocked solve
    nb = (N + 31) // 32
    for q in range(nb):
        lo, hi = 32 * q, min(N - 1, 32 * q + 31)
        n = hi - lo + 1
        rows = np.arange(lo, hi + 1)
        B = np.array([np.floor(log2c[lo:v + 1, v]).max() for v in rows])          # block scale per row
        K = np.zeros((n, n))
        for kk in range(n):
            K[:kk + 1, kk] = 2.0 ** (log2c[lo:lo + kk + 1, lo + kk] - B[kk])     # K[j][k] = c(lo+j, lo+k) 2^-B[k]
        mu = 2.0 ** B / (rows + 1.0)
        T = np.zeros((n, n))
        for kk in range(n):
            T[kk, :kk] = K[1:kk + 1, kk]                                          # T[k][k'] = K[k'+1][k]
        assert np.all(T >= 0) and np.all(mu > 0)
        G = block_inverse(T, mu)
        direct = np.linalg.solve(np.eye(n) - np.diag(mu) @ T, np.diag(mu))       # the same matrix, the textbook way
        assert np.allclose(G, direct, rtol=1e-12, atol=0)
        assert np.all(G[np.tril_indices(n)] > 0) and np.all(G[np.triu_indices(n, 1)] == 0)
        h = G @ K[0, :]
        E = int(np.floor(lw2[lo]))
        om0 = 2.0 ** (lw2[lo] - E)
        rho = np.array([np.sum(2.0 ** (log2c[:lo, lo + kk] + lw2[:lo] - E - B[kk])) if lo else 0.0 for kk in range(n)])
        u = G @ rho + om0 * h
        assert np.all(u > 2.0 ** -700) and np.all(u < 2.0 ** 300)                # the fast-path window of the kernels
        lw2[lo + 1:hi + 2] = np.log2(u) + E
    V_blocked = -lw2 * np.log(2.0) / beta
    assert np.max(np.abs(V_blocked - V)) <= 1e-10 * np.max(np.abs(V))
