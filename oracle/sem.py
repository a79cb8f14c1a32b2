"""ORACLE (test infrastructure, not product): 1-D spectral-element building blocks.

CPU/numpy restatement of the Nek5000 `speclib.f` objects the exptA hot path
uses.  Nek5000 is an un-vendored, un-pinned dependency of
the reference (`Nek5000_setup.sh:56-58`), so these follow the published
formulas (Legendre/GLL quadrature, Lagrange interpolation/derivative
matrices); the in-tree call sites that consume them are
`src/linops/neklab_linops.f90:343-362` (`local_grad3/2` with `dxm1`).

PARITY STATUS: unpinned by the reference's own tests except through the
known-answer tests on shipped fixtures (tests/test_oracle_kat.py) and the
single golden eigenvalue (`test/neklabTests.py:44`).
"""
from __future__ import annotations

import numpy as np


def legendre(n: int, x: np.ndarray):
    """P_n(x) and P_n'(x) by the three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    d0 = np.zeros_like(x)
    d1 = np.ones_like(x)
    for k in range(1, n):
        p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1)
        d2 = d0 + (2 * k + 1) * p1
        p0, p1 = p1, p2
        d0, d1 = d1, d2
    return p1, d1


def gll(n: int):
    """n Gauss-Lobatto-Legendre nodes and weights on [-1,1] (Nek `zwgll`)."""
    if n < 2:
        raise ValueError("GLL needs n>=2")
    N = n - 1
    x = -np.cos(np.pi * np.arange(n) / N)
    for _ in range(100):
        # roots of (1-x^2) P_N'(x): Newton on q(x)=P_N'(x) for interior points
        p, dp = legendre(N, x)
        # second derivative from Legendre ODE: (1-x^2)P'' - 2xP' + N(N+1)P = 0
        with np.errstate(divide="ignore", invalid="ignore"):
            ddp = (2 * x * dp - N * (N + 1) * p) / (1 - x * x)
        dx = np.zeros_like(x)
        dx[1:-1] = -dp[1:-1] / ddp[1:-1]
        x = x + dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # symmetrise
    p, _ = legendre(N, x)
    w = 2.0 / (N * (N + 1) * p * p)
    return x, w


def gl(n: int):
    """n Gauss-Legendre nodes and weights (Nek `zwgl`)."""
    x = -np.cos(np.pi * (np.arange(n) + 0.5) / n)
    for _ in range(100):
        p, dp = legendre(n, x)
        dx = -p / dp
        x = x + dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x = 0.5 * (x - x[::-1])
    _, dp = legendre(n, x)
    w = 2.0 / ((1 - x * x) * dp * dp)
    return x, w


def bary_weights(x: np.ndarray) -> np.ndarray:
    n = len(x)
    w = np.ones(n)
    for j in range(n):
        for k in range(n):
            if k != j:
                w[j] *= (x[j] - x[k])
    return 1.0 / w


def interp_matrix(xto: np.ndarray, xfrom: np.ndarray) -> np.ndarray:
    """I[i,j] = l_j(xto_i), l_j the Lagrange basis on xfrom (Nek `igllm`/`iglm`)."""
    xto = np.asarray(xto, float); xfrom = np.asarray(xfrom, float)
    bw = bary_weights(xfrom)
    m, n = len(xto), len(xfrom)
    out = np.zeros((m, n))
    for i in range(m):
        d = xto[i] - xfrom
        hit = np.where(d == 0.0)[0]
        if len(hit):
            out[i, hit[0]] = 1.0
        else:
            t = bw / d
            out[i] = t / t.sum()
    return out


def deriv_matrix(x: np.ndarray) -> np.ndarray:
    """D[i,j] = l_j'(x_i) on the nodes x themselves (Nek `dgll` / `dgllgl` family)."""
    x = np.asarray(x, float)
    n = len(x)
    bw = bary_weights(x)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = (bw[j] / bw[i]) / (x[i] - x[j])
        D[i, i] = -np.sum(D[i, np.arange(n) != i])
    return D


def deriv_interp_matrix(xto: np.ndarray, xfrom: np.ndarray) -> np.ndarray:
    """D12[i,j] = l_j'(xto_i), basis on xfrom (Nek `dxm12`)."""
    return interp_matrix(xto, xfrom) @ deriv_matrix(xfrom)


class Basis1D:
    """All 1-D operators for one (lx1, lxd) pair; lx2 = lx1-2 (PN-PN-2).

    Names follow Nek: z1/w1 (GLL, mesh 1), z2/w2 (GL, mesh 2), zd/wd (GL, dealias mesh),
    D (dxm1), I12 (ixm12), D12 (dxm12), I1d (GLL->fine GL), Dd (derivative on fine GL).
    """

    def __init__(self, lx1: int, lxd: int):
        self.n, self.m, self.q = lx1, lxd, lx1 - 2
        self.z1, self.w1 = gll(lx1)
        self.z2, self.w2 = gl(lx1 - 2)
        self.zd, self.wd = gl(lxd)
        self.D = deriv_matrix(self.z1)
        self.I12 = interp_matrix(self.z2, self.z1)
        self.D12 = deriv_interp_matrix(self.z2, self.z1)
        self.I21 = interp_matrix(self.z1, self.z2)
        self.I1d = interp_matrix(self.zd, self.z1)
        self.Dd = deriv_matrix(self.zd)
