"""Synthetic tensor-product box meshes (genbox analogue) for tests and the synthetic 3-D bench.

Produces what Nek's tool chain would hand the solver: element corner coordinates
(lexicographic), `.ma2`-style global vertex ids (1-based, periodic directions identified,
`examples/*/**.ma2` layout -- SURVEY App. C) and per-face boundary codes `cbc`
(preprocessor face order: 1:s=-1, 2:r=+1, 3:s=+1, 4:r=-1, 5:t=-1, 6:t=+1).
Host-side helper only (no device code).
"""
from __future__ import annotations

import numpy as np


def box_mesh(nel, lo, hi, periodic=None, bc=None, warp=None):
    """nel=(nx,ny[,nz]); bc = dict face_name->code with names 'xlo','xhi','ylo','yhi','zlo','zhi'.

    Returns dict(corners (E,ndim,2**ndim) lexicographic, vertex (E,2**ndim) int64, cbc (E,2*ndim) 'U3',
    pid (E,) recursive-bisection-like partition ids).
    """
    d = len(nel)
    periodic = periodic or [False] * d
    bc = bc or {}
    grids = [np.linspace(lo[a], hi[a], nel[a] + 1) for a in range(d)]
    nvx = [nel[a] if periodic[a] else nel[a] + 1 for a in range(d)]
    E = int(np.prod(nel))
    nv = 2 ** d
    corners = np.zeros((E, d, nv)); vertex = np.zeros((E, nv), dtype=np.int64)
    cbc = np.full((E, 2 * d), "E  ", dtype="U3")
    names = {0: ("xlo", "xhi"), 1: ("ylo", "yhi"), 2: ("zlo", "zhi")}
    face_of = {(0, 0): 3, (0, 1): 1, (1, 0): 0, (1, 1): 2, (2, 0): 4, (2, 1): 5}
    idx = np.indices(nel[::-1]).reshape(d, -1)[::-1]          # idx[a][e], x fastest
    for c in range(nv):
        off = [(c >> a) & 1 for a in range(d)]
        vid = np.zeros(E, dtype=np.int64)
        stride = 1
        for a in range(d):
            ia = idx[a] + off[a]
            corners[:, a, c] = grids[a][ia]
            vid += (ia % nvx[a]) * stride
            stride *= nvx[a]
        vertex[:, c] = vid + 1
    for a in range(d):
        for side in (0, 1):
            onb = idx[a] == (0 if side == 0 else nel[a] - 1)
            code = "P  " if periodic[a] else bc.get(names[a][side], "W  ")
            cbc[onb, face_of[(a, side)]] = code
    if warp is not None:
        corners = warp(corners)
    # simple partition ids: z-order blocks so that pid // k gives contiguous balanced chunks
    pid = (np.arange(E) * 1024 // E).astype(np.int64)
    return dict(corners=corners, vertex=vertex, cbc=cbc, pid=pid, ndim=d, nel=E)


def gll_points(n: int) -> np.ndarray:
    """n Gauss-Lobatto-Legendre nodes on [-1, 1] (Newton on (1-x^2) P'_{n-1}); host-side input generation only."""
    N = n - 1
    x = -np.cos(np.pi * np.arange(n) / N)
    for _ in range(100):
        p0 = np.ones_like(x); p1 = x.copy(); d0 = np.zeros_like(x); d1 = np.ones_like(x)
        for k in range(1, N):
            p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1); d2 = d0 + (2 * k + 1) * p1
            p0, p1, d0, d1 = p1, p2, d1, d2
        with np.errstate(divide="ignore", invalid="ignore"):
            dd = (2 * x * d1 - N * (N + 1) * p1) / (1 - x * x)
        dx = np.zeros_like(x); dx[1:-1] = -d1[1:-1] / dd[1:-1]
        x = x + dx
        if np.abs(dx).max() < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    return 0.5 * (x - x[::-1])


def lagrange_interp(xto: np.ndarray, xfrom: np.ndarray) -> np.ndarray:
    """I[i, j] = l_j(xto_i) for the Lagrange basis on xfrom."""
    out = np.zeros((len(xto), len(xfrom)))
    for j in range(len(xfrom)):
        lj = np.ones_like(xto)
        for k in range(len(xfrom)):
            if k != j:
                lj *= (xto - xfrom[k]) / (xfrom[j] - xfrom[k])
        out[:, j] = lj
    return out
