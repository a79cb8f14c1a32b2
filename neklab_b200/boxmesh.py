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
