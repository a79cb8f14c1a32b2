"""Shared builders for the tests: the same case as an oracle `SEMesh` (numpy) and as a libnlk `Mesh`."""
from __future__ import annotations

import os

import numpy as np

from neklab_b200.boxmesh import box_mesh
from oracle import ops
from oracle.mesh import SEMesh, coords_from_corners
from oracle.stepper import NekVec, StepParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def warp2d(c):
    x, y = c[:, 0].copy(), c[:, 1].copy()
    c = c.copy()
    Lx, Ly = x.max() - x.min(), y.max() - y.min()
    c[:, 0] = x + 0.06 * Lx * np.sin(2 * np.pi * (y - y.min()) / Ly) * np.sin(np.pi * (x - x.min()) / Lx)
    c[:, 1] = y + 0.05 * Ly * np.sin(2 * np.pi * (x - x.min()) / Lx) * np.sin(np.pi * (y - y.min()) / Ly)
    return c


def warp3d(c):
    x, y, z = c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy()
    c = c.copy()
    sx = np.sin(np.pi * x / x.max()); sy = np.sin(np.pi * y / y.max()); sz = np.sin(np.pi * z / z.max())
    c[:, 0] = x + 0.05 * sx * sy * np.cos(2 * np.pi * z / z.max())      # periodic-compatible in z
    c[:, 1] = y + 0.04 * sx * sy * sz
    c[:, 2] = z + 0.05 * sz * sx * np.cos(0.7 * y)
    return c


def box_case(ndim=2, nel=(4, 3), n=6, lxd=9, warp=True, bc=None, periodic=None, cbc_t=None, hi=None):
    hi = hi or tuple(float(k) for k in nel)
    bm = box_mesh(nel, (0.0,) * ndim, hi, periodic=periodic, bc=bc,
                  warp=(warp2d if ndim == 2 else warp3d) if warp else None)
    coords = coords_from_corners(bm["corners"], n)
    om = SEMesh(coords, bm["vertex"], bm["cbc"], lxd, cbc_t=cbc_t)
    return om, bm, coords


def nlk_mesh(om: SEMesh, cbc_t=None):
    from neklab_b200 import api
    return api.Mesh(om.coords, om.vertex, om.cbc_v, om.m, cbc_t=om.cbc_t if cbc_t is None else cbc_t)


def cylinder_case():
    """examples/cylinder/stability/direct (Re=50, lx1=6, lxd=9) from the committed fixture."""
    z = np.load(os.path.join(GOLDEN, "cylinder_case.npz"))
    om = SEMesh(z["coords"], z["vertex"], z["cbc"], 9)
    bf = NekVec(om, 3)
    bf.v = [z["vel"][:, 0].copy(), z["vel"][:, 1].copy()]
    bf.pr = ops.map12(om, z["pr"])
    prm = StepParams(viscosity=1.0 / 50.0, torder=3, vtol=1e-9, ptol=1e-7)
    return om, bf, prm, z


def bfs_case():
    """examples/back_fstep/transient_growth (Re=600, lx1=6, lxd=9, bdf2, tau=18, filter 0.01/0.84; boundary codes from
    bfs.usr `usrdat`: wall 'W', inlet AND outlet 'v', free-slip 'SYM') from the committed fixture."""
    z = np.load(os.path.join(GOLDEN, "bfs_case.npz"))
    om = SEMesh(z["coords"], z["vertex"], z["cbc"], 9)
    bf = NekVec(om, 2)
    bf.v = [z["vel"][:, 0].copy(), z["vel"][:, 1].copy()]
    bf.pr = ops.map12(om, z["pr"])
    prm = StepParams(viscosity=1.0 / 600.0, torder=2, vtol=1e-8, ptol=1e-6, filter_weight=0.01, filter_cutoff=0.84)
    return om, bf, prm, z


def smooth_fields(om: SEMesh, k, seed=0):
    rng = np.random.default_rng(seed)
    x = om.coords
    out = []
    for i in range(k):
        f = np.sin(0.9 * x[:, 0] + 0.4 * i) * np.cos(0.7 * x[:, 1] - 0.3 * i)
        if om.ndim == 3:
            f = f * np.cos(0.5 * x[:, 2] + 0.2 * i)
        out.append(f + 0.05 * rng.standard_normal(f.shape))
    return out


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def poiseuille_case(n=8, lxd=12):
    """examples/poiseuille/stability/direct_alpha_1 (config C1): the reference's 10 x 12 wall-graded channel mesh
    (`poiseuille.re2` corners scaled by pi in x as `poiseuille.usr` usrdat does, `.ma2` vertex ids, x-periodic), U = 1 - y^2,
    Re = 7500 (`viscosity = -7500`), bdf2, tau = 1."""
    from oracle.mesh import coords_from_corners
    z = np.load(os.path.join(GOLDEN, "poiseuille_case.npz"))
    c = z["corners"].copy(); c[:, 0] *= np.pi
    coords = coords_from_corners(c, n)
    om = SEMesh(coords, z["vertex"], z["cbc"], lxd)
    bf = NekVec(om, 2)
    bf.v = [1.0 - coords[:, 1] ** 2, np.zeros_like(coords[:, 1])]
    prm = StepParams(viscosity=1.0 / 7500.0, torder=2, vtol=1e-10, ptol=1e-10)
    return om, bf, prm, z


def cheb(N):
    x = np.cos(np.pi * np.arange(N + 1) / N)
    c = np.hstack([2.0, np.ones(N - 1), 2.0]) * (-1.0) ** np.arange(N + 1)
    X = np.tile(x, (N + 1, 1)).T
    dX = X - X.T
    D = np.outer(c, 1.0 / c) / (dX + np.eye(N + 1))
    D -= np.diag(D.sum(axis=1))
    return D, x


def orr_sommerfeld_leading(R, N=120):
    """Leading temporal eigenvalue lambda (exp(lambda t), alpha = 1) of plane Poiseuille flow by Chebyshev collocation
    (Trefethen, Spectral Methods in MATLAB, p. 40) -- an anchor independent of the reference and of the oracle."""
    import scipy.linalg as sla
    D, x = cheb(N)
    D2 = (D @ D)[1:N, 1:N]
    S = np.diag(np.hstack([0.0, 1.0 / (1.0 - x[1:N] ** 2), 0.0]))
    D4 = (np.diag(1 - x ** 2) @ np.linalg.matrix_power(D, 4) - 8 * np.diag(x) @ np.linalg.matrix_power(D, 3) - 12 * D @ D) @ S
    D4 = D4[1:N, 1:N]
    I = np.eye(N - 1)
    A = (D4 - 2 * D2 + I) / R - 2j * I - 1j * np.diag(1 - x[1:N] ** 2) @ (D2 - I)
    B = D2 - I
    ee = sla.eigvals(A, B)
    ee = ee[np.isfinite(ee)]
    return ee[np.argmax(ee.real)]


def synth3d_window(half_width=3.0, layers=3, n=8, lxd=12):
    """A window of bench.py's synthetic 3-D config (the Re=50 cylinder mesh extruded periodically in z, lx1 = 8, lxd = 12):
    the 2-D elements whose centroid lies within `half_width` of the cylinder (`bench.window_case`: cut faces become 'v  ' or,
    downstream, 'O  '), `layers` periodic z-layers -- curved near-cylinder elements, the periodic direction and an outflow
    boundary are all present.  Returns (SEMesh, U list) for the oracle; `nlk_mesh(om)` gives the device mesh."""
    import bench
    coords, U, vertex, cbc3 = bench.extrude(bench.window_case(bench.cylinder_inputs(), half_width), n, layers)
    om = SEMesh(coords, vertex, cbc3, lxd)
    return om, [U[:, k].copy() for k in range(3)]
