"""ORACLE (test infrastructure, not product): spectral-element operators.

CPU/numpy restatement of the Nek5000 operator kernels on the exptA hot path
(SURVEY.md §2.3 K1/K3/K4/K5/K6, App. A.2): `axhelm` (hmholtz.f), `multd`/`opdiv`,
`cdtp`/`opgradt` (navier1.f), `convect_new`/`convect_adj` (convect.f), `cdabdtp`.
Nek5000 is un-vendored (`Nek5000_setup.sh:56-58`); the in-tree mirror of the
call ordering of `advabp` is `src/linops/neklab_linops.f90:285-312` and of
`local_grad3` + metric contraction `:343-362`.

PARITY STATUS: pinned by KAT-1/KAT-3 on the shipped cylinder base flow.
"""
from __future__ import annotations

import numpy as np

from .mesh import SEMesh, ax_r, ax_s, ax_t, tensor_apply

_AX = (ax_r, ax_s, ax_t)


def local_grad(mesh: SEMesh, u, D=None):
    """(u_r, u_s[, u_t]) -- Nek `local_grad2/3` (neklab_linops.f90:343-362)."""
    D = mesh.b.D if D is None else D
    return [_AX[k](D, u) for k in range(mesh.ndim)]


def axhelm(mesh: SEMesh, u, h1, h2):
    """w = h1 * (D^T G D) u + h2 * B u  (element-local, no dssum) -- Nek `axhelm`."""
    d = mesh.ndim
    D = mesh.b.D
    ur = local_grad(mesh, u)
    w = h2 * mesh.bm1 * u
    for k in range(d):
        t = sum(mesh.G[k][l] * ur[l] for l in range(d))
        w = w + h1 * _AX[k](D.T, t)
    return w


def axhelm_diag(mesh: SEMesh, h1, h2):
    """diag of the local Helmholtz operator (Nek `setprec`), before dssum.

    Nek's setprec keeps the G_kk terms, and for deformed elements adds the G_kl cross
    contributions only at ... -- we keep the exact diagonal of the tensor form, which for
    k != l contributes 2*G_kl*D_ii*D_jj at node (i,j).
    """
    d = mesh.ndim
    D = mesh.b.D
    D2 = D * D                      # D2[l,i] = D[l,i]^2 ; diag_i = sum_l G(l) D[l,i]^2
    out = h2 * mesh.bm1
    for k in range(d):
        out = out + h1 * _AX[k](D2.T, mesh.G[k][k])
    dd = np.diag(D)
    n = mesh.n
    # cross terms: d^2/du_i^2 of 2*G_kl*(D_k u)(D_l u) at the node itself
    for k in range(d):
        for l in range(k + 1, d):
            shape = [1, 1, 1, 1]
            fk = np.ones((1, 1, 1, 1))
            sk = [1, 1, 1, 1]; sk[3 - k] = n
            sl = [1, 1, 1, 1]; sl[3 - l] = n
            out = out + h1 * 2.0 * mesh.G[k][l] * dd.reshape(sk) * dd.reshape(sl)
    return out


def _interp_dir(mesh, k, M_deriv, M_interp, u):
    """apply M_deriv along direction k and M_interp along the others."""
    out = u
    for dd in range(mesh.ndim):
        out = _AX[dd](M_deriv if dd == k else M_interp, out)
    return out


def multd(mesh: SEMesh, u, c):
    """(D_c u) on mesh 2: w2 * sum_k r_{k,c}2 * d_k(u)|_GL  -- Nek `multd`."""
    b = mesh.b
    out = 0.0
    for k in range(mesh.ndim):
        out = out + mesh.rx2[k][c] * _interp_dir(mesh, k, b.D12, b.I12, u)
    return out * mesh.W2


def opdiv(mesh: SEMesh, u):
    """Weak divergence D u (list of velocity comps -> mesh-2 scalar) -- Nek `opdiv`."""
    return sum(multd(mesh, u[c], c) for c in range(mesh.ndim))


def cdtp(mesh: SEMesh, p, c):
    """(D_c^T p) on mesh 1 -- exact transpose of multd -- Nek `cdtp`."""
    b = mesh.b
    out = 0.0
    wp = p * mesh.W2
    for k in range(mesh.ndim):
        out = out + _interp_dir(mesh, k, b.D12.T, b.I12.T, wp * mesh.rx2[k][c])
    return out


def opgradt(mesh: SEMesh, p):
    return [cdtp(mesh, p, c) for c in range(mesh.ndim)]


def convect_new(mesh: SEMesh, u, C):
    """J^T W [(J C . rx) . grad(J u)] on the fine GL mesh -- Nek `convect_new`, ifcf=.false.

    u: scalar field on mesh 1; C: list of ndim convecting fields on mesh 1.
    Returns the *weak* convective term (already mass-weighted: equals bm1 * conv).
    """
    b = mesh.b
    d = mesh.ndim
    uf = tensor_apply(b.I1d, u, d)
    Cf = [tensor_apply(b.I1d, C[c], d) for c in range(d)]
    acc = 0.0
    for k in range(d):
        tr = sum(mesh.rxd[k][c] * Cf[c] for c in range(d))
        acc = acc + tr * _AX[k](b.Dd, uf)
    return tensor_apply(b.I1d.T, acc, d)


def convect_adj(mesh: SEMesh, U, c):
    """bdu_i = J^T W [ sum_j (J c_j) d(J U_j)/dx_i ] -- Nek `convect_adj`.

    U: list of base-flow comps, c: list of perturbation comps.  Returns list of ndim fields
    (weak form, mass-weighted).
    """
    b = mesh.b
    d = mesh.ndim
    cf = [tensor_apply(b.I1d, c[j], d) for j in range(d)]
    acc = [0.0] * d
    for j in range(d):
        Uf = tensor_apply(b.I1d, U[j], d)
        g = [_AX[k](b.Dd, Uf) for k in range(d)]
        for i in range(d):
            dUj_dxi = sum(mesh.rxd[k][i] * g[k] for k in range(d))
            acc[i] = acc[i] + cf[j] * dUj_dxi
    return [tensor_apply(b.I1d.T, acc[i], d) for i in range(d)]


def opbinv_masked(mesh: SEMesh, v, rho=1.0):
    """mask * binvm1 * dssum(v) / rho per component -- Nek `opbinv` (intype=1 path)."""
    return [mesh.vmask[c] * mesh.binvm1 * mesh.dssum(v[c]) / rho for c in range(mesh.ndim)]


def cdabdtp(mesh: SEMesh, p, rho=1.0):
    """E p = D (mask * B^-1 * dssum (D^T p)) -- Nek `cdabdtp`, intype=1."""
    return opdiv(mesh, opbinv_masked(mesh, opgradt(mesh, p), rho))


def ortho(mesh: SEMesh, p):
    """Remove the mean when the pressure operator has a null space -- Nek `ortho`."""
    if mesh.has_outflow:
        return p
    return p - p.sum() / p.size


def compute_cfl(mesh: SEMesh, U, dt):
    """Nek `compute_cfl`: max dt*(|u_r|/dr + |u_s|/ds + |u_t|/dt) with u_r = (u . rx)/J."""
    d, n = mesh.ndim, mesh.n
    z = mesh.b.z1
    dri = np.zeros(n)
    dri[0] = 1.0 / (z[1] - z[0]); dri[-1] = 1.0 / (z[-1] - z[-2])
    dri[1:-1] = 2.0 / (z[2:] - z[:-2])
    cfl = 0.0
    for k in range(d):
        uk = sum(U[c] * mesh.rx[k][c] for c in range(d)) / mesh.jac
        shp = [1, 1, 1, 1]; shp[3 - k] = n
        cfl = cfl + np.abs(uk) * dri.reshape(shp)
    return dt * float(cfl.max())


def map21(mesh: SEMesh, p2):
    """mesh-2 (GL) -> mesh-1 (GLL) interpolation (Nek `mappr`/`map21t` family) for fld output."""
    return tensor_apply(mesh.b.I21, p2, mesh.ndim)


def map12(mesh: SEMesh, p1):
    """mesh-1 -> mesh-2: evaluate at the GL points (exact inverse of map21 for fld pressures, KAT-2)."""
    return tensor_apply(mesh.b.I12, p1, mesh.ndim)
