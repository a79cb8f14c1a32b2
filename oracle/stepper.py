"""ORACLE (test infrastructure, not product): PN-PN-2 perturbation time stepper + exptA.

CPU/numpy restatement of the path `exptA_matvec` drives
(`src/linops/exponential_propagator.f90:15-60`): `setup_linear_solver`
(`src/neklab_nek_setup.f90:39-247`, dt/nsteps rule :193-224, tolerances :227-230),
then `nsteps` x Nek5000 `nek_advance` in perturbation mode (`fluidp`/`perturbv`,
`makefp`, `advabp`/`advabp_adjoint`, `makextp`, `makebdfp`, `lagfieldp`,
`cresvipp`, `ophinv`->`hmholtz`->`cggo`, `incomprp`->`uzawa_gmres`, `heatp`->`cdscalp`),
with the multistep restart protocol of `exptA_get_rst`/`exptA_compute_rst`
(`exponential_propagator.f90:109-142`).  Nek5000 is an un-vendored dependency
(`Nek5000_setup.sh:56-58`, branch master, unpinned); its routines are restated
from their published algorithm.

PARITY STATUS: parity unpinned at 1e-10; pinned to the reference's single golden
value |lambda_1| = 1.0156 +- 1e-4 (`test/neklabTests.py:44`) by
tests/golden/cylinder_eig_oracle.json (made by oracle/make_golden.py).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional

import numpy as np

from . import ops
from .mesh import SEMesh


# --------------------------------------------------------------------------- time scheme
def bdf_coeffs(nbd: int):
    """Nek `setbd`, constant dt. bd[0] multiplies u^{n+1}; bd[i] (i>=1) multiply u^{n+1-i} on the rhs."""
    return {1: [1.0, 1.0], 2: [1.5, 2.0, -0.5], 3: [11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0]}[nbd]


def ab_coeffs(nab: int, nbd: int):
    """Nek `setabbd`, constant dt (modified AB coefficients paired with BDF-nbd)."""
    if nab == 1:
        return [1.0, 0.0, 0.0]
    if nab == 2:
        if nbd in (1, 2):
            return [1.5, -0.5, 0.0]
        return [2.0, -1.0, 0.0]
    if nbd == 1:
        ab3 = 0.5 * (0.5 + 1.0 / 3.0)       # DTE*(0.5*DTB + DTC/3), DTE=.5, DTB=DTC=1
        ab2 = -0.5 - ab3 * 2.0
        return [1.0 - ab2 - ab3, ab2, ab3]
    if nbd == 2:
        ab3 = 2.0 / 3.0
        ab2 = -1.0 - ab3 * 2.0
        return [1.0 - ab2 - ab3, ab2, ab3]   # (8/3, -7/3, 2/3)
    return [3.0, -3.0, 1.0]


@dataclasses.dataclass
class StepParams:
    viscosity: float                      # h1 = vdiff (=1/Re)
    density: float = 1.0                  # vtrans
    torder: int = 3                       # |param(27)|
    vtol: float = 1e-9                    # param(22) -> TOLHDF
    ptol: float = 1e-7                    # param(21) -> TOLPDF
    ifheat: bool = False
    conductivity: float = 1.0
    rhocp: float = 1.0
    ttol: float = 1e-9
    buoyancy: tuple = (0.0, 0.0, 0.0)     # f_c += buoyancy[c] * T'  (rayBen.usr:75-105)
    filter_weight: float = 0.0            # param(103)
    filter_cutoff: float = 1.0            # filterCutoffRatio
    cg_maxit: int = 1000
    gmres_maxit: int = 100                # Nek: iter.lt.100
    lgmres: int = 30
    pressure_solver: str = "gmres"        # 'gmres' | 'direct'
    helm_solver: str = "cg"               # 'cg' | 'direct'


# --------------------------------------------------------------------------- solvers
def cggo(mesh: SEMesh, f, h1, h2, mask, tol, maxit, stats=None):
    """Jacobi-PCG on (h1 A + h2 B) -- Nek `hmholtz`+`cggo` (rhs already local, un-dssum'ed)."""
    rhs = mask * mesh.dssum(f)
    diag = mesh.dssum(ops.axhelm_diag(mesh, h1, h2))
    dinv = 1.0 / diag
    x = np.zeros_like(rhs)
    if np.abs(rhs).max() == 0.0:
        return x
    r = rhs.copy()
    p = np.zeros_like(rhs)
    rtz1 = 1.0
    mult = mesh.vmult
    wbn = mult * mesh.binvm1
    it = 0
    for it in range(1, maxit + 1):
        z = r * dinv
        rtz2 = rtz1
        rtz1 = float((z * r * mult).sum())
        rbn2 = math.sqrt(float((r * r * wbn).sum()) / mesh.volvm1)
        if rbn2 <= tol:
            it -= 1
            break
        beta = 0.0 if it == 1 else rtz1 / rtz2
        p = z + beta * p
        w = mask * mesh.dssum(ops.axhelm(mesh, p, h1, h2))
        rho = float((w * p * mult).sum())
        alpha = rtz1 / rho
        x += alpha * p
        r -= alpha * w
    if stats is not None:
        stats["cg_iters"] = stats.get("cg_iters", 0) + it
    return x


def uzawa_gmres(mesh: SEMesh, res, apply_E, precond, tol, maxit=100, m=30, stats=None):
    """Right-preconditioned, mass-scaled restarted FGMRES -- Nek `uzawa_gmres` (gmres.f)."""
    ml = np.sqrt(1.0 / mesh.bm2)
    mu = np.sqrt(mesh.bm2)
    norm_fac = 1.0 / math.sqrt(mesh.volvm2)
    x = np.zeros_like(res)
    it = 0
    conv = False
    dot = lambda a, b: float((a * b).sum())
    while not conv and it < maxit:
        if it == 0:
            r = ml * res
        else:
            r = ml * (res - apply_E(x))
        gamma = [math.sqrt(dot(r, r))]
        if gamma[0] == 0.0:
            break
        V = [r / gamma[0]]
        Z = []
        H = np.zeros((m + 1, m))
        c = np.zeros(m); s = np.zeros(m)
        j = 0
        for j in range(1, m + 1):
            it += 1
            z = ops.ortho(mesh, precond(mu * V[j - 1]))
            Z.append(z)
            w = ml * apply_E(z)
            h = np.array([dot(w, V[i]) for i in range(j)])
            for i in range(j):
                w = w - h[i] * V[i]
            H[:j, j - 1] = h
            for i in range(j - 1):
                t = H[i, j - 1]
                H[i, j - 1] = c[i] * t + s[i] * H[i + 1, j - 1]
                H[i + 1, j - 1] = -s[i] * t + c[i] * H[i + 1, j - 1]
            alpha = math.sqrt(dot(w, w))
            if alpha == 0.0:
                conv = True
                break
            l = math.sqrt(H[j - 1, j - 1] ** 2 + alpha ** 2)
            c[j - 1] = H[j - 1, j - 1] / l
            s[j - 1] = alpha / l
            H[j - 1, j - 1] = l
            gamma.append(-s[j - 1] * gamma[j - 1])
            gamma[j - 1] = c[j - 1] * gamma[j - 1]
            rnorm = abs(gamma[j]) * norm_fac
            if rnorm < tol:
                conv = True
                break
            if j == m or it >= maxit:
                break
            V.append(w / alpha)
        # back substitution
        cc = np.zeros(j)
        for k in range(j - 1, -1, -1):
            t = gamma[k]
            for i in range(j - 1, k, -1):
                t -= H[k, i] * cc[i]
            cc[k] = t / H[k, k]
        for i in range(j):
            x = x + cc[i] * Z[i]
    if stats is not None:
        stats["gmres_iters"] = stats.get("gmres_iters", 0) + it
    return ops.ortho(mesh, x)


# --------------------------------------------------------------------------- sparse assembly (direct mode)
def _local_deriv_blocks(mesh: SEMesh, c: int):
    """Dense per-element blocks of multd(., c): shape (E, q^d, n^d)."""
    b = mesh.b
    d = mesh.ndim
    E = mesh.E
    if d == 2:
        # block[(jq,iq),(j,i)]
        t0 = np.einsum("eqp,qj,pi->eqpji", mesh.rx2[0][c][:, 0], b.I12, b.D12)
        t1 = np.einsum("eqp,qj,pi->eqpji", mesh.rx2[1][c][:, 0], b.D12, b.I12)
        blk = (t0 + t1) * mesh.W2[0, 0][None, :, :, None, None]
        return blk.reshape(E, mesh.q ** 2, mesh.n ** 2)
    t0 = np.einsum("erqp,rk,qj,pi->erqpkji", mesh.rx2[0][c], b.I12, b.I12, b.D12)
    t1 = np.einsum("erqp,rk,qj,pi->erqpkji", mesh.rx2[1][c], b.I12, b.D12, b.I12)
    t2 = np.einsum("erqp,rk,qj,pi->erqpkji", mesh.rx2[2][c], b.D12, b.I12, b.I12)
    blk = (t0 + t1 + t2) * mesh.W2[0][None, :, :, :, None, None, None]
    return blk.reshape(E, mesh.q ** 3, mesh.n ** 3)


def assemble_E(mesh: SEMesh, rho=1.0):
    """Sparse E = D (mask B^-1 QQ^T) D^T on the pressure dofs (scipy CSC)."""
    import scipy.sparse as sp
    E = mesh.E
    nq, nn = mesh.q ** mesh.ndim, mesh.n ** mesh.ndim
    Q = sp.csr_matrix((np.ones(E * nn), (np.arange(E * nn), mesh.gidx)), shape=(E * nn, mesh.nglob))
    out = None
    for c in range(mesh.ndim):
        blk = _local_deriv_blocks(mesh, c)
        Dc = sp.bsr_matrix((blk, np.arange(E), np.arange(E + 1)), shape=(E * nq, E * nn)).tocsr()
        G = Dc @ Q                                           # (Np x nglob): D_c Q
        wloc = (mesh.vmask[c] * mesh.binvm1).ravel() / rho
        wg = np.zeros(mesh.nglob); wg[mesh.gidx] = wloc      # identical across copies
        t = G @ sp.diags(wg) @ G.T
        out = t if out is None else out + t
    return out.tocsc()


def assemble_H(mesh: SEMesh, h1, h2, mask):
    """Sparse assembled Helmholtz operator on unmasked unique dofs; returns (H, free_idx, Q)."""
    import scipy.sparse as sp
    d, n, E = mesh.ndim, mesh.n, mesh.E
    D = mesh.b.D
    nn = n ** d
    I = np.eye(n)
    if d == 2:
        Dk = [np.kron(I, D), np.kron(D, I)]
    else:
        Dk = [np.kron(I, np.kron(I, D)), np.kron(I, np.kron(D, I)), np.kron(D, np.kron(I, I))]
    blk = np.zeros((E, nn, nn))
    for k in range(d):
        for l in range(d):
            g = mesh.G[k][l].reshape(E, nn)
            blk += h1 * np.einsum("ak,ea,al->ekl", Dk[k], g, Dk[l], optimize=True)
    blk[:, np.arange(nn), np.arange(nn)] += h2 * mesh.bm1.reshape(E, nn)
    Hl = sp.bsr_matrix((blk, np.arange(E), np.arange(E + 1)), shape=(E * nn, E * nn)).tocsr()
    Q = sp.csr_matrix((np.ones(E * nn), (np.arange(E * nn), mesh.gidx)), shape=(E * nn, mesh.nglob))
    Hg = (Q.T @ Hl @ Q).tocsr()
    mg = np.zeros(mesh.nglob); mg[mesh.gidx] = mask.ravel()
    free = np.where(mg > 0)[0]
    return Hg[free][:, free].tocsc(), free, Q


# --------------------------------------------------------------------------- filter
def filter_matrix(b, weight, cutoff_ratio):
    """Nek `q_filter` 1-D operator: F = (1-w) I + w V diag(sigma) V^-1 (Legendre-modal ramp).

    nmodes damped = n - int(n*cutoff_ratio + 0.01) ... (Nek: `ncut`), quadratic ramp
    amp = wght*(k/ncut)^2 on the top ncut modes.
    """
    from .sem import legendre
    n = b.n
    ncut = n - int(n * cutoff_ratio + 0.01) if cutoff_ratio < 1.0 else 0
    if ncut <= 0 or weight <= 0:
        return np.eye(n)
    # Nek build_new_filter: bubble-function basis phi_k = L_k - L_{k-2} (k>=2), phi_0=L_0, phi_1=L_1
    V = np.zeros((n, n))
    for k in range(n):
        pk, _ = legendre(k, b.z1)
        V[:, k] = pk
        if k >= 2:
            pk2, _ = legendre(k - 2, b.z1)
            V[:, k] = pk - pk2
    diag = np.ones(n)
    k0 = n - ncut
    for k in range(k0, n):
        kk = k + 1 - k0
        amp = weight * (kk * kk) / (ncut * ncut)
        diag[k] = 1.0 - amp
    return V @ np.diag(diag) @ np.linalg.inv(V)


# --------------------------------------------------------------------------- the stepper
class PertStepper:
    """State machine equivalent to Nek's COMMON blocks for one perturbation (lpert=1)."""

    def __init__(self, mesh: SEMesh, prm: StepParams, precond: Optional[Callable] = None):
        self.mesh, self.prm = mesh, prm
        self.d = mesh.ndim
        self.adjoint = False
        self.nonlinear = False                               # ifpert = .false.: Nek `fluid`/`plan3` instead of `fluidp`
        self.U = [np.zeros_like(mesh.bm1) for _ in range(self.d)]
        self.T = np.zeros_like(mesh.bm1)
        self.forcing = None                                  # neklab_ffx/y/z (neklab_nek_forcing.f90:19-21)
        self.stats = {}
        self._precond = precond
        self._lu_E = None
        self._lu_H = {}
        self.F1d = filter_matrix(mesh.b, prm.filter_weight, prm.filter_cutoff)
        self.dt = None
        self.nsteps = None

    # -- src/neklab_nek_setup.f90:193-224
    def setup(self, tau: float, cfl_limit: float = 0.5, transpose: bool = False):
        self.adjoint = transpose
        umax = max(np.abs(u).max() for u in self.U)
        if umax == 0.0:
            if self.dt is None:
                raise ValueError("zero base flow: dt must be preset (recompute_dt disabled)")
            self.nsteps = int(math.ceil(tau / self.dt))
        else:
            ctarg = ops.compute_cfl(self.mesh, self.U, 1.0)
            dt = cfl_limit / ctarg
            self.nsteps = int(math.ceil(tau / dt))
            self.dt = tau / self.nsteps
        return self.dt, self.nsteps

    def set_state(self, v, p, t=None):
        m = self.mesh
        self.vp = [x.copy() for x in v]
        self.prp = p.copy()
        self.tp = t.copy() if t is not None else np.zeros_like(m.bm1)

    def reset_history(self):
        m = self.mesh
        z = lambda: np.zeros_like(m.bm1)
        self.vlag = [[z() for _ in range(self.d)] for _ in range(2)]
        self.exx1 = [z() for _ in range(self.d)]
        self.exx2 = [z() for _ in range(self.d)]
        self.prlag = np.zeros_like(m.bm2)
        self.tlag = [z(), z()]
        self.vgradt1 = z()
        self.vgradt2 = z()

    # -- pressure operator / solver
    def _apply_E(self, p):
        return ops.cdabdtp(self.mesh, p, self.prm.density)

    def _solve_pressure(self, rhs, tol):
        m = self.mesh
        if self.prm.pressure_solver == "direct":
            import scipy.sparse.linalg as spl
            N = rhs.size
            if self._lu_E is None:
                E = assemble_E(m, self.prm.density)
                lu = spl.splu(E)
                # no outflow: E has the constants in its (near-)null space -- exactly on straight-sided meshes, to quadrature
                # error on curved ones (bfs: |E 1| = 6e-5) -- and Nek's uzawa_gmres only ever moves in mean-free directions
                # (`ortho`).  The converged limit of that iteration is the solution of the bordered system
                #   [E 1; 1^T 0] [x; mu] = [rhs; 0],
                # solved here by block elimination with the LU of E (w = E^-1 1) plus iterative refinement on the bordered residual.
                w = None if m.has_outflow else lu.solve(np.ones(N))
                self._lu_E = (E, lu, w)
            E, lu, w = self._lu_E
            b = rhs.ravel()
            if w is None:
                return lu.solve(b).reshape(rhs.shape)
            sw = w.sum()
            x = np.zeros(N); mu = 0.0
            for _ in range(4):
                r = b - E @ x - mu; s_ = -x.sum()
                y = lu.solve(r)
                dmu = (y.sum() - s_) / sw
                x += y - dmu * w; mu += dmu
            return x.reshape(rhs.shape)
        pre = self._precond if self._precond is not None else (lambda r: r / m.bm2)
        return uzawa_gmres(m, rhs, self._apply_E, pre, tol, self.prm.gmres_maxit, self.prm.lgmres, self.stats)

    def _solve_helm(self, f, h1, h2, mask, tol, key):
        m = self.mesh
        if self.prm.helm_solver == "direct":
            k = (key, h1, h2)
            if k not in self._lu_H:
                import scipy.sparse.linalg as spl
                H, free, Q = assemble_H(m, h1, h2, mask)
                self._lu_H = {k: (spl.splu(H), free, Q)}.copy() | {kk: vv for kk, vv in self._lu_H.items() if kk[0] != key}
            lu, free, Q = self._lu_H[k]
            rg = Q.T @ f.ravel()
            xg = np.zeros(m.nglob)
            xg[free] = lu.solve(rg[free])
            return xg[m.gidx].reshape(f.shape)
        return cggo(m, f, h1, h2, mask, tol, self.prm.cg_maxit, self.stats)

    # -- one nek_advance in perturbation mode
    def advance(self, istep: int):
        m, prm, d = self.mesh, self.prm, self.d
        dt = self.dt
        nbd = min(istep, prm.torder)
        nab = min(istep, 3)
        bd = bdf_coeffs(nbd)
        ab = ab_coeffs(nab, nbd)
        rho = prm.density
        B = m.bm1
        # ---------------- igeom = 1 : makefp, lagfieldp
        # makeufp: body force * bm1 (userf: Boussinesq + neklab_forcing)
        bf = []
        for c in range(d):
            f = np.zeros_like(B)
            if prm.ifheat and prm.buoyancy[c] != 0.0:
                f = f + prm.buoyancy[c] * self.tp
            if self.forcing is not None:
                f = f + self.forcing[c]
            bf.append(f * B)
        # advab (nonlinear, Nek plan3/makef) | advabp / advabp_adjoint
        if self.nonlinear:
            for c in range(d):
                bf[c] = bf[c] - rho * ops.convect_new(m, self.vp[c], self.vp)       # u.grad u_c
        elif not self.adjoint:
            for c in range(d):
                bf[c] = bf[c] - rho * ops.convect_new(m, self.U[c], self.vp)     # u'.grad U_c
                bf[c] = bf[c] - rho * ops.convect_new(m, self.vp[c], self.U)     # U.grad u'_c
        else:
            adj = ops.convect_adj(m, self.U, self.vp)                            # (grad U)^T u'
            for c in range(d):
                bf[c] = bf[c] - rho * adj[c]
                bf[c] = bf[c] + rho * ops.convect_new(m, self.vp[c], self.U)
        # makextp
        for c in range(d):
            ta = ab[1] * self.exx1[c] + ab[2] * self.exx2[c]
            self.exx2[c] = self.exx1[c]
            self.exx1[c] = bf[c]
            bf[c] = ab[0] * bf[c] + ta
        # makebdfp
        for c in range(d):
            tb = bd[1] * B * self.vp[c]
            for ilag in range(2, nbd + 1):
                tb = tb + bd[ilag] * B * self.vlag[ilag - 2][c]
            bf[c] = bf[c] + tb * (rho / dt)
        # temperature igeom=1: makeqp (convabp, makeabqp, makebdqp), lagscalp
        if prm.ifheat:
            rcp = prm.rhocp
            bq = -rcp * ops.convect_new(m, self.T, self.vp) - rcp * ops.convect_new(m, self.tp, self.U)
            if self.adjoint:
                raise NotImplementedError("adjoint Boussinesq step is out of scope (no reference config uses it)")
            ta = ab[1] * self.vgradt1 + ab[2] * self.vgradt2
            self.vgradt2 = self.vgradt1
            self.vgradt1 = bq
            bq = ab[0] * bq + ta
            tb = bd[1] * B * self.tp
            for ilag in range(2, nbd + 1):
                tb = tb + bd[ilag] * B * self.tlag[ilag - 2]
            bq = bq + tb * (rcp / dt)
            self.tlag[1] = self.tlag[0]
            self.tlag[0] = self.tp.copy()
        # lagfieldp
        self.vlag[1] = self.vlag[0]
        self.vlag[0] = [x.copy() for x in self.vp]
        # ---------------- igeom = 2 : velocity
        h1 = prm.viscosity
        h2 = rho * bd[0] / dt
        # cresvipp (perturbation Dirichlet values are zero: bcdirvc == mask); cresvif keeps the inhomogeneous values
        if not self.nonlinear:
            self.vp = [m.vmask[c] * self.vp[c] for c in range(d)]
        if nbd == 3:
            pext = 2.0 * self.prp - self.prlag
        else:
            pext = self.prp
        gp = ops.opgradt(m, pext)
        res = [gp[c] + bf[c] - ops.axhelm(m, self.vp[c], h1, h2) for c in range(d)]
        # ophinv
        for c in range(d):
            dv = self._solve_helm(res[c], h1, h2, m.vmask[c], prm.vtol, "v%d" % c)
            self.vp[c] = self.vp[c] + dv
        # incomprp
        bdti = -bd[0] / dt
        dp = ops.ortho(m, bdti * ops.opdiv(m, self.vp))
        scaledt = dt / bd[0]
        dp = self._solve_pressure(dp * scaledt, prm.ptol) / scaledt
        self.prlag = self.prp.copy()                       # lagpresp
        self.prp = pext + dp                               # add3(up, prextr, dp)
        w = ops.opbinv_masked(m, ops.opgradt(m, dp), rho)
        for c in range(d):
            self.vp[c] = self.vp[c] + (dt / bd[0]) * w[c]
        # ---------------- igeom = 2 : temperature (cdscalp)
        if prm.ifheat:
            h1t = prm.conductivity
            h2t = prm.rhocp * bd[0] / dt
            self.tp = m.tmask * self.tp
            rt = bq - ops.axhelm(m, self.tp, h1t, h2t)
            dT = self._solve_helm(rt, h1t, h2t, m.tmask, prm.ttol, "t")
            self.tp = self.tp + dT
        # ---------------- q_filter
        if prm.filter_weight > 0:
            from .mesh import tensor_apply
            self.vp = [tensor_apply(self.F1d, v, d) for v in self.vp]
            if prm.ifheat:
                self.tp = tensor_apply(self.F1d, self.tp, d)


# --------------------------------------------------------------------------- nek_dvector
class NekVec:
    """`nek_dvector` semantics (src/vectors/neklab_vectors.f90:26-50, real_vectors.f90).

    RST_MODE 0 (default) reproduces `nek_daxpby` as written (rst slots += alpha * vec's CURRENT fields,
    real_vectors.f90:186-200); RST_MODE 1 is the consistent variant (rst slots += alpha * vec's rst fields),
    kept only to document the effect on the golden eigenvalue.
    """
    RST_MODE = 0

    def __init__(self, mesh: SEMesh, torder: int = 3, ifheat: bool = False):
        self.mesh, self.torder, self.ifheat = mesh, torder, ifheat
        d = mesh.ndim
        self.v = [np.zeros_like(mesh.bm1) for _ in range(d)]
        self.pr = np.zeros_like(mesh.bm2)
        self.theta = np.zeros_like(mesh.bm1)
        self.nrst = 0
        self.rst = [None] * (torder - 1)          # each: (v list, pr, theta)

    def copy(self):
        o = NekVec(self.mesh, self.torder, self.ifheat)
        o.v = [x.copy() for x in self.v]; o.pr = self.pr.copy(); o.theta = self.theta.copy()
        o.nrst = self.nrst
        o.rst = [None if r is None else ([x.copy() for x in r[0]], r[1].copy(), r[2].copy()) for r in self.rst]
        return o

    def zero(self):                                # real_vectors.f90:37-50
        for x in self.v:
            x[:] = 0
        self.pr[:] = 0; self.theta[:] = 0
        self.nrst = 0; self.rst = [None] * (self.torder - 1)

    def _rst_slot(self, i):
        if self.rst[i] is None:
            m = self.mesh
            self.rst[i] = ([np.zeros_like(m.bm1) for _ in range(m.ndim)], np.zeros_like(m.bm2), np.zeros_like(m.bm1))
        return self.rst[i]

    def scal(self, alpha):                         # :125-160 (pressure and rst included)
        for x in self.v:
            x *= alpha
        self.pr *= alpha
        if self.ifheat:
            self.theta *= alpha
        for i in range(self.nrst):
            rv, rp, rt = self._rst_slot(i)
            for x in rv:
                x *= alpha
            rp *= alpha
            if self.ifheat:
                rt *= alpha

    def axpby(self, alpha, vec: "NekVec", beta):   # :162-206  self = alpha*vec + beta*self
        self.scal(beta)
        for c in range(len(self.v)):
            self.v[c] += alpha * vec.v[c]
        self.pr += alpha * vec.pr
        if self.ifheat:
            self.theta += alpha * vec.theta
        for i in range(self.nrst):                 # quirk :186-200 -- adds vec's CURRENT fields
            rv, rp, rt = self._rst_slot(i)
            if NekVec.RST_MODE == 1 and vec.nrst > i:
                sv, sp, st_ = vec.rst[i]
            else:
                sv, sp, st_ = vec.v, vec.pr, vec.theta
            for c in range(len(self.v)):
                rv[c] += alpha * sv[c]
            rp += alpha * sp
            if self.ifheat:
                rt += alpha * st_

    def dot(self, vec: "NekVec") -> float:         # :208-233  bm1-weighted, pressure excluded
        B = self.mesh.bm1
        a = 0.0
        for c in range(len(self.v)):
            a += float((self.v[c] * vec.v[c] * B).sum())
        if self.ifheat:
            a += float((self.theta * vec.theta * B).sum())
        return a

    def norm(self):
        return math.sqrt(self.dot(self))

    def size(self):                                # :235-247
        n = len(self.v) * self.v[0].size + self.pr.size
        if self.ifheat:
            n += self.theta.size
        return n

    def save_rst(self, state: "NekVec", irst: int):   # :249-291 (irst 1-based)
        self.rst[irst - 1] = ([x.copy() for x in state.v], state.pr.copy(), state.theta.copy())
        self.nrst = max(self.nrst, irst)

    def get_rst(self, irst: int) -> "NekVec":         # :293-333
        o = NekVec(self.mesh, self.torder, self.ifheat)
        rv, rp, rt = self._rst_slot(irst - 1)
        o.v = [x.copy() for x in rv]; o.pr = rp.copy(); o.theta = rt.copy()
        return o


def seeded_field(mesh: SEMesh, seed: int, ifheat=False, torder=3) -> NekVec:
    """Deterministic C0, BC-satisfying test vector (replaces compiler-dependent `nek_drand`, N5)."""
    rng = np.random.default_rng(seed)
    v = NekVec(mesh, torder, ifheat)
    x = mesh.coords
    for c in range(mesh.ndim):
        f = np.sin(0.7 * x[:, 0] + 0.3 * c) * np.cos(0.9 * x[:, 1] - 0.2 * c) + 0.1 * rng.standard_normal(mesh.bm1.shape)
        f = mesh.dssum(f) * mesh.vmult
        v.v[c] = mesh.vmask[c] * f
    if ifheat:
        f = np.cos(0.5 * x[:, 0]) * np.sin(1.1 * x[:, 1]) + 0.1 * rng.standard_normal(mesh.bm1.shape)
        v.theta = mesh.tmask * (mesh.dssum(f) * mesh.vmult)
    return v


# --------------------------------------------------------------------------- nonlinear flow map / Newton
def nonlinear_map(st: PertStepper, x: NekVec, tau: float, cfl_limit: float = 0.4) -> NekVec:
    """`nek_system%response` (src/systems/fixed_point.f90:4-40): F_tau(X) - X with the nonlinear stepper."""
    st.nonlinear = True; st.adjoint = False
    st.U = [v.copy() for v in x.v]                       # CFL of the state itself (setup_nonlinear_solver, recompute_dt)
    st.setup(tau, cfl_limit, False)
    st.set_state(x.v, x.pr, x.theta); st.reset_history()
    for istep in range(1, st.nsteps + 1):
        st.advance(istep)
    out = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)
    out.v = [a.copy() for a in st.vp]; out.pr = st.prp.copy(); out.theta = st.tp.copy()
    out.axpby(-1.0, x, 1.0)                              # vec_out%sub(vec_in)
    st.nonlinear = False
    return out


# --------------------------------------------------------------------------- exptA
class ExptA:
    """`exptA_linop` (src/linops/neklab_linops.f90:35-44; exponential_propagator.f90)."""

    def __init__(self, stepper: PertStepper, tau: float, baseflow: NekVec):
        self.st, self.tau, self.bf = stepper, tau, baseflow
        self.nmatvec = 0

    def init(self):                                   # exponential_propagator.f90:4-13
        self.st.U = [x.copy() for x in self.bf.v]
        self.st.T = self.bf.theta.copy()
        return self.st.setup(self.tau, 0.5, False)

    def _apply(self, vec_in: NekVec, transpose: bool) -> NekVec:
        st = self.st
        nrst = st.prm.torder - 1                      # :23
        st.U = [x.copy() for x in self.bf.v]; st.T = self.bf.theta.copy()      # :25
        st.setup(self.tau, 0.5, transpose)            # :28-32
        st.set_state(vec_in.v, vec_in.pr, vec_in.theta)                        # :35
        st.reset_history()
        for istep in range(1, st.nsteps + 1):         # :39-46
            st.advance(istep)
            if istep <= nrst and vec_in.nrst > 0:     # exptA_get_rst :129-142
                r = vec_in.get_rst(istep)
                st.set_state(r.v, r.pr, r.theta)
        out = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)                     # :49
        out.v = [x.copy() for x in st.vp]; out.pr = st.prp.copy(); out.theta = st.tp.copy()
        for k in range(1, nrst + 1):                  # compute_rst :109-127
            st.advance(st.nsteps + k)
            s = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)
            s.v = [x.copy() for x in st.vp]; s.pr = st.prp.copy(); s.theta = st.tp.copy()
            out.save_rst(s, k)
        self.nmatvec += 1
        return out

    def matvec(self, vec_in):
        return self._apply(vec_in, False)

    def rmatvec(self, vec_in):
        return self._apply(vec_in, True)


# --------------------------------------------------------------------------- exptA_proj
class ExptAProj(ExptA):
    """`exptA_proj_linop` (src/linops/neklab_linops.f90:130-152; exponential_propagator_proj.f90): exptA with the velocity
    projected onto the streamwise wavenumber alpha before the time loop and at tau (`proj_alpha`, :135-173):
    u <- cos(alpha x) <2 u cos(alpha x)> + sin(alpha x) <2 u sin(alpha x)>, <.> = Nek `planar_avg` (bm1-weighted average over
    the points sharing their transverse coordinates)."""

    def __init__(self, stepper, tau, baseflow, alpha, idir=1):
        super().__init__(stepper, tau, baseflow)
        m = stepper.mesh
        ax = idir - 1
        x = m.coords
        ext = max(float(np.ptp(x[:, c])) for c in range(m.ndim))
        keys = np.stack([np.round(x[:, c] / (1e-8 * ext)).astype(np.int64).ravel() for c in range(m.ndim) if c != ax], axis=1)
        _, self.gid = np.unique(keys, axis=0, return_inverse=True)
        self.gid = self.gid.ravel()
        self.cv = np.cos(alpha * x[:, ax]); self.sv = np.sin(alpha * x[:, ax])
        self.msum = np.bincount(self.gid, weights=m.bm1.ravel())

    def planar_avg(self, u):
        m = self.st.mesh
        return (np.bincount(self.gid, weights=(u * m.bm1).ravel()) / self.msum)[self.gid].reshape(u.shape)

    def proj(self, v):
        return [self.cv * self.planar_avg(2.0 * u * self.cv) + self.sv * self.planar_avg(2.0 * u * self.sv) for u in v]

    def _apply(self, vec_in, transpose):
        st = self.st
        nrst = st.prm.torder - 1
        st.U = [x.copy() for x in self.bf.v]; st.T = self.bf.theta.copy()
        st.setup(self.tau, 0.5, transpose)
        st.set_state(self.proj(vec_in.v), vec_in.pr, vec_in.theta)            # :46-47
        st.reset_history()
        for istep in range(1, st.nsteps + 1):
            st.advance(istep)
            if istep <= nrst and vec_in.nrst > 0:
                r = vec_in.get_rst(istep)
                st.set_state(r.v, r.pr, r.theta)
        st.set_state(self.proj(st.vp), st.prp, st.tp)                         # :64-65
        out = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)
        out.v = [x.copy() for x in st.vp]; out.pr = st.prp.copy(); out.theta = st.tp.copy()
        for k in range(1, nrst + 1):
            st.advance(st.nsteps + k)
            s = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)
            s.v = [x.copy() for x in st.vp]; s.pr = st.prp.copy(); s.theta = st.tp.copy()
            out.save_rst(s, k)
        self.nmatvec += 1
        return out


# --------------------------------------------------------------------------- resolvent
class Resolvent:
    """`resolvent_linop` (src/linops/neklab_linops.f90:198-205; src/linops/resolvent.f90): vec_in = (f_re, f_im) -> (re, im).
    tau = 2 pi/|omega| (:22); b = evaluate_rhs (:80-112); re = gmres(I - exptA, b) with kdim 64, atol 1e-12, rtol 1e-6
    (:114-134); im = evaluate_imaginary_part over tau/4 (:136-166)."""

    def __init__(self, stepper: PertStepper, omega: float, baseflow: NekVec, rtol: float = 1.0e-6):
        self.st, self.omega, self.bf, self.rtol = stepper, float(omega), baseflow, rtol

    def integrate(self, tau, f_re: NekVec, f_im: NekVec, x0: Optional[NekVec] = None, adjoint=False) -> NekVec:
        st = self.st
        st.U = [x.copy() for x in self.bf.v]; st.T = self.bf.theta.copy()
        st.setup(tau, 0.5, adjoint)                                   # exptA%init()
        z = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)
        src = x0 if x0 is not None else z                             # opzero(vxp, vyp, vzp)
        st.set_state(src.v, src.pr, src.theta); st.reset_history()
        sign = -1.0 if adjoint else 1.0
        for istep in range(1, st.nsteps + 1):
            th = sign * self.omega * (istep - 1) * st.dt              # alpha = exp(sign i omega time)
            st.forcing = [math.cos(th) * f_re.v[c] - math.sin(th) * f_im.v[c] for c in range(st.d)]       # Re(alpha f), ipert = 1
            st.advance(istep)
        st.forcing = None                                             # zero_neklab_forcing()
        st.adjoint = False
        out = NekVec(st.mesh, st.prm.torder, st.prm.ifheat)
        out.v = [x.copy() for x in st.vp]; out.pr = st.prp.copy(); out.theta = st.tp.copy()
        return out

    def _apply(self, f_re, f_im, adjoint):
        from .krylov import gmres
        tau = 1.0 if self.omega == 0.0 else 2.0 * math.pi / abs(self.omega)
        b = self.integrate(tau, f_re, f_im, None, adjoint)
        A = ExptA(self.st, tau, self.bf)

        def S(x):                                                     # axpby_linop(Id, exptA, 1, -1)
            y = A._apply(x, adjoint); y.axpby(1.0, x, -1.0)
            return y
        x0 = NekVec(self.st.mesh, self.st.prm.torder, self.st.prm.ifheat)
        re = gmres(S, b, x0, kdim=64, atol=1.0e-12, rtol=self.rtol, maxiter=10)
        re.nrst = 0
        im = self.integrate(tau / 4.0, f_re, f_im, re, adjoint)
        return re, im

    def matvec(self, f_re, f_im):
        return self._apply(f_re, f_im, False)

    def rmatvec(self, f_re, f_im):
        return self._apply(f_re, f_im, True)


# --------------------------------------------------------------------------- periodic orbits
class NekExtVec:
    """`nek_ext_dvector` (src/vectors/real_extended_vectors.f90): NekVec + period T (dot adds T*T', :243; axpby :193)."""

    def __init__(self, vec: NekVec, T: float = 0.0):
        self.vec, self.T = vec, float(T)

    def copy(self):
        return NekExtVec(self.vec.copy(), self.T)

    def scal(self, a):
        self.vec.scal(a); self.T *= a

    def axpby(self, alpha, other: "NekExtVec", beta):
        self.vec.axpby(alpha, other.vec, beta); self.T = beta * self.T + alpha * other.T

    def dot(self, other):
        return self.vec.dot(other.vec) + self.T * other.T

    def norm(self):
        return math.sqrt(self.dot(self))


class UPOJacobian:
    """`nek_upo_jacobian` (src/systems/periodic_orbit.f90:46-181) with two steppers: `lin` (perturbation) and `nl` (base flow,
    nonlinear mode), advanced together as Nek does with ifbase: the perturbation step sees the base flow of the previous time
    level (SURVEY.md call stack), then the base flow moves."""

    def __init__(self, lin, nl, X: NekExtVec):
        self.lin, self.nl, self.X = lin, nl, X

    def _coupled(self, istep):
        lin, nl = self.lin, self.nl
        lin.U = [x.copy() for x in nl.vp]; lin._dirty = True       # CPertStepper pushes base flow + mode on the next advance
        lin.advance(istep)
        nl.advance(istep)

    def _fdot(self):                                               # compute_fdot (neklab_systems.f90:202-223)
        nl = self.nl
        v0 = [x.copy() for x in nl.vp]; p0 = nl.prp.copy()
        nl.reset_history()
        nl.advance(1)
        out = NekVec(nl.mesh, nl.prm.torder, nl.prm.ifheat)
        out.v = [(a - b) / nl.dt for a, b in zip(nl.vp, v0)]; out.pr = (nl.prp - p0) / nl.dt
        return NekExtVec(out, 0.0)

    def _apply(self, vin: NekExtVec, transpose: bool) -> NekExtVec:
        lin, nl, X = self.lin, self.nl, self.X
        nrst = lin.prm.torder - 1
        atol = lin.prm.vtol
        fac = 0.5 if transpose else 0.1
        for s in (lin, nl):
            s.prm.vtol = atol * fac; s.prm.ptol = atol * fac
            if hasattr(s, "ref"):
                s.ref.set_params(s.prm, getattr(s, "variant", 0))
        # base flow := X, dt from its CFL at 0.4 over T_X
        nl.nonlinear = True; nl.adjoint = False
        nl.U = [x.copy() for x in X.vec.v]
        nl.setup(X.T, 0.4, False)
        nl.set_state(X.vec.v, X.vec.pr, X.vec.theta); nl.reset_history()
        lin.U = [x.copy() for x in X.vec.v]
        lin.setup(X.T, 0.4, transpose)
        lin.set_state(vin.vec.v, vin.vec.pr, vin.vec.theta); lin.reset_history()
        for istep in range(1, lin.nsteps + 1):
            self._coupled(istep)
            if istep <= nrst and vin.vec.nrst > 0:
                r = vin.vec.get_rst(istep)
                lin.set_state(r.v, r.pr, r.theta)
        out = NekVec(lin.mesh, lin.prm.torder, lin.prm.ifheat)
        out.v = [x.copy() for x in lin.vp]; out.pr = lin.prp.copy(); out.theta = lin.tp.copy()
        for k in range(1, nrst + 1):
            self._coupled(lin.nsteps + k)
            s = NekVec(lin.mesh, lin.prm.torder, lin.prm.ifheat)
            s.v = [x.copy() for x in lin.vp]; s.pr = lin.prp.copy(); s.theta = lin.tp.copy()
            out.save_rst(s, k)
        res = NekExtVec(out, 0.0)
        res.axpby(-1.0, vin, 1.0)                                  # vec_out%sub(vec_in)
        res.axpby(vin.T, self._fdot(), 1.0)                        # + f'(X(T)) dT (base flow where the trajectory stands)
        nl.set_state(X.vec.v, X.vec.pr, X.vec.theta)
        res.T = vin.dot(self._fdot())                              # phase condition at X(0)
        for s in (lin, nl):
            s.prm.vtol = atol; s.prm.ptol = atol
            if hasattr(s, "ref"):
                s.ref.set_params(s.prm, getattr(s, "variant", 0))
        lin.adjoint = False
        return res

    def matvec(self, vin):
        return self._apply(vin, False)

    def rmatvec(self, vin):
        return self._apply(vin, True)
