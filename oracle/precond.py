"""ORACLE (test infrastructure, not product): pressure preconditioner for uzawa_gmres.

Restates the *structure* of Nek5000's `hsmg_solve` for PN-PN-2 (`.par`
`preconditioner = semg_xxt`, `examples/cylinder/stability/direct/1cyl.par:21`;
SURVEY.md App. A.3 item 8, kernels K10/K11): an additive two-level method

    z = W . sum_e R_e^T (S (x) S [(x) S]) Lambda^-1 (S^T (x) S^T [(x) S^T]) R_e r   (overlapping Schwarz, FDM)
      + R_0^T A_0^-1 R_0 r                                                            (vertex coarse grid)

The local problems live on the element's GL pressure points extended by one point
into each face neighbour ((lx2+2)^d = lx1^d points, so the overlap exchange reuses
the velocity-grid gather-scatter, as Nek's `hsmg_extrude`/`hsmg_schwarz_dssum` do).
1-D operators: linear-FE stiffness on the extended point set with the true GL
quadrature mass (`set_up_fast_1D_fem` analogue).  The coarse operator is the exact
Galerkin product R_0 E R_0^T with bilinear vertex interpolation, inverted densely
(`crs_solve` analogue).  hsmg itself is un-vendored; any SPD-like preconditioner gives
the same converged pressure, so only the converged solution is
a parity quantity -- iteration counts are reported, not pinned.
"""
from __future__ import annotations

import numpy as np

from . import ops
from .mesh import SEMesh, face_slices, ax_r, ax_s, ax_t

_AX = (ax_r, ax_s, ax_t)

BC_OVERLAP, BC_DIRICHLET, BC_NEUMANN = 0, 1, 2


def element_lengths(mesh: SEMesh):
    """l[e, k]: mean distance between the two faces normal to local direction k."""
    d = mesh.ndim
    x = mesh.coords                        # (E, d, nz, ny, nx)
    out = np.zeros((mesh.E, d))
    for k in range(d):
        ax = 4 - k                         # array axis of direction k in (E,d,nz,ny,nx)
        lo = np.take(x, 0, axis=ax); hi = np.take(x, -1, axis=ax)
        dist = np.sqrt(((hi - lo) ** 2).sum(axis=1))
        out[:, k] = dist.reshape(mesh.E, -1).mean(axis=1)
    return out


def neighbour_lengths(mesh: SEMesh, lens):
    """(left, right) neighbour lengths per element/direction via the face dssum trick; 0 = no neighbour."""
    d, n = mesh.ndim, mesh.n
    ll = np.zeros_like(lens); lr = np.zeros_like(lens)
    for k in range(d):
        w = np.zeros_like(mesh.bm1)
        f_lo = {0: 3, 1: 0, 2: 4}[k]; f_hi = {0: 1, 1: 2, 2: 5}[k]
        for f in (f_lo, f_hi):
            sl = face_slices(f, d)
            w[(slice(None),) + sl] = lens[:, k].reshape((-1,) + (1,) * (3 - 1))
        s = mesh.dssum(w)
        mid = n // 2
        # sample a face-interior node (never an edge node)
        def sample(f):
            sl = list(face_slices(f, d))
            idx = []
            for a_i, a in enumerate(sl):
                if isinstance(a, slice):
                    dim = (mesh.nz, n, n)[a_i]
                    idx.append(0 if dim == 1 else mid if (n % 2 == 1 or True) else mid)
                else:
                    idx.append(a)
            return s[(slice(None),) + tuple(idx)]
        ll[:, k] = sample(f_lo) - lens[:, k]
        lr[:, k] = sample(f_hi) - lens[:, k]
    ll[np.abs(ll) < 1e-12 * lens] = 0.0
    lr[np.abs(lr) < 1e-12 * lens] = 0.0
    return ll, lr


def face_bc(mesh: SEMesh):
    """bc[e, k, side]: OVERLAP if a neighbour exists, DIRICHLET on outflow, else NEUMANN."""
    d = mesh.ndim
    lens = element_lengths(mesh)
    ll, lr = neighbour_lengths(mesh, lens)
    bc = np.zeros((mesh.E, d, 2), dtype=np.int64)
    for k in range(d):
        f_lo = {0: 3, 1: 0, 2: 4}[k]; f_hi = {0: 1, 1: 2, 2: 5}[k]
        for side, (f, ln) in enumerate(((f_lo, ll), (f_hi, lr))):
            has_nb = ln[:, k] > 0
            outflow = np.isin(mesh.cbc_v[:, f], ("O  ", "o  ", "ON ", "on "))
            bc[:, k, side] = np.where(has_nb, BC_OVERLAP, np.where(outflow, BC_DIRICHLET, BC_NEUMANN))
    return bc, lens, ll, lr


def fdm_1d(z2, w2, lm, ll, lr, bcl, bcr):
    """1-D generalized eigenpairs on the extended grid (n = q+2 slots).

    Returns S (n x n) and lam (n) with S^T B S = I, S^T A S = diag(lam); inactive slots get
    a zero column/row in S (so they neither read nor write).
    """
    q = len(z2)
    n = q + 2
    xm = 0.5 * lm * (z2 + 1.0)
    bm = 0.5 * lm * w2
    pts = []; mass = []; slot = []
    g0 = 0.5 * (1.0 + z2[0])          # distance (in half-lengths) of the GL end point from the face
    if bcl == BC_OVERLAP:
        pts += [-ll * (g0 + 0.5 * (z2[1] - z2[0])), -ll * g0]; mass += [None, 0.5 * ll * w2[0]]; slot += [-1, 0]
    elif bcl == BC_DIRICHLET:
        pts += [0.0]; mass += [None]; slot += [-1]
    for i in range(q):
        pts.append(xm[i]); mass.append(bm[i]); slot.append(i + 1)
    if bcr == BC_OVERLAP:
        pts += [lm + lr * g0, lm + lr * (g0 + 0.5 * (z2[1] - z2[0]))]; mass += [0.5 * lr * w2[0], None]; slot += [n - 1, -1]
    elif bcr == BC_DIRICHLET:
        pts += [lm]; mass += [None]; slot += [-1]
    pts = np.array(pts); N = len(pts)
    A = np.zeros((N, N))
    for i in range(N - 1):
        h = pts[i + 1] - pts[i]
        A[i, i] += 1 / h; A[i + 1, i + 1] += 1 / h; A[i, i + 1] -= 1 / h; A[i + 1, i] -= 1 / h
    act = [i for i in range(N) if slot[i] >= 0]
    Aa = A[np.ix_(act, act)]
    Ba = np.array([mass[i] for i in act])
    # symmetric generalized problem via B^-1/2 scaling
    bi = 1.0 / np.sqrt(Ba)
    lam, V = np.linalg.eigh(Aa * bi[:, None] * bi[None, :])
    Sa = V * bi[:, None]
    S = np.zeros((n, n)); L = np.ones(n)
    sl = [slot[i] for i in act]
    for jj in range(len(act)):
        S[sl, jj] = Sa[:, jj]
        L[jj] = lam[jj]
    nact = len(act)
    return S, L, nact


class SchwarzCoarse:
    def __init__(self, mesh: SEMesh, rho=1.0, use_coarse=True, use_schwarz=True):
        self.mesh = mesh
        d, n, q, E = mesh.ndim, mesh.n, mesh.q, mesh.E
        self.rho = rho
        bc, lens, ll, lr = face_bc(mesh)
        self.S = np.zeros((E, d, n, n)); self.lam = np.ones((E, d, n)); self.nact = np.zeros((E, d), dtype=int)
        b = mesh.b
        for e in range(E):
            for k in range(d):
                self.S[e, k], self.lam[e, k], self.nact[e, k] = fdm_1d(b.z2, b.w2, lens[e, k], ll[e, k], lr[e, k], bc[e, k, 0], bc[e, k, 1])
        # inverse eigenvalue tensor (zero where any direction slot is inactive)
        if d == 2:
            den = self.lam[:, 1, :, None] + self.lam[:, 0, None, :]
            act = (np.arange(n)[None, :] < self.nact[:, 1, None])[:, :, None] & (np.arange(n)[None, :] < self.nact[:, 0, None])[:, None, :]
            self.dinv = np.where(act, 1.0 / np.where(den > 0, den, 1.0), 0.0)[:, None]
            zero_mode = act & (den <= 1e-12 * np.abs(den).max())
            self.dinv[:, 0][zero_mode] = 0.0
        else:
            den = self.lam[:, 2, :, None, None] + self.lam[:, 1, None, :, None] + self.lam[:, 0, None, None, :]
            a = lambda k: (np.arange(n)[None, :] < self.nact[:, k, None])
            act = a(2)[:, :, None, None] & a(1)[:, None, :, None] & a(0)[:, None, None, :]
            self.dinv = np.where(act, 1.0 / np.where(den > 0, den, 1.0), 0.0)
            self.dinv[act & (den <= 1e-12 * np.abs(den).max())] = 0.0
        self.bc = bc
        # overlap-count weights on the pressure grid
        one = np.ones_like(mesh.bm2)
        cnt = self._exchange_back(self._exchange_fwd(self._embed(one)), count=True)
        self.wt = 1.0 / self._interior(cnt)
        self.use_coarse, self.use_schwarz = use_coarse, use_schwarz
        if use_coarse:
            self._setup_coarse()

    # ---- embedding q^d <-> n^d
    def _embed(self, p):
        m = self.mesh
        w = np.zeros_like(m.bm1)
        if m.ndim == 2:
            w[:, :, 1:-1, 1:-1] = p
        else:
            w[:, 1:-1, 1:-1, 1:-1] = p
        return w

    def _interior(self, w):
        return w[:, :, 1:-1, 1:-1] if self.mesh.ndim == 2 else w[:, 1:-1, 1:-1, 1:-1]

    def _face_and_inner(self, k, side):
        """index tuples (into (E,nz,ny,nx)) of the outer layer and first interior layer, tangentially interior."""
        d, n = self.mesh.ndim, self.mesh.n
        inner = slice(1, n - 1)
        idx_f = [slice(None)] + [inner if (d == 3 or a > 0) else slice(None) for a in range(3)]
        idx_i = list(idx_f)
        ax = 3 - k
        idx_f[ax] = 0 if side == 0 else n - 1
        idx_i[ax] = 1 if side == 0 else n - 2
        return tuple(idx_f), tuple(idx_i)

    def _exchange_fwd(self, w):
        """outer layer <- neighbour's first interior layer (hsmg_extrude / dssum / extrude)."""
        m = self.mesh
        w = w.copy()
        for k in range(m.ndim):
            for side in (0, 1):
                f, i = self._face_and_inner(k, side)
                w[f] = w[i]
        s = m.dssum(w)
        for k in range(m.ndim):
            for side in (0, 1):
                f, i = self._face_and_inner(k, side)
                w[f] = s[f] - w[i]
        return w

    def _exchange_back(self, z, count=False):
        """first interior layer += neighbour's outer-layer value."""
        m = self.mesh
        t = np.zeros_like(z)
        for k in range(m.ndim):
            for side in (0, 1):
                f, i = self._face_and_inner(k, side)
                t[f] = z[f]
        s = m.dssum(t)
        out = z.copy()
        for k in range(m.ndim):
            for side in (0, 1):
                f, i = self._face_and_inner(k, side)
                out[i] = out[i] + (s[f] - t[f])
        return out

    def _fdm(self, w):
        d = self.mesh.ndim
        t = w
        for k in range(d):
            St = np.swapaxes(self.S[:, k], 1, 2)
            t = _apply_per_elem(St, t, k)
        t = t * self.dinv
        for k in range(d):
            t = _apply_per_elem(self.S[:, k], t, k)
        return t

    def schwarz(self, r):
        w = self._exchange_fwd(self._embed(r))
        z = self._fdm(w)
        z = self._exchange_back(z)
        return self._interior(z) * self.wt

    # ---- coarse grid
    def _setup_coarse(self):
        m = self.mesh
        d, q = m.ndim, m.q
        z2 = m.b.z2
        H = [0.5 * (1 - z2), 0.5 * (1 + z2)]
        self.shape = []
        for c in range(2 ** d):
            i, j, k = c & 1, (c >> 1) & 1, (c >> 2) & 1
            if d == 2:
                self.shape.append((H[j][:, None] * H[i][None, :])[None])
            else:
                self.shape.append(H[k][:, None, None] * H[j][None, :, None] * H[i][None, None, :])
        self.nv = int(m.vertex.max())
        if self.nv <= 300:                                  # column by column through the operator itself (small cases: cross-check)
            A0 = np.zeros((self.nv, self.nv))
            for v in range(self.nv):
                e0 = np.zeros(self.nv); e0[v] = 1.0
                A0[:, v] = self.restrict(ops.cdabdtp(m, self.prolong(e0), self.rho))
        else:
            A0 = self.galerkin_sparse().toarray()
        A0 = 0.5 * (A0 + A0.T)
        self.A0 = A0
        if m.has_outflow:
            self.A0inv = np.linalg.inv(A0)
        else:
            self.A0inv = np.linalg.pinv(A0, rcond=1e-12, hermitian=True)

    def galerkin_sparse(self):
        """R0 E R0^T assembled algebraically: E = sum_c D_c W_c D_c^T with W_c = mask_c binvm1 / rho on the unique velocity
        nodes, so A0 = sum_c M_c^T W_c M_c with M_c = Q^T D_c^T R0^T (the velocity-space image of every vertex hat function,
        one `cdtp` per corner and component on the whole mesh).  Equal to the column-by-column construction to round-off."""
        import scipy.sparse as sp
        m = self.mesh
        E, nn = m.E, m.bm1[0].size
        rows = np.asarray(m.gidx).reshape(E, nn)
        A0 = None
        for c in range(m.ndim):
            Mc = None
            for k in range(2 ** m.ndim):
                loc = ops.cdtp(m, np.broadcast_to(self.shape[k][None], m.bm2.shape), c).reshape(E, nn)
                cols = np.repeat((m.vertex[:, k] - 1)[:, None], nn, axis=1)
                t = sp.csr_matrix((loc.ravel(), (rows.ravel(), cols.ravel())), shape=(m.nglob, self.nv))
                Mc = t if Mc is None else Mc + t
            wg = np.zeros(m.nglob); wg[m.gidx] = (m.vmask[c] * m.binvm1).ravel() / self.rho
            t = (Mc.T @ sp.diags(wg) @ Mc)
            A0 = t if A0 is None else A0 + t
        return A0.tocsr()

    def prolong(self, c):
        m = self.mesh
        out = np.zeros_like(m.bm2)
        for k in range(2 ** m.ndim):
            out += c[m.vertex[:, k] - 1].reshape((-1, 1, 1, 1)) * self.shape[k][None]
        return out

    def restrict(self, r):
        m = self.mesh
        out = np.zeros(self.nv)
        for k in range(2 ** m.ndim):
            np.add.at(out, m.vertex[:, k] - 1, (r * self.shape[k][None]).reshape(m.E, -1).sum(axis=1))
        return out

    def coarse(self, r):
        return self.prolong(self.A0inv @ self.restrict(r))

    def __call__(self, r):
        z = 0.0
        if self.use_schwarz:
            z = z + self.schwarz(r)
        if self.use_coarse:
            z = z + self.coarse(r)
        return z


def _apply_per_elem(M, u, k):
    """apply per-element matrix M[e] (n x n) along direction k of u (E,nz,ny,nx)."""
    if k == 0:
        return np.einsum("eij,ezyj->ezyi", M, u, optimize=True)
    if k == 1:
        return np.einsum("eij,ezjx->ezix", M, u, optimize=True)
    return np.einsum("eij,ejyx->eiyx", M, u, optimize=True)
