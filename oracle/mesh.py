"""ORACLE (test infrastructure, not product): spectral-element mesh setup.

CPU/numpy restatement of Nek5000's geometry/numbering setup that the exptA
hot path depends on: metrics without 1/J (`geom1`,
`coef.f`), mass matrices `bm1/bm2`, stiffness factors `g1..g6` (`geom2`),
mesh-2 metrics by interpolation (`map12`), dealiasing metrics
(`set_dealias_rx`), direct-stiffness numbering from `.ma2` vertex ids
(`set_vert`/`setvert2d/3d`), Dirichlet masks from `cbc` (`bcmask`),
`binvm1`, `vmult`.  Nek5000 is un-vendored (`Nek5000_setup.sh:56-58`); in-tree
consumers: `src/vectors/real_vectors.f90:100-113` (opdssum, vmult, masks),
`src/linops/neklab_linops.f90:343-362` (metrics in local_grad).

Field layout: arrays of shape (E, nz, ny, nx) -- x (r) fastest, exactly Nek's
`(lx1,ly1,lz1,lelv)` column-major storage read as C order.

PARITY STATUS: pinned only by KAT-1..6 on shipped fixtures (tests/test_oracle_kat.py).
"""
from __future__ import annotations

import numpy as np

from .sem import Basis1D

DIRICHLET_VEL = ("v  ", "V  ", "W  ", "w  ", "vl ", "VL ")   # all components fixed
DIRICHLET_TEMP = ("t  ", "T  ")


def ax_r(M, u):
    return np.einsum("ij,ezyj->ezyi", M, u, optimize=True)


def ax_s(M, u):
    return np.einsum("ij,ezjx->ezix", M, u, optimize=True)


def ax_t(M, u):
    return np.einsum("ij,ejyx->eiyx", M, u, optimize=True)


def tensor_apply(M, u, ndim):
    """(M (x) M [(x) M]) u for every element."""
    out = ax_s(M, ax_r(M, u))
    if ndim == 3:
        out = ax_t(M, out)
    return out


def partition_rank(pid: np.ndarray, nranks: int) -> np.ndarray:
    """Nek's power-of-two partition rule on genmap leaf ids.

    rank(e) = pid(e) // (npstar / P), npstar = 2**ceil(log2(max(pid)+1)).
    """
    if nranks & (nranks - 1):
        raise ValueError("nranks must be a power of two")
    npstar = 1
    while npstar < int(pid.max()) + 1:
        npstar *= 2
    if nranks > npstar:
        raise ValueError("more ranks than partition leaves")
    return (pid // (npstar // nranks)).astype(np.int64)


# local corner index (lexicographic) -> (i,j,k) in {0,1}
def _corner_ijk(c, ndim):
    return (c & 1, (c >> 1) & 1, (c >> 2) & 1 if ndim == 3 else 0)


def glo_num_from_vertices(vertex: np.ndarray, n: int, ndim: int) -> np.ndarray:
    """Global node ids (1-based) from `.ma2` vertex ids, hierarchical like Nek `setvert2d/3d`.

    vertices keep their id; edge-interior nodes are numbered per unique edge
    (oriented from the smaller to the larger endpoint id); face-interior nodes
    (3-D) per unique face (origin at the smallest corner id, first axis towards the
    smaller of its two neighbours); element-interior nodes are private.
    Returns int64 (E, nz, ny, nx).
    """
    E = vertex.shape[0]
    nz = n if ndim == 3 else 1
    g = np.zeros((E, nz, n, n), dtype=np.int64)
    nvert = int(vertex.max())
    ni = n - 2
    # --- vertices
    for c in range(2 ** ndim):
        i, j, k = _corner_ijk(c, ndim)
        g[:, k * (nz - 1), j * (n - 1), i * (n - 1)] = vertex[:, c]
    nxt = nvert + 1
    if ni <= 0:
        return g
    # --- edges: (corner a, corner b, axis) with a<b along axis
    edges = []
    for c in range(2 ** ndim):
        i, j, k = _corner_ijk(c, ndim)
        if i == 0:
            edges.append((c, c | 1, 0))
        if j == 0:
            edges.append((c, c | 2, 1))
        if ndim == 3 and k == 0:
            edges.append((c, c | 4, 2))
    ekeys = []
    for (a, b, ax) in edges:
        va, vb = vertex[:, a], vertex[:, b]
        ekeys.append(np.stack([np.minimum(va, vb), np.maximum(va, vb)], axis=1))
    allk = np.concatenate(ekeys, axis=0)
    uniq, inv = np.unique(allk, axis=0, return_inverse=True)
    inv = inv.reshape(len(edges), E)
    pos = np.arange(1, n - 1)
    for ie, (a, b, ax) in enumerate(edges):
        va, vb = vertex[:, a], vertex[:, b]
        fwd = (va < vb)[:, None]
        p = np.where(fwd, pos[None, :] - 1, (n - 1 - pos[None, :]) - 1)     # 0..ni-1 canonical
        ids = nxt + inv[ie][:, None] * ni + p
        i, j, k = _corner_ijk(a, ndim)
        if ax == 0:
            g[:, k * (nz - 1), j * (n - 1), 1:n - 1] = ids
        elif ax == 1:
            g[:, k * (nz - 1), 1:n - 1, i * (n - 1)] = ids
        else:
            g[:, 1:n - 1, j * (n - 1), i * (n - 1)] = ids
    nxt += len(uniq) * ni
    # --- faces
    if ndim == 3:
        # (fixed axis, side) ; in-face axes (a,b) in lexicographic order
        faces = []
        for ax in range(3):
            for side in (0, 1):
                faces.append((ax, side))
        fkeys = []
        fc = []
        for (ax, side) in faces:
            oth = [d for d in range(3) if d != ax]
            cs = []
            for bb in (0, 1):
                for aa in (0, 1):
                    ijk = [0, 0, 0]
                    ijk[ax] = side; ijk[oth[0]] = aa; ijk[oth[1]] = bb
                    cs.append(ijk[0] | (ijk[1] << 1) | (ijk[2] << 2))
            fc.append(cs)                                   # [c00, c10, c01, c11]
            v4 = np.sort(vertex[:, cs], axis=1)
            fkeys.append(v4)
        allf = np.concatenate(fkeys, axis=0)
        uniq, inv = np.unique(allf, axis=0, return_inverse=True)
        inv = inv.reshape(len(faces), E)
        A, B = np.meshgrid(np.arange(1, n - 1), np.arange(1, n - 1), indexing="xy")   # A: first in-face axis (fast)
        for iff, (ax, side) in enumerate(faces):
            cs = fc[iff]
            v = vertex[:, cs]                               # (E,4): c00 c10 c01 c11
            cmin = np.argmin(v, axis=1)                     # which corner holds smallest id
            amin = cmin & 1; bmin = (cmin >> 1) & 1
            # neighbours of the min corner along a and along b
            na = v[np.arange(E), (1 - amin) | (bmin << 1)]
            nb = v[np.arange(E), amin | ((1 - bmin) << 1)]
            swap = nb < na
            ia = np.where(amin[:, None, None] == 0, A[None] - 1, (n - 1 - A[None]) - 1)
            ib = np.where(bmin[:, None, None] == 0, B[None] - 1, (n - 1 - B[None]) - 1)
            p = np.where(swap[:, None, None], ib + ni * ia, ia + ni * ib)
            ids = nxt + inv[iff][:, None, None] * ni * ni + p      # (E, b, a)
            s = side * (n - 1)
            if ax == 0:      # in-face axes (y,z): a=y, b=z -> g[:, z, y, s]
                g[:, 1:n - 1, 1:n - 1, s] = ids
            elif ax == 1:    # in-face axes (x,z): a=x, b=z -> g[:, z, s, x]
                g[:, 1:n - 1, s, 1:n - 1] = ids
            else:            # in-face axes (x,y): a=x, b=y -> g[:, s, y, x]
                g[:, s, 1:n - 1, 1:n - 1] = ids
        nxt += len(uniq) * ni * ni
    # --- interiors
    nint = ni ** ndim
    loc = np.arange(nint).reshape((ni,) * ndim)
    if ndim == 3:
        g[:, 1:n - 1, 1:n - 1, 1:n - 1] = nxt + (np.arange(E) * nint)[:, None, None, None] + loc[None]
    else:
        g[:, 0, 1:n - 1, 1:n - 1] = nxt + (np.arange(E) * nint)[:, None, None] + loc[None]
    return g


def glo_num_from_coords(coords: np.ndarray, periods=None, tol=1e-6) -> np.ndarray:
    """Independent numbering by coordinate coincidence (test cross-check)."""
    E, d = coords.shape[:2]
    pts = coords.transpose(0, 2, 3, 4, 1).reshape(-1, d).copy()
    if periods is not None:
        for ax, (lo, L) in periods.items():
            x = pts[:, ax] - lo
            x = np.where(np.abs(x - L) < 10 * tol, 0.0, x)
            pts[:, ax] = np.mod(x, L)
            pts[:, ax] = np.where(np.abs(pts[:, ax] - L) < 10 * tol, 0.0, pts[:, ax])
    key = np.round(pts / tol).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    return (inv.reshape(coords.shape[0], *coords.shape[2:]) + 1).astype(np.int64)


def same_partition(g1: np.ndarray, g2: np.ndarray) -> bool:
    """True iff two numberings induce the same coincidence classes."""
    a = g1.ravel(); b = g2.ravel()
    _, ia = np.unique(a, return_inverse=True)
    _, ib = np.unique(b, return_inverse=True)
    # bijection check
    pa = np.unique(np.stack([ia, ib], 1), axis=0)
    return len(pa) == ia.max() + 1 == ib.max() + 1


def face_slices(f: int, ndim: int):
    """Preprocessor face number (0-based) -> index tuple into (nz,ny,nx)."""
    s = slice(None)
    if f == 0:
        return (s, 0, s)
    if f == 1:
        return (s, s, -1)
    if f == 2:
        return (s, -1, s)
    if f == 3:
        return (s, s, 0)
    if f == 4:
        return (0, s, s)
    return (-1, s, s)


class SEMesh:
    """Everything geometric the PN-PN-2 perturbation stepper needs, on the host in numpy."""

    def __init__(self, coords: np.ndarray, vertex: np.ndarray, cbc_v: np.ndarray, lxd: int,
                 cbc_t: np.ndarray | None = None):
        """coords (E, ndim, nz, ny, nx); vertex (E, 2**ndim) ma2 ids; cbc (E, 2*ndim) 'U3'."""
        self.E, self.ndim = coords.shape[0], coords.shape[1]
        d = self.ndim
        n = coords.shape[-1]
        self.n, self.m, self.q = n, lxd, n - 2
        self.b = b = Basis1D(n, lxd)
        self.nz = n if d == 3 else 1
        self.coords = coords
        x = [coords[:, c] for c in range(d)]
        D = b.D
        w = b.w1
        if d == 2:
            xr, xs = ax_r(D, x[0]), ax_s(D, x[0])
            yr, ys = ax_r(D, x[1]), ax_s(D, x[1])
            self.jac = xr * ys - xs * yr
            # rx[k][c] = d(r_k)/d(x_c) * J
            self.rx = [[ys, -xs], [-yr, xr]]
            W = (w[:, None] * w[None, :])[None, None]
        else:
            xr, xs, xt = ax_r(D, x[0]), ax_s(D, x[0]), ax_t(D, x[0])
            yr, ys, yt = ax_r(D, x[1]), ax_s(D, x[1]), ax_t(D, x[1])
            zr, zs, zt = ax_r(D, x[2]), ax_s(D, x[2]), ax_t(D, x[2])
            self.jac = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr)
            self.rx = [[ys * zt - yt * zs, xt * zs - xs * zt, xs * yt - xt * ys],
                       [yt * zr - yr * zt, xr * zt - xt * zr, xt * yr - xr * yt],
                       [yr * zs - ys * zr, xs * zr - xr * zs, xr * ys - xs * yr]]
            W = (w[:, None, None] * w[None, :, None] * w[None, None, :])[None]
        if np.any(self.jac <= 0):
            raise ValueError("non-positive Jacobian")
        self.W1 = W
        self.bm1 = self.jac * W
        # stiffness factors G[k][l] = sum_c rx[k][c] rx[l][c] * W / J
        self.G = [[sum(self.rx[k][c] * self.rx[l][c] for c in range(d)) * W / self.jac
                   for l in range(d)] for k in range(d)]
        # mesh 2 (GL, q points): interpolated metrics (map12), bm2
        I12 = b.I12
        self.rx2 = [[tensor_apply(I12, self.rx[k][c], d) for c in range(d)] for k in range(d)]
        self.jac2 = tensor_apply(I12, self.jac, d)
        w2 = b.w2
        if d == 2:
            self.W2 = (w2[:, None] * w2[None, :])[None, None]
        else:
            self.W2 = (w2[:, None, None] * w2[None, :, None] * w2[None, None, :])[None]
        self.bm2 = self.jac2 * self.W2
        # dealias mesh: rx_d = I1d(rx) * wd (x) wd  (set_dealias_rx)
        wd = b.wd
        if d == 2:
            Wd = (wd[:, None] * wd[None, :])[None, None]
        else:
            Wd = (wd[:, None, None] * wd[None, :, None] * wd[None, None, :])[None]
        self.rxd = [[tensor_apply(b.I1d, self.rx[k][c], d) * Wd for c in range(d)] for k in range(d)]
        # numbering
        self.vertex = vertex
        self.glo = glo_num_from_vertices(vertex, n, d)
        flat = self.glo.ravel()
        self.uniq, self.gidx = np.unique(flat, return_inverse=True)
        self.nglob = len(self.uniq)
        self.mult_count = np.bincount(self.gidx, minlength=self.nglob)
        self.vmult = 1.0 / self.dssum(np.ones_like(self.bm1))
        self.binvm1 = 1.0 / self.dssum(self.bm1)
        self.volvm1 = float(self.bm1.sum())
        self.volvm2 = float(self.bm2.sum())
        # masks
        self.cbc_v = cbc_v
        self.vmask = [self._mask(cbc_v, c) for c in range(d)]
        self.cbc_t = cbc_t
        self.tmask = self._mask_t(cbc_t) if cbc_t is not None else None
        # does the pressure operator have a null space? (no outflow => yes)
        self.has_outflow = bool(np.isin(cbc_v, ("O  ", "o  ", "ON ", "on ")).any())

    # ------------------------------------------------------------------
    def dssum(self, u: np.ndarray) -> np.ndarray:
        """Direct stiffness summation (Nek `dssum` -> gs_op add)."""
        s = np.bincount(self.gidx, weights=u.ravel(), minlength=self.nglob)
        return s[self.gidx].reshape(u.shape)

    def dsop_min(self, u):
        s = np.full(self.nglob, np.inf)
        np.minimum.at(s, self.gidx, u.ravel())
        return s[self.gidx].reshape(u.shape)

    def _mask(self, cbc, comp):
        m = np.ones_like(self.bm1)
        d = self.ndim
        for f in range(2 * d):
            sl = face_slices(f, d)
            dirich = np.isin(cbc[:, f], DIRICHLET_VEL)
            # SYM (Nek `bcmask`, axis-aligned symmetry planes only): zero the component along the face's PHYSICAL normal,
            # found from the coordinates (unstructured meshes rotate their elements); 'SYx'/'SYy'/'SYz' name it explicitly
            codes = np.char.upper(cbc[:, f].astype(str))
            sym = np.char.startswith(codes, "SY")
            fix = dirich.copy()
            if sym.any():
                ext = np.stack([np.ptp(self.coords[(slice(None), c) + sl].reshape(len(cbc), -1), axis=1) for c in range(d)], axis=1)
                phys = np.argmin(ext, axis=1)
                explicit = np.array([{"X": 0, "Y": 1, "Z": 2}.get(s[2:3], -1) for s in codes])
                bad = sym & (explicit < 0) & (ext[np.arange(len(cbc)), phys] > 1e-8 * ext.max(axis=1))
                if bad.any():
                    raise ValueError("SYM face is not a coordinate plane")
                axis = np.where(explicit >= 0, explicit, phys)
                fix |= sym & (axis == comp)
            idx = np.where(fix)[0]
            if len(idx):
                m[(idx,) + sl] = 0.0
        # a Dirichlet node is Dirichlet in every element sharing it
        return self.dsop_min(m)

    def _mask_t(self, cbc):
        m = np.ones_like(self.bm1)
        for f in range(2 * self.ndim):
            idx = np.where(np.isin(cbc[:, f], DIRICHLET_TEMP))[0]
            if len(idx):
                m[(idx,) + face_slices(f, self.ndim)] = 0.0
        return self.dsop_min(m)


def lex_corners_from_re2(xyz: np.ndarray, ndim: int) -> np.ndarray:
    """re2 corner order (preprocessor, ccw) -> lexicographic order. (E, ndim, 2**ndim)."""
    perm = [0, 1, 3, 2] if ndim == 2 else [0, 1, 3, 2, 4, 5, 7, 6]
    return xyz[:, :, perm]


def coords_from_corners(xyz_lex: np.ndarray, n: int) -> np.ndarray:
    """Straight-sided (bi/tri-linear) GLL coordinates from lexicographic corners."""
    from .sem import gll
    z, _ = gll(n)
    E, d, nv = xyz_lex.shape
    h0 = 0.5 * (1 - z); h1 = 0.5 * (1 + z)
    H = [h0, h1]
    nz = n if d == 3 else 1
    out = np.zeros((E, d, nz, n, n))
    for c in range(nv):
        i, j, k = _corner_ijk(c, d)
        if d == 2:
            shp = (H[j][:, None] * H[i][None, :])[None]
        else:
            shp = H[k][:, None, None] * H[j][None, :, None] * H[i][None, None, :]
        out += xyz_lex[:, :, c][:, :, None, None, None] * shp[None, None]
    return out
