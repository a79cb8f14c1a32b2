"""Python host mirror of neklab's operator/vector interface over the libnlk C-ABI (include/nlk.h).

Names follow the reference: `nek_dvector` (src/vectors/neklab_vectors.f90:26-50), `exptA_linop`
(src/linops/neklab_linops.f90:35-44), `linear_stability_analysis_fixed_point`
(src/neklab_analysis.f90:38-105).  Everything numerical happens inside libnlk.so on the GPU; this
module only marshals arguments.  There is no CPU fallback: if the library or a device is missing the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libnlk.so")
_lib = None


class NlkError(RuntimeError):
    pass


class MeshDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("lx1", C.c_int32), ("lxd", C.c_int32), ("nelg", C.c_int64), ("nel", C.c_int64),
                ("xm1", C.c_void_p), ("ym1", C.c_void_p), ("zm1", C.c_void_p), ("vertex", C.c_void_p),
                ("cbc_v", C.c_void_p), ("cbc_t", C.c_void_p), ("gllnid", C.c_void_p), ("rank", C.c_int32), ("nranks", C.c_int32)]


class MeshInfo(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("lx1", C.c_int32), ("lx2", C.c_int32), ("lxd", C.c_int32), ("nel", C.c_int64),
                ("nelg", C.c_int64), ("np1", C.c_int64), ("np2", C.c_int64), ("nglob_local", C.c_int64),
                ("nshared_local", C.c_int64), ("nvert", C.c_int64), ("has_outflow", C.c_int32), ("nneigh", C.c_int32),
                ("volvm1", C.c_double), ("volvm2", C.c_double)]


class Params(C.Structure):
    _fields_ = [("viscosity", C.c_double), ("density", C.c_double), ("torder", C.c_int32), ("vtol", C.c_double),
                ("ptol", C.c_double), ("ifheat", C.c_int32), ("conductivity", C.c_double), ("rhocp", C.c_double),
                ("ttol", C.c_double), ("buoyancy", C.c_double * 3), ("filter_weight", C.c_double),
                ("filter_cutoff", C.c_double), ("cg_maxit", C.c_int32), ("gmres_maxit", C.c_int32), ("lgmres", C.c_int32),
                ("precond", C.c_int32), ("pr_proj", C.c_int32), ("cfl_limit", C.c_double), ("rst_mode", C.c_int32), ("coarse_iters", C.c_int32), ("step_variant", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("nsteps", C.c_int32), ("dt", C.c_double), ("cg_iters", C.c_int64), ("gmres_iters", C.c_int64),
                ("steps", C.c_int64), ("matvecs", C.c_int64), ("ms_total", C.c_double), ("launches", C.c_int64)]


EIGS_CB = C.CFUNCTYPE(None, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

# every symbol include/nlk.h declares (checked by tests/test_cabi.py)
SYMBOLS = """nlk_last_error nlk_version nlk_partition nlk_mesh_create nlk_mesh_destroy nlk_mesh_info nlk_mesh_glo_num
nlk_mesh_field nlk_mesh_neighbor nlk_mesh_basis nlk_dense_eig nlk_params_default nlk_ctx_create nlk_ctx_destroy nlk_ctx_set_tol
nlk_ctx_set_dt nlk_comm_unique_id nlk_ctx_comm_init nlk_ctx_sync nlk_ctx_stream nlk_vec_create nlk_vec_destroy nlk_vec_copy nlk_vec_zero
nlk_vec_rand nlk_vec_scal nlk_vec_axpby nlk_vec_dot nlk_vec_norm nlk_vec_size nlk_vec_save_rst nlk_vec_get_rst nlk_vec_nrst
nlk_vec_clear_rst nlk_vec_upload nlk_vec_download nlk_zvec_scal nlk_zvec_axpby nlk_zvec_dot nlk_basis_innerprod nlk_basis_axpy nlk_basis_dgs nlk_exptA_create
nlk_exptA_destroy nlk_exptA_init nlk_exptA_set_tau nlk_exptA_matvec nlk_exptA_rmatvec nlk_exptA_stats nlk_exptA_time_steps nlk_exptA_set_baseflow nlk_exptA_set_projection nlk_exptA_apply_projection nlk_nonlinear_map nlk_newton_fixed_point nlk_ctx_set_forcing nlk_set_neklab_forcing nlk_get_neklab_forcing nlk_zero_neklab_forcing nlk_zero_neklab_forcing_ipert nlk_nek2vec nlk_vec2nek
nlk_eigs nlk_svds nlk_gmres nlk_resolvent_matvec nlk_resolvent_integrate nlk_upo_jacobian nlk_test_axhelm nlk_test_dssum nlk_test_opdiv nlk_test_opgradt nlk_test_cdabdtp nlk_test_convect
nlk_test_convect_adj nlk_test_helmholtz nlk_test_pressure nlk_test_precond nlk_test_cfl nlk_bench_kernel""".split()


def lib():
    """Load libnlk.so (fails loudly if it has not been built: `python -m neklab_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise NlkError(f"{_LIBPATH} not built -- run `python -m neklab_b200.build` (no CPU fallback exists)")
        _lib = C.CDLL(_LIBPATH)
        _lib.nlk_last_error.restype = C.c_char_p
        _lib.nlk_ctx_stream.restype = C.c_void_p
    return _lib


def _chk(rc):
    if rc != 0:
        raise NlkError(lib().nlk_last_error().decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def cbc_bytes(cbc):
    """(E, nface) array of 3-char codes -> contiguous uint8 buffer [E][nface][3]."""
    a = np.char.encode(np.asarray(cbc, dtype="U3"), "ascii").astype("S3")
    return np.ascontiguousarray(a).view(np.uint8).copy()


def partition(pid, nranks):
    pid = np.ascontiguousarray(pid, dtype=np.int64)
    out = np.zeros(len(pid), dtype=np.int32)
    _chk(lib().nlk_partition(_p(pid), C.c_int64(len(pid)), C.c_int32(nranks), _p(out)))
    return out


def dense_eig(A):
    A = _f64(A); n = A.shape[0]
    wr = np.zeros(n); wi = np.zeros(n); VR = np.zeros((n, n))
    _chk(lib().nlk_dense_eig(C.c_int32(n), _p(A), _p(wr), _p(wi), _p(VR)))
    return wr, wi, VR


def resolve_sym(coords, cbc):
    """Rewrite 'SYM' boundary codes as 'SYx' / 'SYy' / 'SYz' (PHYSICAL normal of the symmetry plane, read off the
    coordinates) -- unstructured meshes rotate their elements, so the plane's normal is not the element's reference axis.
    coords: (E, ndim, nz, ny, nx) for the SAME elements as cbc (E, nface): call it on the global arrays before partitioning.
    Only coordinate planes are supported (all the reference's configs: back_fstep `bfs.usr` usrdat, `setbc(4,1,'SYM')`)."""
    cbc = np.array(cbc, dtype="U3")
    coords = np.asarray(coords)
    d = coords.shape[1]
    s = slice(None)
    faces = [(s, 0, s), (s, s, -1), (s, -1, s), (s, s, 0), (0, s, s), (-1, s, s)][:2 * d]
    for f, sl in enumerate(faces):
        idx = np.where(cbc[:, f] == "SYM")[0]
        if not len(idx):
            continue
        ext = np.stack([np.ptp(coords[(idx, c) + sl].reshape(len(idx), -1), axis=1) for c in range(d)], axis=1)
        ax = np.argmin(ext, axis=1)
        if (ext[np.arange(len(idx)), ax] > 1e-8 * ext.max(axis=1)).any():
            raise NlkError("a 'SYM' face is not a coordinate plane (general symmetry planes are not supported)")
        cbc[idx, f] = np.array(["SYx", "SYy", "SYz"])[ax]
    return cbc


class Mesh:
    """Host-side mesh (geometry + numbering).  coords: (E_local, ndim, nz, ny, nx); vertex: (E_global, 2**ndim).
    'SYM' codes are resolved to physical axes here when the coordinates cover every element (single rank); multi-rank
    callers run `resolve_sym` on the global arrays first."""

    def __init__(self, coords, vertex, cbc_v, lxd, cbc_t=None, gllnid=None, rank=0, nranks=1):
        L = lib()
        coords = _f64(coords)
        if coords.shape[0] == np.asarray(vertex).shape[0]:
            cbc_v = resolve_sym(coords, cbc_v)
        self.ndim = coords.shape[1]; self.lx1 = coords.shape[-1]; self.lxd = lxd
        self._x = [np.ascontiguousarray(coords[:, c]) for c in range(self.ndim)]
        self._vertex = np.ascontiguousarray(vertex, dtype=np.int64)
        self._cbcv = cbc_bytes(cbc_v)
        self._cbct = cbc_bytes(cbc_t) if cbc_t is not None else None
        self._gllnid = None if gllnid is None else np.ascontiguousarray(gllnid, dtype=np.int32)
        self._rank, self._nranks = int(rank), int(nranks)
        d = MeshDesc(self.ndim, self.lx1, lxd, self._vertex.shape[0], coords.shape[0], _p(self._x[0]), _p(self._x[1]),
                     _p(self._x[2]) if self.ndim == 3 else None, _p(self._vertex), _p(self._cbcv), _p(self._cbct),
                     _p(self._gllnid), rank, nranks)
        self.h = C.c_void_p()
        _chk(L.nlk_mesh_create(C.byref(d), C.byref(self.h)))
        self.info = MeshInfo()
        _chk(L.nlk_mesh_info(self.h, C.byref(self.info)))
        self.nel = int(self.info.nel)
        n = self.lx1; q = n - 2
        self.shape1 = (self.nel, n if self.ndim == 3 else 1, n, n)
        self.shape2 = (self.nel, q if self.ndim == 3 else 1, q, q)

    def glo_num(self):
        g = np.zeros(self.shape1, dtype=np.int64)
        _chk(lib().nlk_mesh_glo_num(self.h, _p(g)))
        return g

    def field(self, name):
        a = np.zeros(self.shape2 if name == "bm2" else self.shape1)
        _chk(lib().nlk_mesh_field(self.h, name.encode(), _p(a)))
        return a

    def basis(self, name):
        buf = np.zeros(32 * 32)
        n = lib().nlk_mesh_basis(self.h, name.encode(), _p(buf), C.c_int64(buf.size))
        if n < 0:
            _chk(1)
        return buf[:n].copy()

    def neighbors(self):
        out = []
        for k in range(self.info.nneigh):
            r = C.c_int32(); cnt = C.c_int64()
            _chk(lib().nlk_mesh_neighbor(self.h, k, C.byref(r), C.byref(cnt), None))
            g = np.zeros(cnt.value, dtype=np.int64)
            _chk(lib().nlk_mesh_neighbor(self.h, k, C.byref(r), C.byref(cnt), _p(g)))
            out.append((r.value, g))
        return out

    def __del__(self):
        try:
            if self.h:
                lib().nlk_mesh_destroy(self.h); self.h = None
        except Exception:
            pass


def default_params(**kw):
    p = Params()
    _chk(lib().nlk_params_default(C.byref(p)))
    for k, v in kw.items():
        if k == "buoyancy":
            for i, x in enumerate(v):
                p.buoyancy[i] = x
        else:
            setattr(p, k, v)
    return p


class Context:
    """Device-resident solver state (the reference's Nek COMMON blocks)."""

    def __init__(self, mesh: Mesh, params: Params, device=0, nccl_id: bytes | None = None):
        self.mesh = mesh; self.params = params
        self.h = C.c_void_p()
        _chk(lib().nlk_ctx_create(mesh.h, C.byref(params), C.c_int32(device), C.byref(self.h)))
        if mesh._nranks > 1:
            if nccl_id is None:
                raise NlkError("multi-rank mesh needs an NCCL unique id (nlk_comm_unique_id on rank 0, broadcast)")
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            _chk(lib().nlk_ctx_comm_init(self.h, buf, C.c_int32(mesh._rank), C.c_int32(mesh._nranks)))

    def vec(self):
        return nek_dvector(self)

    def set_tol(self, vtol, ptol):
        _chk(lib().nlk_ctx_set_tol(self.h, C.c_double(vtol), C.c_double(ptol)))
        self.params.vtol = vtol; self.params.ptol = ptol; self.params.ttol = vtol

    def set_dt(self, dt):
        _chk(lib().nlk_ctx_set_dt(self.h, C.c_double(dt)))

    def sync(self):
        _chk(lib().nlk_ctx_sync(self.h))

    # ---- kernel-level hooks (tests / bench)
    def axhelm(self, u, h1, h2):
        u = _f64(u); w = np.zeros_like(u)
        _chk(lib().nlk_test_axhelm(self.h, _p(u), C.c_double(h1), C.c_double(h2), _p(w))); return w

    def dssum(self, u):
        u = _f64(u).copy()
        _chk(lib().nlk_test_dssum(self.h, _p(u))); return u

    def opdiv(self, u):
        u = [_f64(x) for x in u]
        p = np.zeros(self.mesh.shape2)
        _chk(lib().nlk_test_opdiv(self.h, _p(u[0]), _p(u[1]), _p(u[2]) if len(u) > 2 else None, _p(p))); return p

    def opgradt(self, p):
        p = _f64(p); d = self.mesh.ndim
        w = [np.zeros(self.mesh.shape1) for _ in range(d)]
        _chk(lib().nlk_test_opgradt(self.h, _p(p), _p(w[0]), _p(w[1]), _p(w[2]) if d == 3 else None)); return w

    def cdabdtp(self, p):
        p = _f64(p); ep = np.zeros_like(p)
        _chk(lib().nlk_test_cdabdtp(self.h, _p(p), _p(ep))); return ep

    def convect(self, u, Cv):
        u = _f64(u); Cv = [_f64(x) for x in Cv]; out = np.zeros_like(u)
        _chk(lib().nlk_test_convect(self.h, _p(u), _p(Cv[0]), _p(Cv[1]), _p(Cv[2]) if len(Cv) > 2 else None, _p(out))); return out

    def convect_adj(self, U, cf):
        d = self.mesh.ndim
        U = [_f64(x) for x in U]; cf = [_f64(x) for x in cf]; out = [np.zeros(self.mesh.shape1) for _ in range(d)]
        arr = lambda lst: (C.c_void_p * 3)(*[_p(x).value if x is not None else None for x in (lst + [None] * 3)[:3]])
        _chk(lib().nlk_test_convect_adj(self.h, arr(U), arr(cf), arr(out))); return out

    def helmholtz(self, f, h1, h2, comp, tol):
        f = _f64(f); x = np.zeros_like(f); it = C.c_int32()
        _chk(lib().nlk_test_helmholtz(self.h, _p(f), C.c_double(h1), C.c_double(h2), C.c_int32(comp), C.c_double(tol), _p(x), C.byref(it)))
        return x, it.value

    def pressure(self, rhs, tol):
        rhs = _f64(rhs); x = np.zeros_like(rhs); it = C.c_int32()
        _chk(lib().nlk_test_pressure(self.h, _p(rhs), C.c_double(tol), _p(x), C.byref(it))); return x, it.value

    def precond(self, r):
        r = _f64(r); z = np.zeros_like(r)
        _chk(lib().nlk_test_precond(self.h, _p(r), _p(z))); return z

    def cfl(self, u, dt):
        u = [_f64(x) for x in u]; out = C.c_double()
        _chk(lib().nlk_test_cfl(self.h, _p(u[0]), _p(u[1]), _p(u[2]) if len(u) > 2 else None, C.c_double(dt), C.byref(out))); return out.value

    def bench_kernel(self, which, nrep):
        ms = C.c_double(); by = C.c_double()
        _chk(lib().nlk_bench_kernel(self.h, C.c_int32(which), C.c_int32(nrep), C.byref(ms), C.byref(by))); return ms.value, by.value

    def nek2vec(self, vec=None):
        """`nek2vec` (src/neklab_utils.f90:84-134) on the device-resident state."""
        vec = vec or nek_dvector(self)
        _chk(lib().nlk_nek2vec(self.h, vec.h)); return vec

    def vec2nek(self, vec):
        _chk(lib().nlk_vec2nek(self.h, vec.h))

    def set_forcing(self, f, ipert=1):
        """`set_neklab_forcing(fx, fy, fz, ipert)` (src/neklab_nek_forcing.f90:57-76)."""
        f = [_f64(x) for x in f]
        _chk(lib().nlk_set_neklab_forcing(self.h, _p(f[0]), _p(f[1]), _p(f[2]) if len(f) > 2 else None, C.c_int32(ipert)))

    def get_forcing(self, ipert=1):
        f = [np.zeros(self.mesh.shape1) for _ in range(self.mesh.ndim)]
        _chk(lib().nlk_get_neklab_forcing(self.h, _p(f[0]), _p(f[1]), _p(f[2]) if len(f) > 2 else None, C.c_int32(ipert))); return f

    def zero_forcing(self, ipert=None):
        _chk(lib().nlk_zero_neklab_forcing(self.h) if ipert is None else lib().nlk_zero_neklab_forcing_ipert(self.h, C.c_int32(ipert)))

    def close(self):
        if self.h:
            lib().nlk_ctx_destroy(self.h); self.h = None


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(lib().nlk_comm_unique_id(buf))
    return buf.raw


class nek_dvector:
    """Device-resident `nek_dvector` (src/vectors/neklab_vectors.f90:26-50; TBPs in real_vectors.f90)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.h = C.c_void_p()
        _chk(lib().nlk_vec_create(ctx.h, C.byref(self.h)))

    def copy(self):
        o = nek_dvector(self.ctx); _chk(lib().nlk_vec_copy(o.h, self.h)); return o

    def zero(self):
        _chk(lib().nlk_vec_zero(self.h))

    def rand(self, ifnorm=False, seed=12345):
        _chk(lib().nlk_vec_rand(self.h, C.c_int32(int(ifnorm)), C.c_uint64(seed)))

    def scal(self, alpha):
        _chk(lib().nlk_vec_scal(self.h, C.c_double(alpha)))

    def axpby(self, alpha, vec, beta):
        _chk(lib().nlk_vec_axpby(C.c_double(alpha), vec.h, C.c_double(beta), self.h))

    def dot(self, vec):
        out = C.c_double(); _chk(lib().nlk_vec_dot(self.h, vec.h, C.byref(out))); return out.value

    def norm(self):
        out = C.c_double(); _chk(lib().nlk_vec_norm(self.h, C.byref(out))); return out.value

    def get_size(self):
        n = C.c_int64(); _chk(lib().nlk_vec_size(self.h, C.byref(n))); return n.value

    def save_rst(self, state, irst):
        _chk(lib().nlk_vec_save_rst(self.h, state.h, C.c_int32(irst)))

    def get_rst(self, irst):
        o = nek_dvector(self.ctx); _chk(lib().nlk_vec_get_rst(self.h, o.h, C.c_int32(irst))); return o

    @property
    def nrst(self):
        n = C.c_int32(); _chk(lib().nlk_vec_nrst(self.h, C.byref(n))); return n.value

    def has_rst_fields(self):
        return self.nrst > 0

    def clear_rst_fields(self):
        _chk(lib().nlk_vec_clear_rst(self.h))

    def upload(self, v=None, pr=None, theta=None):
        v = [_f64(x) for x in v] if v is not None else [None, None, None]
        v = (list(v) + [None] * 3)[:3]
        pr = _f64(pr); theta = _f64(theta)
        _chk(lib().nlk_vec_upload(self.h, _p(v[0]), _p(v[1]), _p(v[2]), _p(pr), _p(theta)))

    def download(self):
        m = self.ctx.mesh
        v = [np.zeros(m.shape1) for _ in range(m.ndim)]
        pr = np.zeros(m.shape2); th = np.zeros(m.shape1)
        _chk(lib().nlk_vec_download(self.h, _p(v[0]), _p(v[1]), _p(v[2]) if m.ndim == 3 else None, _p(pr), _p(th)))
        return v, pr, th

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib().nlk_vec_destroy(self.h)
            self.h = None
        except Exception:
            pass


class nek_ext_dvector:
    """`nek_ext_dvector` (src/vectors/neklab_vectors.f90:121-147; real_extended_vectors.f90): a `nek_dvector` plus the scalar
    period `T` of a periodic orbit (and its restart copies `Trst`).  The fields live on the device; `T` is a host scalar exactly as
    in the reference, where it is one double next to the field arrays.  dot adds T*T' (:243), axpby combines T and -- unlike the
    fields -- Trst consistently (:193, :210), get_size counts it."""

    def __init__(self, ctx: Context, T: float = 0.0):
        self.ctx = ctx; self.vec = nek_dvector(ctx); self.T = float(T); self.Trst = [0.0, 0.0]

    def zero(self):
        self.vec.zero(); self.T = 0.0; self.Trst = [0.0, 0.0]

    def rand(self, ifnorm=False, seed=12345):
        self.vec.rand(False, seed)
        self.T = float(np.random.default_rng(seed).random())                     # random_number(self%T) (:117)
        if ifnorm:
            self.scal(1.0 / self.norm())

    def scal(self, alpha):
        self.vec.scal(alpha); self.T *= alpha; self.Trst = [alpha * t for t in self.Trst]

    def axpby(self, alpha, vec: "nek_ext_dvector", beta):
        self.vec.axpby(alpha, vec.vec, beta)
        self.T = beta * self.T + alpha * vec.T
        self.Trst = [beta * a + alpha * b for a, b in zip(self.Trst, vec.Trst)]

    def dot(self, vec: "nek_ext_dvector"):
        return self.vec.dot(vec.vec) + self.T * vec.T

    def norm(self):
        return float(np.sqrt(self.dot(self)))

    def get_size(self):
        return self.vec.get_size() + 1

    def save_rst(self, state: "nek_ext_dvector", irst):
        self.vec.save_rst(state.vec, irst); self.Trst[irst - 1] = state.T

    def get_rst(self, irst):
        o = nek_ext_dvector(self.ctx); o.vec = self.vec.get_rst(irst); o.T = self.Trst[irst - 1]; return o


class nek_zvector:
    """`nek_zvector` (src/vectors/neklab_vectors.f90:219-237): complex vector as a (re, im) pair of `nek_dvector`."""

    def __init__(self, ctx: "Context", re: nek_dvector | None = None, im: nek_dvector | None = None):
        self.ctx = ctx
        self.re = re if re is not None else nek_dvector(ctx)
        self.im = im if im is not None else nek_dvector(ctx)

    def zero(self):
        self.re.zero(); self.im.zero()

    def scal(self, alpha: complex):
        _chk(lib().nlk_zvec_scal(self.re.h, self.im.h, C.c_double(alpha.real), C.c_double(alpha.imag)))

    def axpby(self, alpha: complex, vec: "nek_zvector", beta: complex):
        _chk(lib().nlk_zvec_axpby(C.c_double(alpha.real), C.c_double(alpha.imag), vec.re.h, vec.im.h,
                                  C.c_double(beta.real), C.c_double(beta.imag), self.re.h, self.im.h))

    def dot(self, vec: "nek_zvector") -> complex:
        a = C.c_double(); b = C.c_double()
        _chk(lib().nlk_zvec_dot(self.re.h, self.im.h, vec.re.h, vec.im.h, C.byref(a), C.byref(b)))
        return complex(a.value, b.value)

    def get_size(self):
        return self.re.get_size()          # complex_vectors.f90: size of one part


def svds(A: "exptA_linop", nsv, kdim, tol=0.0, x0=None, want_vectors=False):
    """LightKrylov `svds` as called at src/neklab_analysis.f90:136 (Golub-Kahan, device-resident U and V bases)."""
    sig = np.zeros(nsv); res = np.zeros(nsv); nit = C.c_int32(); info = C.c_int32()
    U = V = None; ua = va = None
    if want_vectors:
        U = [nek_dvector(A.ctx) for _ in range(nsv)]; V = [nek_dvector(A.ctx) for _ in range(nsv)]
        ua = (C.c_void_p * nsv)(*[u.h for u in U]); va = (C.c_void_p * nsv)(*[v.h for v in V])
    _chk(lib().nlk_svds(A.h, C.c_int32(nsv), C.c_int32(kdim), C.c_double(tol), x0.h if x0 is not None else None,
                        _p(sig), _p(res), ua, va, C.byref(nit), C.byref(info)))
    return dict(sigma=sig, resid=res, niter=nit.value, info=info.value, U=U, V=V)


class exptA_linop:
    """`exptA_linop` (src/linops/neklab_linops.f90:35-44): u'(tau) = exp(tau L) u'(0) by time stepping."""

    def __init__(self, ctx: Context, tau: float, baseflow: nek_dvector):
        self.ctx = ctx; self.tau = tau
        self.h = C.c_void_p()
        _chk(lib().nlk_exptA_create(ctx.h, C.c_double(tau), baseflow.h, C.byref(self.h)))

    def init(self):
        _chk(lib().nlk_exptA_init(self.h)); return self.stats()

    def matvec(self, vec_in, vec_out=None):
        vec_out = vec_out or nek_dvector(self.ctx)
        _chk(lib().nlk_exptA_matvec(self.h, vec_in.h, vec_out.h)); return vec_out

    def rmatvec(self, vec_in, vec_out=None):
        vec_out = vec_out or nek_dvector(self.ctx)
        _chk(lib().nlk_exptA_rmatvec(self.h, vec_in.h, vec_out.h)); return vec_out

    def time_steps(self, vec_in, nwarm, nsteps):
        ms = C.c_double()
        _chk(lib().nlk_exptA_time_steps(self.h, vec_in.h, C.c_int32(nwarm), C.c_int32(nsteps), C.byref(ms)))
        return ms.value

    def set_projection(self, alpha, idir=1):
        """turn this operator into `exptA_proj_linop(tau, baseflow, alpha)` (src/linops/neklab_linops.f90:130-152); idir = 0: off."""
        _chk(lib().nlk_exptA_set_projection(self.h, C.c_double(alpha), C.c_int32(idir)))

    def proj(self, vec):
        _chk(lib().nlk_exptA_apply_projection(self.h, vec.h)); return vec

    def stats(self):
        s = Stats(); _chk(lib().nlk_exptA_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib().nlk_exptA_destroy(self.h)
            self.h = None
        except Exception:
            pass


class resolvent_linop:
    """`resolvent_linop` (src/linops/neklab_linops.f90:198-205, src/linops/resolvent.f90): R(i omega) on a `nek_zvector`, evaluated
    by time stepping: forced response over one period, GMRES on I - exptA for the periodic state, a quarter period for the
    imaginary part."""

    def __init__(self, ctx: Context, omega: float, baseflow: nek_dvector, rtol: float = 0.0):
        self.ctx = ctx; self.omega = float(omega); self.rtol = rtol
        self._A = exptA_linop(ctx, 1.0, baseflow)
        self.info = None

    def _apply(self, vec_in: "nek_zvector", vec_out, adjoint):
        vec_out = vec_out or nek_zvector(self.ctx)
        info = C.c_int32()
        _chk(lib().nlk_resolvent_matvec(self._A.h, C.c_double(self.omega), vec_in.re.h, vec_in.im.h, vec_out.re.h, vec_out.im.h,
                                        C.c_int32(1 if adjoint else 0), C.c_double(self.rtol), C.byref(info)))
        self.info = info.value
        return vec_out

    def matvec(self, vec_in, vec_out=None):
        return self._apply(vec_in, vec_out, False)

    def rmatvec(self, vec_in, vec_out=None):
        return self._apply(vec_in, vec_out, True)

    def integrate(self, tau, forcing: "nek_zvector", x0: nek_dvector | None = None, adjoint=False) -> nek_dvector:
        """`evaluate_rhs` (x0 None, resolvent.f90:80-112) / `evaluate_imaginary_part` (:136-166)."""
        out = nek_dvector(self.ctx)
        _chk(lib().nlk_resolvent_integrate(self._A.h, C.c_double(tau), C.c_double(self.omega), forcing.re.h, forcing.im.h,
                                           C.c_int32(1 if adjoint else 0), x0.h if x0 is not None else None, out.h))
        return out


# ---- mesh-1 <-> mesh-2 pressure maps and field-file I/O (Nek `mappr` / `load_fld` / `outpost`; src/neklab_utils.f90:305-361)
def _tensor(M, a, ndim):
    a = np.einsum("pi,ezyi->ezyp", M, a)
    a = np.einsum("qj,ezjx->ezqx", M, a)
    return np.einsum("rk,ekyx->eryx", M, a) if ndim == 3 else a


def map12(mesh: Mesh, p1):
    """mesh 1 (GLL) -> mesh 2 (GL): evaluate the fld pressure at the Gauss points (exact inverse of `map21` for fld pressures)."""
    n = mesh.lx1
    return _tensor(mesh.basis("I12").reshape(n - 2, n), _f64(p1), mesh.ndim)


def map21(mesh: Mesh, p2):
    """mesh 2 (GL) -> mesh 1 (GLL): what Nek's `outpost` (prepost.f `mappr`) writes into the P block of a field file."""
    from .boxmesh import lagrange_interp
    return _tensor(lagrange_interp(mesh.basis("z1"), mesh.basis("z2")), _f64(p2), mesh.ndim)


def load_fld(ctx: Context, path) -> nek_dvector:
    """`call load_fld(file)` + `nek2vec(bf, vx, vy, vz, pr, t)` (examples/cylinder/stability/direct/1cyl.usr:15-16): read a Nek field
    file (elements in global-id order), map the pressure onto mesh 2 and upload everything into a device-resident `nek_dvector`."""
    from .formats import read_fld
    f = read_fld(path); m = ctx.mesh
    if f.vel is None or f.vel.shape[0] != m.nel or f.nx != m.lx1:
        raise NlkError(f"{path}: field file does not match the mesh ({f.nel} elements, nx = {f.nx})")
    v = ctx.vec()
    v.upload([f.vel[:, c] for c in range(m.ndim)], map12(m, f.pr) if f.pr is not None else None, f.temp)
    return v


def outpost_dnek(vec: nek_dvector, prefix: str, case: str, index: int = 1, outdir: str = ".", time: float = 0.0, istep: int = 0, wdsize: int = 8):
    """`outpost_dnek(vec, prefix)` (src/neklab_utils.f90:305-312 -> Nek `outpost`): writes `<prefix><case>0.f%05d` with the X, U, P
    (mesh 1) and T blocks, readable by Nek's `load_fld`, VisIt/ParaView and the reference's python tools.  Returns the path."""
    from .formats import write_fld
    if len(prefix) != 3:
        raise NlkError("outpost prefix must have 3 characters (Nek convention: 'dir', 'adj', 'BF_', 'nwt', ...)")
    m = vec.ctx.mesh
    v, pr, th = vec.download()
    path = os.path.join(outdir, "%s%s0.f%05d" % (prefix, case, index))
    write_fld(path, coords=np.stack(m._x, axis=1), vel=np.stack(v, axis=1), pr=map21(m, pr),
              temp=th if vec.ctx.params.ifheat else None, time=time, istep=istep, wdsize=wdsize)
    return path


def nonlinear_map(ctx: Context, tau, vec_in: nek_dvector, cfl_limit=0.4) -> nek_dvector:
    """`nek_system%response` (src/systems/fixed_point.f90:4-40): F_tau(X) - X."""
    out = nek_dvector(ctx)
    _chk(lib().nlk_nonlinear_map(ctx.h, C.c_double(tau), C.c_double(cfl_limit), vec_in.h, out.h))
    return out


class nek_upo_system:
    """`nek_upo_system%response` (src/systems/periodic_orbit.f90:4-44): F_T(X) - X with T the period carried by the vector,
    dt from the CFL of X at 0.4, vtol = ptol = atol*0.1; the period component of the residual is T - T = 0."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def response(self, vec_in: nek_ext_dvector, atol: float) -> nek_ext_dvector:
        ctx = self.ctx
        vt, pt = ctx.params.vtol, ctx.params.ptol
        ctx.set_tol(atol * 0.1, atol * 0.1)
        try:
            out = nek_ext_dvector(ctx)
            _chk(lib().nlk_nonlinear_map(ctx.h, C.c_double(vec_in.T), C.c_double(0.4), vec_in.vec.h, out.vec.h))
            out.T = vec_in.T - vec_in.T
        finally:
            ctx.set_tol(vt, pt)
        return out


class nek_upo_jacobian:
    """`nek_upo_jacobian` (src/systems/neklab_systems.f90:157-164; periodic_orbit.f90:46-181): Jacobian of the periodic-orbit
    residual at `X` = (X(0), T) on `nek_ext_dvector`s -- base flow and perturbation advanced together over T, the period column
    f'(X(T)) dT and the phase-condition row <dx, f'(X(0))>."""

    def __init__(self, ctx: Context, X: nek_ext_dvector):
        self.ctx = ctx; self.X = X

    def _apply(self, vec_in: nek_ext_dvector, vec_out, transpose):
        vec_out = vec_out or nek_ext_dvector(self.ctx)
        T = C.c_double()
        _chk(lib().nlk_upo_jacobian(self.ctx.h, self.X.vec.h, C.c_double(self.X.T), vec_in.vec.h, C.c_double(vec_in.T), vec_out.vec.h,
                                    C.byref(T), C.c_int32(1 if transpose else 0)))
        vec_out.T = T.value
        p = self.ctx.params; p.ptol = p.vtol                     # the call leaves vtol = ptol = atol, as the reference does
        return vec_out

    def matvec(self, vec_in, vec_out=None):
        return self._apply(vec_in, vec_out, False)

    def rmatvec(self, vec_in, vec_out=None):
        return self._apply(vec_in, vec_out, True)


def newton_fixed_point_iteration(ctx: Context, bf: nek_dvector, tol, tau=1.0, tol_mode=1, maxiter=40, gmres_kdim=30):
    """src/neklab_analysis.f90:158-212: Newton-GMRES for F_tau(X) = X; `bf` is updated in place."""
    hist = np.zeros(maxiter + 2); nit = C.c_int32(); info = C.c_int32()
    _chk(lib().nlk_newton_fixed_point(ctx.h, C.c_double(tau), bf.h, C.c_double(tol), C.c_int32(tol_mode), C.c_int32(maxiter),
                                      C.c_int32(gmres_kdim), _p(hist), C.byref(nit), C.byref(info)))
    return dict(residuals=hist[:nit.value + 1].copy(), niter=nit.value, info=info.value)


def eigs(A: exptA_linop, nev, kdim, tol=0.0, transpose=False, x0=None, want_vectors=False, callback=None):
    """LightKrylov `eigs` as called at src/neklab_analysis.f90:80-81 (device-resident Krylov basis)."""
    lr = np.zeros(nev); li = np.zeros(nev); rs = np.zeros(nev)
    niter = C.c_int32(); info = C.c_int32()
    vecs = None; arr = None
    if want_vectors:
        vecs = [nek_dvector(A.ctx) for _ in range(2 * nev)]
        arr = (C.c_void_p * (2 * nev))(*[v.h for v in vecs])
    hist = []

    def _cb(it, k, pre, pim, pres, user):
        re = np.ctypeslib.as_array(pre, shape=(k,)).copy(); im = np.ctypeslib.as_array(pim, shape=(k,)).copy()
        rr = np.ctypeslib.as_array(pres, shape=(k,)).copy()
        hist.append((it, k, re, im, rr))
        if callback:
            callback(it, k, re + 1j * im, rr)
    cb = EIGS_CB(_cb)
    _chk(lib().nlk_eigs(A.h, C.c_int32(nev), C.c_int32(kdim), C.c_double(tol), C.c_int32(int(transpose)),
                        x0.h if x0 is not None else None, _p(lr), _p(li), _p(rs), arr, C.byref(niter), cb, None, C.byref(info)))
    return dict(lam=lr + 1j * li, resid=rs, niter=niter.value, info=info.value, vecs=vecs, history=hist)


def linear_stability_analysis_fixed_point(A: exptA_linop, kdim, nev, adjoint=False, outdir=None, tol=0.0, x0=None, case=None):
    """src/neklab_analysis.f90:38-105: eigs, then lambda = log(mu)/tau (:84); writes `eigs_output.txt`
    (6 columns parsed by test/lib/neklabTestCase.py:413-455) and `dir_/adj_eigenspectrum.npy` (n x 3)."""
    lines = []

    def cb(it, k, lam, res):
        i = int(np.argmax(np.abs(lam)))
        lines.append("%6d %18.10E %18.10E %18.10E %18.10E %s" % (it, lam[i].real, lam[i].imag, abs(lam[i]), res[i], "T" if res[i] < (tol or 3.1622776601683794e-08) else "F"))
    r = eigs(A, nev, kdim, tol=tol, transpose=adjoint, x0=x0, callback=cb, want_vectors=bool(outdir and case))
    r["eigvals"] = np.log(r["lam"]) / A.tau
    if outdir:
        os.makedirs(outdir, exist_ok=True)
        if r.get("vecs") and case:                 # outpost_dnek(eigvecs, "dir"|"adj") (src/neklab_analysis.f90:93)
            for i in range(nev):
                outpost_dnek(r["vecs"][2 * i], "adj" if adjoint else "dir", case, 2 * i + 1, outdir)
                outpost_dnek(r["vecs"][2 * i + 1], "adj" if adjoint else "dir", case, 2 * i + 2, outdir)
        with open(os.path.join(outdir, "eigs_output.txt"), "w") as f:
            f.write("  iter            Re                 Im              modulus            residual     conv\n")
            f.write("\n".join(lines) + "\n")
        np.save(os.path.join(outdir, ("adj" if adjoint else "dir") + "_eigenspectrum.npy"),
                np.stack([r["eigvals"].real, r["eigvals"].imag, r["resid"]], axis=1))
    return r
