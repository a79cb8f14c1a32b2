"""ORACLE (test infrastructure, not product): Python driver of the C++/OpenMP restatement oracle/cpp/nekref.cpp.

`CPertStepper` is a drop-in for `oracle.stepper.PertStepper` (same attributes and methods, so `oracle.stepper.ExptA`,
`oracle.krylov.*` and `nonlinear_map` work unchanged) whose `advance()` runs in C++ on all host cores.  Geometry,
numbering, masks and the preconditioner's setup data are taken from the numpy oracle (`oracle.mesh.SEMesh`,
`oracle.precond.SchwarzCoarse`); only the per-step arithmetic is restated in C++ (see the header of nekref.cpp for the
reference routines it follows).  Used by tests/ (full-length parity applies on the reference's configs) and by
bench.py's `cpu_baseline` / `--impl reference` arm.  Never imported from neklab_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import ops
from .mesh import SEMesh
from .stepper import StepParams, filter_matrix

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cpp", "nekref.cpp")
_LIB = os.path.join(_HERE, "cpp", "libnekref.so")
_lib = None


def build(force=False):
    """g++ -O3 -fopenmp (x86-64-v3 so the prebuilt .so also loads on the GPU box's host)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["g++", "-O3", "-mavx2", "-mfma", "-fopenmp", "-shared", "-fPIC", "-std=c++17", _SRC, "-o", _LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/cpp build failed:\n" + r.stderr)
    return _LIB


class _Desc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("E", C.c_int64), ("nglob", C.c_int64)] + \
        [(k, C.c_void_p) for k in ("D", "I12", "D12", "I1d", "Dd", "F1d", "G", "bm1", "binvm1", "vmult", "bm2", "rxw2", "rxd",
                                   "mask0", "mask1", "mask2", "mask3", "diagA", "gidx")] + \
        [("has_outflow", C.c_int32), ("volvm1", C.c_double), ("volvm2", C.c_double)] + \
        [(k, C.c_void_p) for k in ("S", "dinv", "wt", "A0inv", "shape", "vertex")] + [("nv", C.c_int64)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.nekref_create.restype = C.c_void_p
        _lib.nekref_threads.restype = C.c_int
        _lib.nekref_helmholtz.restype = C.c_int
        _lib.nekref_pressure.restype = C.c_int
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CRef:
    """Handle on one nekref context (mesh + optional SchwarzCoarse preconditioner)."""

    def __init__(self, mesh: SEMesh, prm: StepParams, precond=None):
        self.mesh, self.prm = mesh, prm
        d, b = mesh.ndim, mesh.b
        order = [(0, 0), (1, 1), (0, 1)] if d == 2 else [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
        keep = self._keep = {}
        keep["G"] = _c(np.stack([mesh.G[k][l] for k, l in order], axis=1))
        keep["rxw2"] = _c(np.stack([mesh.rx2[k][c] * mesh.W2 for k in range(d) for c in range(d)], axis=1))
        keep["rxd"] = _c(np.stack([mesh.rxd[k][c] for k in range(d) for c in range(d)], axis=1))
        for name, a in (("D", b.D), ("I12", b.I12), ("D12", b.D12), ("I1d", b.I1d), ("Dd", b.Dd),
                        ("F1d", filter_matrix(b, prm.filter_weight, prm.filter_cutoff)), ("bm1", mesh.bm1), ("binvm1", mesh.binvm1),
                        ("vmult", mesh.vmult), ("bm2", mesh.bm2), ("diagA", ops.axhelm_diag(mesh, 1.0, 0.0))):
            keep[name] = _c(a)
        for c in range(d):
            keep["mask%d" % c] = _c(mesh.vmask[c])
        if mesh.tmask is not None:
            keep["mask3"] = _c(mesh.tmask)
        keep["gidx"] = np.ascontiguousarray(mesh.gidx, dtype=np.int64)
        if precond is not None:
            keep["S"] = _c(precond.S); keep["dinv"] = _c(precond.dinv); keep["wt"] = _c(precond.wt)
            if precond.use_coarse:
                keep["A0inv"] = _c(precond.A0inv); keep["shape"] = _c(np.stack(precond.shape, axis=0))
                keep["vertex"] = np.ascontiguousarray(mesh.vertex, dtype=np.int64)
        ds = _Desc()
        ds.ndim, ds.n, ds.m, ds.E, ds.nglob = d, mesh.n, mesh.m, mesh.E, mesh.nglob
        for k, _ in _Desc._fields_:
            if k in keep:
                setattr(ds, k, _p(keep[k]))
        ds.has_outflow = int(mesh.has_outflow); ds.volvm1 = mesh.volvm1; ds.volvm2 = mesh.volvm2
        ds.nv = int(precond.nv) if (precond is not None and precond.use_coarse) else 0
        self.h = C.c_void_p(lib().nekref_create(C.byref(ds)))
        self.set_params(prm)

    def set_params(self, prm: StepParams, variant=0):
        self.prm = prm
        bu = (C.c_double * 3)(*prm.buoyancy)
        lib().nekref_set_params(self.h, C.c_double(prm.viscosity), C.c_double(prm.density), C.c_int(prm.torder), C.c_double(prm.vtol),
                                C.c_double(prm.ptol), C.c_int(int(prm.ifheat)), C.c_double(prm.conductivity), C.c_double(prm.rhocp),
                                C.c_double(prm.ttol), bu, C.c_double(prm.filter_weight), C.c_int(prm.cg_maxit), C.c_int(prm.gmres_maxit),
                                C.c_int(prm.lgmres), C.c_int(variant))

    # ---- operator-level hooks
    def axhelm(self, u, h1, h2):
        u = _c(u); w = np.zeros_like(u); lib().nekref_axhelm(self.h, _p(u), C.c_double(h1), C.c_double(h2), _p(w)); return w

    def dssum(self, u):
        u = _c(u).copy(); lib().nekref_dssum(self.h, _p(u)); return u

    def opdiv(self, u):
        u = [_c(x) for x in u]; p = np.zeros_like(self.mesh.bm2)
        lib().nekref_opdiv(self.h, _p(u[0]), _p(u[1]), _p(u[2]) if len(u) > 2 else None, _p(p)); return p

    def opgradt(self, p):
        p = _c(p); w = [np.zeros_like(self.mesh.bm1) for _ in range(self.mesh.ndim)]
        lib().nekref_opgradt(self.h, _p(p), _p(w[0]), _p(w[1]), _p(w[2]) if len(w) > 2 else None); return w

    def convect(self, u, Cv):
        u = _c(u); Cv = [_c(x) for x in Cv]; o = np.zeros_like(u)
        lib().nekref_convect(self.h, _p(u), _p(Cv[0]), _p(Cv[1]), _p(Cv[2]) if len(Cv) > 2 else None, _p(o)); return o

    def convect_adj(self, U, cf):
        d = self.mesh.ndim
        U = [_c(x) for x in U]; cf = [_c(x) for x in cf]; o = [np.zeros_like(self.mesh.bm1) for _ in range(d)]
        arr = lambda lst: (C.c_void_p * 3)(*[(_p(x).value if x is not None else None) for x in (lst + [None] * 3)[:3]])
        lib().nekref_convect_adj(self.h, arr(U), arr(cf), arr(o)); return o

    def cdabdtp(self, p):
        p = _c(p); o = np.zeros_like(p); lib().nekref_cdabdtp(self.h, _p(p), _p(o)); return o

    def precond(self, r):
        r = _c(r); o = np.zeros_like(r); lib().nekref_precond(self.h, _p(r), _p(o)); return o

    def helmholtz(self, f, h1, h2, comp, tol):
        f = _c(f); x = np.zeros_like(f)
        it = lib().nekref_helmholtz(self.h, _p(f), C.c_double(h1), C.c_double(h2), C.c_int(comp), C.c_double(tol), _p(x)); return x, it

    def pressure(self, rhs, tol):
        rhs = _c(rhs); x = np.zeros_like(rhs); it = lib().nekref_pressure(self.h, _p(rhs), C.c_double(tol), _p(x)); return x, it

    @staticmethod
    def use_all_cores():
        """Let the C++ oracle use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        lib().nekref_set_threads(C.c_int(n))
        return n

    def threads(self):
        return int(lib().nekref_threads())

    def counters(self):
        a = C.c_int64(); b = C.c_int64(); lib().nekref_counters(self.h, C.byref(a), C.byref(b)); return a.value, b.value

    def __del__(self):
        try:
            if self.h:
                lib().nekref_destroy(self.h); self.h = None
        except Exception:
            pass


class CPertStepper:
    """Same interface as oracle.stepper.PertStepper; the state lives in the C++ context."""

    def __init__(self, mesh: SEMesh, prm: StepParams, precond=None, variant=0, pr_proj=0):
        self.mesh, self.prm, self.d = mesh, prm, mesh.ndim
        self.ref = CRef(mesh, prm, precond)
        lib().nekref_set_proj(self.ref.h, C.c_int(pr_proj))
        if variant:
            self.ref.set_params(prm, variant)
        self.adjoint = False; self.nonlinear = False
        self.U = [np.zeros_like(mesh.bm1) for _ in range(self.d)]
        self.T = np.zeros_like(mesh.bm1)
        self.forcing = None
        self.dt = None; self.nsteps = None
        self.stats = {}

    def setup(self, tau, cfl_limit=0.5, transpose=False):
        import math
        self.adjoint = transpose
        self._dirty = True
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

    def _push_mode(self):
        L = lib(); U = [_c(u) for u in self.U]; T = _c(self.T)
        L.nekref_set_base(self.ref.h, _p(U[0]), _p(U[1]), _p(U[2]) if self.d == 3 else None, _p(T))
        L.nekref_set_mode(self.ref.h, C.c_double(self.dt), C.c_int(int(self.adjoint)), C.c_int(int(self.nonlinear)))
        self._push_forcing()

    def _push_forcing(self):
        L = lib()
        if self._forcing is not None:
            f = [_c(x) for x in self._forcing]
            L.nekref_set_forcing(self.ref.h, _p(f[0]), _p(f[1]), _p(f[2]) if self.d == 3 else None)
        else:
            L.nekref_set_forcing(self.ref.h, None, None, None)
        self._fdirty = False

    # a forcing assigned between two steps (time-dependent forcing: resolvent, OTD) is pushed before the next advance
    @property
    def forcing(self):
        return self._forcing

    @forcing.setter
    def forcing(self, f):
        self._forcing = f; self._fdirty = True

    def set_state(self, v, p, t=None):
        v = [_c(x) for x in v]; p = _c(p); t = _c(t) if t is not None else np.zeros_like(self.mesh.bm1)
        lib().nekref_set_state(self.ref.h, _p(v[0]), _p(v[1]), _p(v[2]) if self.d == 3 else None, _p(p), _p(t))

    def reset_history(self):
        lib().nekref_reset_history(self.ref.h)

    def _get(self):
        m = self.mesh
        v = [np.zeros_like(m.bm1) for _ in range(self.d)]; p = np.zeros_like(m.bm2); t = np.zeros_like(m.bm1)
        lib().nekref_get_state(self.ref.h, _p(v[0]), _p(v[1]), _p(v[2]) if self.d == 3 else None, _p(p), _p(t))
        return v, p, t

    vp = property(lambda self: self._get()[0])
    prp = property(lambda self: self._get()[1])
    tp = property(lambda self: self._get()[2])

    def advance(self, istep):
        if getattr(self, "_dirty", True):          # base flow / mode / forcing are pushed once per setup()
            self._push_mode(); self._dirty = False
        elif self._fdirty:
            self._push_forcing()
        lib().nekref_advance(self.ref.h, C.c_int(istep))
        cg, gm = self.ref.counters()
        self.stats["cg_iters"] = cg; self.stats["gmres_iters"] = gm
