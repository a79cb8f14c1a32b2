"""CPU tests of the oracle's own consistency: direct vs iterative solvers, projection property, time-order,
nek_dvector semantics, Krylov-Schur on a small operator."""
import numpy as np
import pytest

from oracle import ops
from oracle.krylov import eigs
from oracle.precond import SchwarzCoarse
from oracle.stepper import ExptA, NekVec, PertStepper, StepParams, ab_coeffs, bdf_coeffs, seeded_field
from tests.util import box_case


@pytest.fixture(scope="module")
def small():
    om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "})
    x = om.coords
    bf = NekVec(om, 3)
    bf.v = [1.0 + 0.3 * np.sin(0.5 * x[:, 1]), 0.2 * np.cos(0.4 * x[:, 0])]
    return om, bf


def test_time_scheme_coefficients():
    assert bdf_coeffs(3) == [11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0]
    assert ab_coeffs(3, 3) == [3.0, -3.0, 1.0]
    assert np.allclose(ab_coeffs(3, 2), [8.0 / 3.0, -7.0 / 3.0, 2.0 / 3.0])
    for nab in (1, 2, 3):
        for nbd in (1, 2, 3):
            assert abs(sum(ab_coeffs(nab, nbd)) - 1.0) < 1e-14


def test_iterative_equals_direct_and_divergence_free(small):
    om, bf = small
    x0 = seeded_field(om, 1)
    outs = []
    for ps, hs in (("direct", "direct"), ("gmres", "cg")):
        prm = StepParams(viscosity=0.05, torder=3, vtol=1e-13, ptol=1e-13, pressure_solver=ps, helm_solver=hs, gmres_maxit=2000)
        st = PertStepper(om, prm, precond=SchwarzCoarse(om))
        y = ExptA(st, 0.1, bf).matvec(x0)
        assert np.abs(ops.opdiv(om, y.v)).max() < 1e-11
        outs.append(y)
    for c in range(2):
        assert np.abs(outs[0].v[c] - outs[1].v[c]).max() < 1e-10
    assert outs[0].nrst == 2


def test_adjoint_is_close_to_discrete_transpose():
    # closed box (no outflow boundary term), smooth fields: <A u, v>_B = <u, A^+ v>_B up to discretisation error
    om, _, _ = box_case(ndim=2, nel=(4, 4), n=8, lxd=12, warp=False)
    x = om.coords
    bf = NekVec(om, 3)
    bf.v = [om.vmask[0] * np.sin(np.pi * x[:, 1] / 4) * np.cos(np.pi * x[:, 0] / 4), -om.vmask[1] * np.cos(np.pi * x[:, 1] / 4) * np.sin(np.pi * x[:, 0] / 4)]
    prm = StepParams(viscosity=0.05, torder=3, vtol=1e-12, ptol=1e-12, gmres_maxit=1000)
    A = ExptA(PertStepper(om, prm, precond=SchwarzCoarse(om)), 0.2, bf)

    def smooth(k):
        v = NekVec(om, 3)
        v.v = [om.vmask[0] * np.sin(np.pi * x[:, 0] / 4 * k) * np.sin(np.pi * x[:, 1] / 4), om.vmask[1] * np.sin(np.pi * x[:, 0] / 4) * np.sin(np.pi * x[:, 1] / 4 * k)]
        return v
    u, v = smooth(1), smooth(2)
    lhs = A.matvec(u).dot(v); rhs = u.dot(A.rmatvec(v))
    assert abs(lhs - rhs) < 5e-3 * max(abs(lhs), abs(rhs), 1e-3)


def test_nek_dvector_semantics(small):
    om, bf = small
    a = seeded_field(om, 4); b = seeded_field(om, 5)
    a.pr[:] = 1.0; b.pr[:] = 2.0
    d0 = a.dot(b)
    a.pr[:] = 7.0
    assert a.dot(b) == d0                               # pressure excluded from dot (real_vectors.f90:217-227)
    a.save_rst(b, 1)
    a.axpby(2.0, b, 3.0)                                # rst slot gets beta*rst + alpha*b.CURRENT
    assert np.allclose(a.rst[0][0][0], 3.0 * b.v[0] + 2.0 * b.v[0])
    assert np.allclose(a.pr, 3.0 * 7.0 + 2.0 * 2.0)     # axpby includes pressure (:176)
    a.zero(); assert a.nrst == 0


def test_krylov_schur_small_matrix():
    rng = np.random.default_rng(0)
    n = 60
    M = np.diag(np.linspace(0.1, 1.0, n)) + 0.01 * rng.standard_normal((n, n))

    class V:
        def __init__(s, a): s.a = np.array(a, float)
        def copy(s): return V(s.a)
        def zero(s): s.a[:] = 0
        def scal(s, c): s.a *= c
        def axpby(s, al, o, be): s.a = al * o.a + be * s.a
        def dot(s, o): return float(s.a @ o.a)
        def norm(s): return float(np.linalg.norm(s.a))
    lam, res, X, Y, k = eigs(lambda v: V(M @ v.a), V(rng.standard_normal(n)), nev=3, kdim=20, tol=1e-10, maxiter=50)
    ref = np.linalg.eigvals(M); ref = ref[np.argsort(-np.abs(ref))]
    assert np.abs(np.sort(np.abs(lam[:3])) - np.sort(np.abs(ref[:3]))).max() < 1e-8


def test_direct_mode_without_outflow_solves_the_bordered_system():
    """No outflow: the pressure operator has the constants in its (near-)null space and uzawa_gmres moves only in mean-free
    directions; the sparse-direct mode must give the same converged limit -- on a straight-sided box (E exactly singular) and
    on a warped one, where the weak divergence is integrated inexactly and E 1 != 0 (the back_fstep situation)."""
    from oracle.stepper import uzawa_gmres
    for warp in (False, True):
        om, _, _ = box_case(ndim=2, nel=(4, 4), n=6, lxd=9, warp=warp)          # walls all round
        assert not om.has_outflow
        rng = np.random.default_rng(3)
        u = [om.vmask[c] * rng.standard_normal(om.bm1.shape) for c in range(2)]
        rhs = ops.ortho(om, -ops.opdiv(om, u))
        st = PertStepper(om, StepParams(viscosity=0.05, pressure_solver="direct"))
        xd = st._solve_pressure(rhs, 0.0)
        assert abs(xd.mean()) < 1e-13
        xg = uzawa_gmres(om, rhs, lambda p: ops.cdabdtp(om, p, 1.0), SchwarzCoarse(om), 1e-14, 3000, 30, {})
        assert np.abs(xd - xg).max() < 1e-9 * np.abs(xg).max(), warp
        r = ops.cdabdtp(om, xd, 1.0) - rhs
        assert np.abs(ops.ortho(om, r)).max() < 1e-10 * np.abs(rhs).max()       # residual is a multiple of the constant
