"""CPU tests: the C++/OpenMP oracle (oracle/cpp/nekref.cpp) against the numpy oracle, operator by operator and through
whole exptA applies (2-D / 3-D, bdf2 / bdf3, direct / adjoint, Boussinesq, filter).  The numpy oracle is the one pinned
by the known-answer tests (tests/test_oracle_kat.py); the C++ one exists so that full-length applies on the reference's
own configs and the same-config CPU baseline of bench.py are affordable."""
import numpy as np
import pytest

from oracle import ops
from oracle.cref import CPertStepper, CRef
from oracle.precond import SchwarzCoarse
from oracle.stepper import ExptA, NekVec, PertStepper, StepParams, seeded_field
from tests.util import box_case, rel, smooth_fields


def _case(ndim):
    if ndim == 2:
        om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "})
    else:
        om, _, _ = box_case(ndim=3, nel=(2, 2, 3), n=5, lxd=8, bc={"xlo": "v  ", "xhi": "O  "}, periodic=(False, False, True))
    return om


@pytest.mark.parametrize("ndim", [2, 3])
def test_operators_match_numpy(ndim):
    om = _case(ndim)
    pre = SchwarzCoarse(om)
    R = CRef(om, StepParams(viscosity=0.05), pre)
    f = smooth_fields(om, 2 * ndim + 1, seed=3)
    u = f[0]
    assert rel(R.axhelm(u, 0.3, 2.0), ops.axhelm(om, u, 0.3, 2.0)) < 1e-13
    assert rel(R.dssum(u), om.dssum(u)) < 1e-14
    U = f[1:1 + ndim]; C = f[1 + ndim:1 + 2 * ndim]
    assert rel(R.opdiv(U), ops.opdiv(om, U)) < 1e-13
    p = ops.opdiv(om, U)
    for a, b in zip(R.opgradt(p), ops.opgradt(om, p)):
        assert rel(a, b) < 1e-13
    assert rel(R.convect(u, C), ops.convect_new(om, u, C)) < 1e-13
    for a, b in zip(R.convect_adj(U, C), ops.convect_adj(om, U, C)):
        assert rel(a, b) < 1e-13
    assert rel(R.cdabdtp(p), ops.cdabdtp(om, p)) < 1e-12
    assert rel(R.precond(p), pre(p)) < 1e-11
    x, it = R.helmholtz(u, 0.05, 30.0, 0, 1e-12)
    from oracle.stepper import cggo, uzawa_gmres
    assert rel(x, cggo(om, u.copy(), 0.05, 30.0, om.vmask[0], 1e-12, 1000)) < 1e-9
    rhs = ops.ortho(om, p)
    xp, itp = R.pressure(rhs, 1e-12)
    xo = uzawa_gmres(om, rhs, lambda q: ops.cdabdtp(om, q), pre, 1e-12, 100, 30)
    assert rel(xp, xo) < 1e-8


CASES = [
    dict(ndim=2, torder=3, transpose=False),
    dict(ndim=2, torder=2, transpose=True, filt=True),
    dict(ndim=3, torder=3, transpose=False),
    dict(ndim=2, torder=3, transpose=False, heat=True),
    dict(ndim=2, torder=3, transpose=True, heat=True),
]


@pytest.mark.parametrize("cs", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_exptA_matches_numpy(cs):
    ndim = cs["ndim"]; heat = cs.get("heat", False)
    if heat:
        cbt = np.array([["t  ", "I  ", "t  ", "I  "]] * 12)
        om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, cbc_t=cbt)
    else:
        om = _case(ndim)
    x = om.coords
    prm = StepParams(viscosity=0.05, torder=cs["torder"], vtol=1e-13, ptol=1e-13, ttol=1e-13, gmres_maxit=2000, ifheat=heat,
                     conductivity=0.07, buoyancy=(0.0, 3.0, 0.0) if heat else (0.0, 0.0, 0.0),
                     filter_weight=0.05 if cs.get("filt") else 0.0, filter_cutoff=0.7 if cs.get("filt") else 1.0)
    bf = NekVec(om, cs["torder"], heat)
    bf.v = [1.0 + 0.3 * np.sin(0.5 * x[:, 1]), 0.2 * np.cos(0.4 * x[:, 0])] + ([0.1 * np.sin(0.3 * x[:, 0])] if ndim == 3 else [])
    if heat:
        bf.v = [om.vmask[c] * bf.v[c] for c in range(2)]
        bf.theta = 1.0 - x[:, 1] / 3.0 + 0.1 * np.sin(x[:, 0])
    x0 = seeded_field(om, 5, heat, cs["torder"])
    pre = SchwarzCoarse(om)
    if heat and cs["transpose"]:
        return _adjoint_heat_identity()
    y_np = ExptA(PertStepper(om, prm, precond=pre), 0.05, bf)._apply(x0, cs["transpose"])
    y_c = ExptA(CPertStepper(om, prm, precond=pre), 0.05, bf)._apply(x0, cs["transpose"])
    den = y_np.norm()
    diff = y_c.copy(); diff.nrst = 0; diff.axpby(-1.0, y_np, 1.0)
    assert diff.norm() / den < 1e-10
    assert rel(y_c.pr, y_np.pr) < 1e-7
    if heat:
        assert rel(y_c.theta, y_np.theta) < 1e-9
    for k in range(cs["torder"] - 1):
        for c in range(ndim):
            assert rel(y_c.rst[k][0][c], y_np.rst[k][0][c]) < 1e-9


def _adjoint_heat_identity():
    """exptA_temp_linop%rmatvec (exponential_propagator_temp.f90:62-107): the numpy oracle has no adjoint Boussinesq step, so
    the C++ one is checked through the adjoint identity <A x, y>_B = <x, A^+ y>_B on a closed box with smooth solenoidal
    fields (agreement to the temporal discretisation error: 1.9e-2 at CFL 0.5 with 5 steps, 2.9e-3 at CFL 0.125)."""
    cbt = np.array([["t  ", "t  ", "t  ", "t  "]] * 16)
    om, _, _ = box_case(ndim=2, nel=(4, 4), n=8, lxd=12, warp=False, cbc_t=cbt)
    X, Y = om.coords[:, 0], om.coords[:, 1]; k = np.pi / 4; s, c = np.sin, np.cos
    pre = SchwarzCoarse(om)

    def products(buoy, Tamp):
        prm = StepParams(viscosity=0.05, torder=3, vtol=1e-12, ptol=1e-12, ttol=1e-12, gmres_maxit=1000, ifheat=True, conductivity=0.07, buoyancy=(0.0, buoy, 0.0))
        bf = NekVec(om, 3, True)
        bf.v = [s(k * X) ** 2 * 2 * k * s(k * Y) * c(k * Y), -2 * k * s(k * X) * c(k * X) * s(k * Y) ** 2]      # psi = sin^2 sin^2: solenoidal, zero on the walls
        bf.theta = Tamp * (1.0 - Y / 2.0 + 0.5 * s(k * X) * s(2 * k * Y))

        def sf(a, b, ph):
            v = NekVec(om, 3, True)
            v.v = [s(a * k * X) ** 2 * 2 * b * k * s(b * k * Y) * c(b * k * Y) + ph * s(k * X) ** 2 * 4 * k * s(2 * k * Y) * c(2 * k * Y),
                   -(2 * a * k * s(a * k * X) * c(a * k * X) * s(b * k * Y) ** 2 + ph * 2 * k * s(k * X) * c(k * X) * s(2 * k * Y) ** 2)]
            v.theta = s(b * k * X) * s(a * k * Y) + ph * s(k * X) * s(k * Y)
            return v
        u, v = sf(1, 2, 0.3), sf(2, 1, -0.4)
        st = CPertStepper(om, prm, precond=pre); orig = st.setup
        st.setup = lambda tau, cfl, tr: orig(tau, 0.125, tr)
        A = ExptA(st, 0.2, bf)
        return A.matvec(u).dot(v), u.dot(A.rmatvec(v))
    lhs, rhs = products(6.0, 1.0)
    assert abs(lhs - rhs) < 5e-3 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    lhs0, _ = products(0.0, 0.0)
    assert abs(lhs0 - lhs) > 0.05 * abs(lhs)          # the coupling terms matter in this set-up
