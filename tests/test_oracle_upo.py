"""CPU checks of the oracle restatement of the periodic-orbit Jacobian (src/systems/periodic_orbit.f90): with the base flow
advanced together with the perturbation, the field block is the tangent of the discrete nonlinear flow map, the period column
is f'(X(T)) and the phase row is <dx, f'(X(0))>."""
import numpy as np

from oracle.cref import CPertStepper
from oracle.precond import SchwarzCoarse
from oracle.stepper import NekExtVec, NekVec, StepParams, UPOJacobian
from tests.util import box_case


def upo_case(torder=3):
    om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "})
    x = om.coords
    prof = 4.0 * x[:, 1] * (3.0 - x[:, 1]) / 9.0
    X = NekVec(om, torder)
    X.v = [prof * (1.0 + 0.15 * np.sin(0.8 * x[:, 0]) * om.vmask[0]), 0.1 * om.vmask[1] * np.sin(1.1 * x[:, 0]) * np.sin(np.pi * x[:, 1] / 3.0)]
    X.v = [om.dssum(v) * om.vmult for v in X.v]
    dx = NekVec(om, torder)
    dx.v = [om.vmask[0] * np.sin(0.9 * x[:, 0]) * np.sin(2 * np.pi * x[:, 1] / 3.0), om.vmask[1] * np.cos(0.7 * x[:, 0]) * np.sin(np.pi * x[:, 1] / 3.0) ** 2]
    dx.v = [om.dssum(v) * om.vmult for v in dx.v]
    kw = dict(viscosity=0.05, torder=torder, vtol=1e-12, ptol=1e-12, gmres_maxit=2000, cg_maxit=2000)
    return om, X, dx, kw


def steppers(om, kw):
    pre = SchwarzCoarse(om)
    return CPertStepper(om, StepParams(**kw), precond=pre), CPertStepper(om, StepParams(**kw), precond=pre)


def wnorm(om, v):
    return np.sqrt(sum(float((a * a * om.bm1).sum()) for a in v))


def test_upo_jacobian_is_the_tangent_of_the_flow_map():
    om, X, dx, kw = upo_case()
    lin, nl = steppers(om, kw)
    T = 0.2
    J = UPOJacobian(lin, nl, NekExtVec(X, T))
    vin = NekExtVec(dx.copy(), 0.0)
    out = J.matvec(vin)
    dt, nsteps = nl.dt, nl.nsteps

    def flow(x0):                                     # the nonlinear map with the SAME dt and step count
        nl.nonlinear = True; nl.adjoint = False; nl._dirty = True
        nl.set_state(x0.v, x0.pr, x0.theta); nl.reset_history()
        for i in range(1, nsteps + 1):
            nl.advance(i)
        return [a.copy() for a in nl.vp]
    # tolerances as inside the Jacobian (atol*0.1): finite differences need them tight
    nl.prm.vtol = nl.prm.ptol = 1e-13; nl.ref.set_params(nl.prm, 0)
    eps = 1e-5
    xp = X.copy(); xm = X.copy()
    for c in range(2):
        xp.v[c] = X.v[c] + eps * dx.v[c]; xm.v[c] = X.v[c] - eps * dx.v[c]
    fp, fm = flow(xp), flow(xm)
    tang = [(a - b) / (2 * eps) for a, b in zip(fp, fm)]
    # out = u'(T) - dx  (T_in = 0: no period column)
    got = [out.vec.v[c] + dx.v[c] for c in range(2)]
    err = wnorm(om, [got[c] - tang[c] for c in range(2)]) / wnorm(om, tang)
    assert err < 1e-7, err
    assert nl.dt == dt
    # phase row: <dx, f'(X(0))>, f' by one first-order step
    nl.set_state(X.v, X.pr, X.theta); nl.reset_history(); nl.advance(1)
    fd = [(a - b) / dt for a, b in zip(nl.vp, X.v)]
    assert abs(out.T - sum(float((dx.v[c] * fd[c] * om.bm1).sum()) for c in range(2))) < 1e-8 * max(1.0, abs(out.T))
    assert abs(out.T) > 1e-6
    # period column: J (0, dT) = dT * f'(X(T)) and f'(X(T)) approximates (F_{T+h} - F_T)/h
    z = NekVec(om, 3)
    col = J.matvec(NekExtVec(z, 1.0))
    assert col.T == 0.0
    xT = flow(X)
    nl.advance(nsteps + 1)                            # one more step of the same scheme
    fdT = [(a - b) / dt for a, b in zip(nl.vp, xT)]
    cosang = sum(float((col.vec.v[c] * fdT[c] * om.bm1).sum()) for c in range(2)) / (wnorm(om, col.vec.v) * wnorm(om, fdT))
    assert cosang > 0.95, cosang                      # same vector up to O(dt): first-order restart vs BDF3 continuation, and ...
    # ... the reference takes f' where the base trajectory stands after the nrst extra steps of compute_rst (periodic_orbit.f90:90-97)
    nl.advance(nsteps + 2)
    x2 = [a.copy() for a in nl.vp]; p2 = nl.prp.copy()
    nl.set_state(x2, p2); nl.reset_history(); nl.advance(1)
    fq = [(a - b) / dt for a, b in zip(nl.vp, x2)]
    assert wnorm(om, [col.vec.v[c] - fq[c] for c in range(2)]) / wnorm(om, fq) < 1e-9
