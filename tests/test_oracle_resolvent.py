"""CPU checks of the oracle restatement of `resolvent_linop` (src/linops/resolvent.f90): the real part it returns is the
time-periodic state of the harmonically forced linearised flow, and the quarter-period leg is consistent with it."""
import numpy as np

from oracle.cref import CPertStepper
from oracle.precond import SchwarzCoarse
from oracle.stepper import NekVec, Resolvent, StepParams
from tests.util import box_case


def resolvent_case(torder=3):
    om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "})
    x = om.coords
    bf = NekVec(om, torder); bf.v = [4.0 * x[:, 1] * (3.0 - x[:, 1]) / 9.0 + 0 * x[:, 0], 0.05 * np.sin(0.5 * x[:, 0]) * np.sin(np.pi * x[:, 1] / 3.0)]
    g = lambda x0, y0, s: np.exp(-((x[:, 0] - x0) ** 2 + (x[:, 1] - y0) ** 2) / (2 * s * s))
    fre = NekVec(om, torder); fre.v = [g(1.0, 1.5, 0.4), 0.5 * g(1.5, 1.0, 0.5)]
    fim = NekVec(om, torder); fim.v = [0.3 * g(2.0, 2.0, 0.4), -0.7 * g(1.2, 1.8, 0.5)]
    kw = dict(viscosity=0.05, torder=torder, vtol=1e-13, ptol=1e-13, gmres_maxit=2000, cg_maxit=2000)
    return om, bf, fre, fim, kw


def wdiff(om, a, b):
    return np.sqrt(sum(float(((a.v[c] - b.v[c]) ** 2 * om.bm1).sum()) for c in range(2)) / sum(float((b.v[c] ** 2 * om.bm1).sum()) for c in range(2)))


def test_resolvent_real_part_is_the_periodic_state():
    """torder = 1 (no rst fields, so the GMRES operator is exactly the map the check integrates): forced integration over one
    period started from `re` returns `re` to the GMRES tolerance; the adjoint differs from the direct response."""
    om, bf, fre, fim, kw = resolvent_case(torder=1)
    st = CPertStepper(om, StepParams(**kw), precond=SchwarzCoarse(om))
    omega = 2.0
    R = Resolvent(st, omega, bf, rtol=1e-9)
    re, im = R.matvec(fre, fim)
    tau = 2 * np.pi / omega
    again = R.integrate(tau, fre, fim, re)
    assert wdiff(om, again, re) < 1e-7
    # the quarter-period leg is what `im` is
    q = R.integrate(tau / 4, fre, fim, re)
    assert wdiff(om, q, im) < 1e-13
    # linearity in the forcing: R(2 f) = 2 R(f)
    f2r = fre.copy(); f2r.scal(2.0); f2i = fim.copy(); f2i.scal(2.0)
    re2, im2 = R.matvec(f2r, f2i)
    re.scal(2.0); im.scal(2.0)
    assert wdiff(om, re2, re) < 1e-7 and wdiff(om, im2, im) < 1e-7
    rea, _ = R.rmatvec(fre, fim)
    rea.scal(2.0)
    assert wdiff(om, rea, re) > 1e-2
