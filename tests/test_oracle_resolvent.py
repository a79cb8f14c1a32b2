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


def test_resolvent_against_the_frequency_domain_solve():
    """Physics anchor that does not go through the time stepper: the semi-discrete linearised operator (same weak forms: mass,
    dealiased convection about U, viscous stiffness, D and D^T) is assembled densely on the unique unmasked dofs and the harmonic
    response solved directly,
        (i omega B + rho C(U) + nu A) q - D^T p = B f ,   D q = 0 .
    The time-stepping resolvent must return Re(q) in `re` and -- because `im` is the periodic state a quarter period later,
    Re(q e^{i pi/2}) -- MINUS Im(q) in `im` (resolvent.f90:40-41 with alpha = exp(+i omega t), :95): (re, im) = conj(q), up to the
    O((omega dt)^2..3) error of the BDF/EXT scheme."""
    from oracle import ops
    om, bf, fre, fim, kw = resolvent_case(torder=3)
    st = CPertStepper(om, StepParams(**kw), precond=SchwarzCoarse(om))
    omega = 2.0
    re, im = Resolvent(st, omega, bf, rtol=1e-9).matvec(fre, fim)
    # ---- dense semi-discrete operator on global dofs
    d, ng = 2, om.nglob
    shape = om.bm1.shape
    loc = lambda g: g[om.gidx].reshape(shape)                                  # continuous global vector -> local copies
    gat = lambda l: np.bincount(om.gidx, weights=l.ravel(), minlength=ng)      # Q^T (dssum to unique dofs)
    mg = [np.zeros(ng) for _ in range(d)]
    for c in range(d):
        mg[c][om.gidx] = om.vmask[c].ravel()
    free = [np.where(mg[c] > 0)[0] for c in range(d)]
    nf = [len(f) for f in free]; off = [0, nf[0]]; nu_ = nf[0] + nf[1]
    n2 = om.bm2.size
    nu, rho = kw["viscosity"], 1.0
    K = np.zeros((nu_, nu_)); Mb = np.zeros(nu_); Dm = np.zeros((n2, nu_))
    zero = np.zeros(shape)
    for cj in range(d):
        for jj, g in enumerate(free[cj]):
            e = np.zeros(ng); e[g] = 1.0
            up = [zero, zero]; up[cj] = loc(e)
            col = off[cj] + jj
            for c in range(d):
                r = rho * ops.convect_new(om, bf.v[c], up)                       # u'.grad U_c
                if c == cj:
                    r = r + rho * ops.convect_new(om, up[c], bf.v) + ops.axhelm(om, up[c], nu, 0.0)
                K[off[c]:off[c] + nf[c], col] = gat(r)[free[c]]
            Mb[col] = gat(om.bm1 * up[cj])[g]
            Dm[:, col] = ops.opdiv(om, up).ravel()
    # D^T really is the transpose of D in these weak forms (one spot check)
    p = np.random.default_rng(0).standard_normal(om.bm2.shape)
    gt = ops.opgradt(om, p)
    assert np.abs(np.concatenate([gat(gt[c])[free[c]] for c in range(d)]) - Dm.T @ p.ravel()).max() < 1e-12
    fhat = np.concatenate([(gat(om.bm1 * fre.v[c]) + 1j * gat(om.bm1 * fim.v[c]))[free[c]] for c in range(d)])
    A = np.zeros((nu_ + n2, nu_ + n2), dtype=complex)
    A[:nu_, :nu_] = K + 1j * omega * rho * np.diag(Mb)
    A[:nu_, nu_:] = -Dm.T
    A[nu_:, :nu_] = Dm
    sol = np.linalg.solve(A, np.concatenate([fhat, np.zeros(n2)]))
    q = sol[:nu_]
    got_re = np.concatenate([gat(re.v[c] * om.vmult)[free[c]] for c in range(d)])
    got_im = np.concatenate([gat(im.v[c] * om.vmult)[free[c]] for c in range(d)])
    w = np.sqrt(Mb)
    nrm = np.linalg.norm(w * np.abs(q))
    e_re = np.linalg.norm(w * (got_re - q.real)) / nrm
    e_conj = np.linalg.norm(w * (got_im + q.imag)) / nrm
    e_plain = np.linalg.norm(w * (got_im - q.imag)) / nrm
    assert e_re < 1e-2 and e_conj < 1e-2, (e_re, e_conj)                          # measured: 1.4e-3 and 3.2e-3
    assert e_plain > 0.3, e_plain                                                 # it is the conjugate, not q itself
