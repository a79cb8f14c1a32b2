"""GPU parity of the full hot path: one exptA matvec (time stepping + rst protocol) and nek_dvector semantics
against the numpy oracle (north_star: single apply within 1e-10 relative, mass-weighted L2)."""
import numpy as np
import pytest

from oracle import ops
from oracle.precond import SchwarzCoarse
from oracle.stepper import ExptA, NekVec, PertStepper, StepParams, seeded_field
from tests.util import box_case, nlk_mesh, rel, smooth_fields

pytestmark = pytest.mark.gpu

TOL_APPLY = 1e-10      # north_star tolerance for a single exptA apply (relative, bm1-weighted L2)


def _baseflow(om):
    x = om.coords
    bf = NekVec(om, 3)
    if om.ndim == 2:
        bf.v = [1.0 + 0.3 * np.sin(0.5 * x[:, 1]), 0.2 * np.cos(0.4 * x[:, 0])]
    else:
        bf.v = [1.0 + 0.3 * np.sin(0.5 * x[:, 1]), 0.2 * np.cos(0.4 * x[:, 0]), 0.1 * np.sin(0.3 * x[:, 0] + 0.2 * x[:, 2])]
    return bf


def _wnorm(om, v):
    return np.sqrt(sum(float((a * a * om.bm1).sum()) for a in v))


def _to_dev(ctx, nv: NekVec):
    d = ctx.vec()
    d.upload(nv.v, nv.pr, nv.theta)
    return d


CASES = {
    "box2d_outflow_bdf3": dict(mesh=dict(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "}), torder=3, tau=0.25),
    "box2d_periodic_bdf2": dict(mesh=dict(ndim=2, nel=(4, 4), n=6, lxd=9, periodic=[True, False]), torder=2, tau=0.2),
    "box3d_bdf3": dict(mesh=dict(ndim=3, nel=(3, 2, 2), n=5, lxd=8, bc={"xlo": "v  ", "xhi": "O  "}), torder=3, tau=0.1),
    "box2d_outflow_bdf3_proj": dict(mesh=dict(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "}), torder=3, tau=0.25, proj=8),
}


@pytest.fixture(scope="module", params=list(CASES))
def setup(request, nlk_lib):
    from neklab_b200 import api
    cfg = CASES[request.param]
    om, _, _ = box_case(**cfg["mesh"])
    m = nlk_mesh(om)
    nu = 0.05
    prm = StepParams(viscosity=nu, torder=cfg["torder"], vtol=1e-13, ptol=1e-13, gmres_maxit=2000, cg_maxit=2000)
    pc = SchwarzCoarse(om)
    st = PertStepper(om, prm, precond=pc)
    bf = _baseflow(om)
    A_or = ExptA(st, cfg["tau"], bf)
    ctx = api.Context(m, api.default_params(viscosity=nu, torder=cfg["torder"], vtol=1e-13, ptol=1e-13, gmres_maxit=2000, cg_maxit=2000,
                                            pr_proj=cfg.get("proj", 0)))
    A_dev = api.exptA_linop(ctx, cfg["tau"], _to_dev(ctx, bf))
    yield request.param, om, ctx, A_or, A_dev, cfg
    ctx.close()


def test_dt_nsteps(setup):
    name, om, ctx, A_or, A_dev, cfg = setup
    dt, nsteps = A_or.init()
    s = A_dev.init()
    assert s["nsteps"] == nsteps
    assert abs(s["dt"] - dt) < 1e-15


def test_matvec_and_rst_parity(setup):
    name, om, ctx, A_or, A_dev, cfg = setup
    x0 = seeded_field(om, 3, torder=cfg["torder"])
    y_or = A_or.matvec(x0)
    xd = _to_dev(ctx, x0)
    yd = A_dev.matvec(xd)
    v, pr, _ = yd.download()
    err = _wnorm(om, [v[c] - y_or.v[c] for c in range(om.ndim)]) / _wnorm(om, y_or.v)
    assert err < TOL_APPLY, err
    assert yd.nrst == cfg["torder"] - 1 == y_or.nrst
    for k in range(1, yd.nrst + 1):
        rv, rp, _ = yd.get_rst(k).download()
        ro = y_or.get_rst(k)
        assert _wnorm(om, [rv[c] - ro.v[c] for c in range(om.ndim)]) / _wnorm(om, ro.v) < TOL_APPLY
    # second application consumes the rst fields of its input (exptA_get_rst)
    z_or = A_or.matvec(y_or)
    zd = A_dev.matvec(yd)
    v2, _, _ = zd.download()
    err2 = _wnorm(om, [v2[c] - z_or.v[c] for c in range(om.ndim)]) / _wnorm(om, z_or.v)
    assert err2 < TOL_APPLY, err2
    # divergence-free to solver tolerance
    assert np.abs(ops.opdiv(om, v2)).max() < 1e-10


def test_rmatvec_parity(setup):
    name, om, ctx, A_or, A_dev, cfg = setup
    x0 = seeded_field(om, 5, torder=cfg["torder"])
    y_or = A_or.rmatvec(x0)
    yd = A_dev.rmatvec(_to_dev(ctx, x0))
    v, _, _ = yd.download()
    err = _wnorm(om, [v[c] - y_or.v[c] for c in range(om.ndim)]) / _wnorm(om, y_or.v)
    assert err < TOL_APPLY, err


def test_vector_semantics(setup):
    name, om, ctx, A_or, A_dev, cfg = setup
    a = seeded_field(om, 11, torder=cfg["torder"]); b = seeded_field(om, 12, torder=cfg["torder"])
    a.pr = np.random.default_rng(1).standard_normal(om.bm2.shape); b.pr = np.random.default_rng(2).standard_normal(om.bm2.shape)
    ad, bd = _to_dev(ctx, a), _to_dev(ctx, b)
    assert abs(ad.dot(bd) - a.dot(b)) < 1e-12 * abs(a.dot(b))          # bm1-weighted, pressure excluded
    assert ad.get_size() == a.size()
    # rst quirk of nek_daxpby: rst slots receive alpha * (x's CURRENT fields)
    a.save_rst(b, 1); ad.save_rst(bd, 1)
    a.axpby(0.7, b, -1.3); ad.axpby(0.7, bd, -1.3)
    v, pr, _ = ad.download()
    assert rel(v[0], a.v[0]) < 1e-14 and rel(pr, a.pr) < 1e-14
    rv, rp, _ = ad.get_rst(1).download()
    ro = a.get_rst(1)
    assert rel(rv[1], ro.v[1]) < 1e-14 and rel(rp, ro.pr) < 1e-14
    ad.zero(); assert ad.nrst == 0 and ad.norm() == 0.0
