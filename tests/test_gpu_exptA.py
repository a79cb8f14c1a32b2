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


# ----------------------------------------------------------------------------------------------- exptA_temp_linop
def test_boussinesq_filter_parity(nlk_lib):
    """exptA_temp_linop path (src/linops/exponential_propagator_temp.f90; rayBen config): velocity + temperature,
    Boussinesq forcing through the declarative body force, explicit modal filter (q_filter)."""
    from neklab_b200 import api
    cbt = None
    om, bm, _ = box_case(ndim=2, nel=(5, 3), n=8, lxd=12, periodic=[True, False], hi=(4.0, 1.0), warp=False)
    cbc_t = om.cbc_v.copy(); cbc_t[cbc_t == "W  "] = "t  "
    from oracle.mesh import SEMesh
    om = SEMesh(om.coords, om.vertex, om.cbc_v, 12, cbc_t=cbc_t)
    kw = dict(viscosity=1.0, torder=3, vtol=1e-13, ptol=1e-13, ttol=1e-13, ifheat=True, conductivity=1.0, rhocp=1.0,
              buoyancy=(0.0, 1900.0, 0.0), filter_weight=0.01, filter_cutoff=0.84, gmres_maxit=2000, cg_maxit=2000)
    prm = StepParams(**kw)
    st = PertStepper(om, prm, precond=SchwarzCoarse(om))
    bf = NekVec(om, 3, ifheat=True)
    bf.theta = 1.0 - om.coords[:, 1]                       # conduction state (rayBen.usr userbc: temp = 1 - y)
    st.dt = 0.004                                          # zero base flow: dt is preset (setup_nek: recompute_dt disabled)
    A_or = ExptA(st, 0.02, bf)
    x0 = seeded_field(om, 21, ifheat=True)
    y_or = A_or.matvec(x0)
    m = api.Mesh(om.coords, om.vertex, om.cbc_v, 12, cbc_t=cbc_t)
    kw2 = dict(kw); kw2["ifheat"] = 1
    ctx = api.Context(m, api.default_params(**kw2))
    bd = ctx.vec(); bd.upload(bf.v, bf.pr, bf.theta)
    xd = ctx.vec(); xd.upload(x0.v, x0.pr, x0.theta)
    A = api.exptA_linop(ctx, 0.02, bd)
    api.lib().nlk_ctx_set_dt(ctx.h, __import__("ctypes").c_double(0.004))
    yd = A.matvec(xd)
    v, pr, th = yd.download()
    assert A.stats()["nsteps"] == st.nsteps == 5
    err_v = _wnorm(om, [v[c] - y_or.v[c] for c in range(2)]) / _wnorm(om, y_or.v)
    err_t = _wnorm(om, [th - y_or.theta]) / _wnorm(om, [y_or.theta])
    assert err_v < TOL_APPLY and err_t < TOL_APPLY, (err_v, err_t)
    assert yd.get_size() == y_or.size()
    assert abs(yd.dot(yd) - y_or.dot(y_or)) < 1e-9 * y_or.dot(y_or)      # theta enters the inner product
    ctx.close()


# ----------------------------------------------------------------------------------------------- reference config C4
def test_back_fstep_matvec_parity(nlk_lib):
    """examples/back_fstep/transient_growth from the committed fixture (2 760 gmsh elements, lx1 = 6, Re = 600, bdf2, explicit
    filter 0.01/0.84, walls 'W', inlet and outlet 'v', free-slip 'SYM' planes, singular pressure operator): a short direct and
    adjoint exptA apply (4 steps + 1 restart step) against the oracle with sparse-direct inner solves."""
    from neklab_b200 import api
    from tests.util import bfs_case
    om, bf, prm, z = bfs_case()
    prm.pressure_solver = "direct"; prm.helm_solver = "direct"
    st = PertStepper(om, prm)
    A_or = ExptA(st, 18.0, bf)
    dt, nsteps = A_or.init()
    assert nsteps == 1933
    A_or.tau = 4 * dt + 1e-12
    m = nlk_mesh(om)
    ctx = api.Context(m, api.default_params(viscosity=1.0 / 600.0, torder=2, vtol=1e-13, ptol=1e-13, filter_weight=0.01, filter_cutoff=0.84,
                                            gmres_maxit=3000, cg_maxit=3000, pr_proj=8))
    A = api.exptA_linop(ctx, 18.0, _to_dev(ctx, bf))
    s = A.init()
    assert s["nsteps"] == 1933 and abs(s["dt"] - dt) < 1e-15
    api.lib().nlk_exptA_set_tau(A.h, __import__("ctypes").c_double(4 * dt + 1e-12))
    x0 = seeded_field(om, 7, torder=2)
    for name in ("matvec", "rmatvec"):
        y_or = getattr(A_or, name)(x0)
        yd = getattr(A, name)(_to_dev(ctx, x0))
        v, pr, _ = yd.download()
        err = _wnorm(om, [v[c] - y_or.v[c] for c in range(2)]) / _wnorm(om, y_or.v)
        assert err < 1e-9, (name, err)                           # iterative (1e-13) vs sparse-direct inner solves
        assert yd.nrst == 1 == y_or.nrst
        assert A.stats()["nsteps"] == 4
        # free-slip planes: v_y = 0, v_x free
        top = np.isclose(om.coords[:, 1], 20.0) & (om.coords[:, 0] > -19.9) & (om.coords[:, 0] < 99.9)
        assert np.abs(v[1][top]).max() == 0.0 and np.abs(v[0][top]).max() > 0.0
    ctx.close()


def test_exptA_proj_linop(nlk_lib):
    """exptA_proj_linop (src/linops/exponential_propagator_proj.f90; examples/poiseuille/stability/direct_alpha_1/poiseuille.usr:24):
    the streamwise-wavenumber projection (planar average of 2 u cos / 2 u sin, bm1-weighted) around the time loop, matvec and
    rmatvec, on an x-periodic channel.  Properties: the projection is idempotent and keeps exactly the alpha-harmonic."""
    from neklab_b200 import api
    from oracle.cref import CPertStepper
    from oracle.stepper import ExptAProj
    om, _, _ = box_case(ndim=2, nel=(4, 4), n=6, lxd=9, periodic=[True, False], warp=False, hi=(2 * np.pi, 2.0))
    x = om.coords
    bf = NekVec(om, 2); bf.v = [1.0 - (x[:, 1] - 1.0) ** 2, np.zeros_like(x[:, 1])]
    kw = dict(viscosity=0.01, torder=2, vtol=1e-13, ptol=1e-13, gmres_maxit=2000, cg_maxit=2000)
    alpha = 1.0
    A_or = ExptAProj(CPertStepper(om, StepParams(**kw), precond=SchwarzCoarse(om)), 0.3, bf, alpha, idir=1)
    ctx = api.Context(nlk_mesh(om), api.default_params(**kw))
    A = api.exptA_linop(ctx, 0.3, _to_dev(ctx, bf)); A.set_projection(alpha, 1)
    x0 = seeded_field(om, 5, torder=2)
    # the projection itself
    pv = A.proj(_to_dev(ctx, x0)).download()[0]
    po = A_or.proj(x0.v)
    for c in range(2):
        assert rel(pv[c], po[c]) < 1e-12
    pp = A_or.proj(po)
    assert rel(pp[0], po[0]) < 1e-12                                               # idempotent
    h = [np.cos(alpha * x[:, 0]) * (1 - (x[:, 1] - 1) ** 2), np.sin(alpha * x[:, 0]) * x[:, 1] * (2 - x[:, 1])]
    hp = A_or.proj(h)
    assert rel(hp[0], h[0]) < 1e-10 and rel(hp[1], h[1]) < 1e-10                   # the alpha-harmonic passes unchanged
    for tr in (False, True):
        y_or = A_or._apply(x0, tr)
        y = (A.rmatvec if tr else A.matvec)(_to_dev(ctx, x0))
        v, _, _ = y.download()
        num = np.sqrt(sum(float(((v[c] - y_or.v[c]) ** 2 * om.bm1).sum()) for c in range(2)))
        assert num / _wnorm(om, y_or.v) < TOL_APPLY
    ctx.close()
