"""GPU parity on the reference's OWN configurations at full length (BASELINE.json configs[1..4]), through the C-ABI, against
the C++/OpenMP oracle (oracle/cpp/nekref.cpp; itself checked against the numpy oracle in tests/test_oracle_cpp.py):

  * cylinder Re=50 (examples/cylinder/stability/direct): one full exptA matvec = 100 + 2 time steps, 1e-10 (north_star), and
    the reference-produced base flow as a fixed point of the nonlinear flow map (KAT-7 on the GPU path);
  * Rayleigh-Benard (examples/rayBen/baseflow fixtures): exptA_temp_linop matvec AND rmatvec;
  * backward-facing step (examples/back_fstep/transient_growth): 60 time steps direct and adjoint (bdf2, filter, SYM);
  * the synthetic 3-D extruded cylinder of bench.py (lx1 = 8, lxd = 12; a 864-element window, three periodic z-layers): the stepper
    itself, not only its kernels, against the oracle.
Inner tolerances are tightened on both sides (setup_nek takes vtol/ptol: src/neklab_nek_setup.f90:74-75) as SURVEY.md §7
prescribes for 1e-10 parity of iterative solvers."""
import os

import numpy as np
import pytest

from oracle import ops
from oracle.cref import CPertStepper
from oracle.mesh import SEMesh
from oracle.precond import SchwarzCoarse
from oracle.stepper import ExptA, NekVec, StepParams, nonlinear_map, seeded_field
from tests.util import GOLDEN, bfs_case, cylinder_case, nlk_mesh

pytestmark = pytest.mark.gpu

TOL_APPLY = 1e-10


def _wdiff(om, dev_vec, ref: NekVec, heat=False):
    """relative bm1-weighted L2 distance over the fields of nek_ddot (velocity [+ temperature])."""
    v, pr, th = dev_vec.download()
    num = sum(float(((v[c] - ref.v[c]) ** 2 * om.bm1).sum()) for c in range(om.ndim))
    den = sum(float((ref.v[c] ** 2 * om.bm1).sum()) for c in range(om.ndim))
    if heat:
        num += float(((th - ref.theta) ** 2 * om.bm1).sum()); den += float((ref.theta ** 2 * om.bm1).sum())
    return np.sqrt(num / den)


def _dev(ctx, nv: NekVec):
    d = ctx.vec(); d.upload(nv.v, nv.pr, nv.theta); return d


def _rst_diff(om, dev_vec, ref: NekVec, k, heat=False):
    r = dev_vec.get_rst(k)
    o = NekVec(om, ref.torder, heat); o.v, o.pr, o.theta = ref.rst[k - 1]
    return _wdiff(om, r, o, heat)


# ---------------------------------------------------------------------------------------------------------------- cylinder
@pytest.fixture(scope="module")
def cyl(nlk_lib):
    from neklab_b200 import api
    om, bf, prm, z = cylinder_case()
    pre = SchwarzCoarse(om)
    tight = dict(vtol=1e-13, ptol=1e-12, gmres_maxit=2000, cg_maxit=2000)
    prm = StepParams(viscosity=1 / 50.0, torder=3, **tight)
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=1 / 50.0, torder=3, **tight))
    yield om, bf, prm, pre, ctx
    ctx.close()


def test_cylinder_full_length_matvec_parity(cyl):
    """exptA_matvec on the reference's golden-test config at full length: tau = 1 -> 100 steps + 2 restart steps."""
    from neklab_b200 import api
    om, bf, prm, pre, ctx = cyl
    A_or = ExptA(CPertStepper(om, prm, precond=pre), 1.0, bf)
    A = api.exptA_linop(ctx, 1.0, _dev(ctx, bf))
    x0 = seeded_field(om, 3)
    y_or = A_or.matvec(x0)
    y = A.matvec(_dev(ctx, x0))
    st = A.stats()
    assert st["nsteps"] == 100 and st["steps"] == 102 and abs(st["dt"] - 0.01) < 1e-15
    assert _wdiff(om, y, y_or) < TOL_APPLY
    assert y.nrst == 2
    for k in (1, 2):
        assert _rst_diff(om, y, y_or, k) < TOL_APPLY
    # second apply: the input carries restart fields (exptA_get_rst overrides the first two steps)
    z_or = A_or.matvec(y_or); z = A.matvec(y)
    assert _wdiff(om, z, z_or) < 10 * TOL_APPLY


def test_cylinder_shipped_base_flow_is_a_fixed_point(cyl):
    """KAT-7 through the C-ABI: the reference-produced BF_1cyl0.f00001 is a fixed point of nlk_nonlinear_map
    (nek_system%response, src/systems/fixed_point.f90:4-40) to the reference's Newton tolerance."""
    from neklab_b200 import api
    om, bf, prm, pre, ctx = cyl
    r = api.nonlinear_map(ctx, 1.0, _dev(ctx, bf), cfl_limit=0.5)
    assert r.norm() < 5e-6
    r_or = nonlinear_map(CPertStepper(om, prm, precond=pre), bf, 1.0, 0.5)
    v, _, _ = r.download()
    err = np.sqrt(sum(float(((v[c] - r_or.v[c]) ** 2 * om.bm1).sum()) for c in range(2))) / bf.norm()
    assert err < TOL_APPLY


# ---------------------------------------------------------------------------------------------------------------- rayBen
def _rayben():
    z = np.load(os.path.join(GOLDEN, "rayben_case.npz"))
    om = SEMesh(z["coords"], z["vertex"], z["cbc"], 15, cbc_t=z["cbc_t"])
    Pr, Ra = 1.0, 1900.0                                        # rayBen.par userParam05/06; rayBen.usr usrdat2 + userf
    kw = dict(viscosity=np.sqrt(Pr / Ra), conductivity=1.0 / np.sqrt(Pr * Ra), torder=3, ifheat=True, filter_weight=0.01, filter_cutoff=0.84)
    bf = NekVec(om, 3, True)
    bf.v = [z["vel"][:, 0].copy(), z["vel"][:, 1].copy()]
    bf.pr = ops.map12(om, z["pr"]); bf.theta = z["temp"].copy()
    return om, bf, kw, (0.0, Ra * Pr, 0.0)


@pytest.mark.parametrize("mode", ["literal", "conduction", "conduction_adjoint"])
def test_rayben_fixture_parity(nlk_lib, mode):
    """exptA_temp_linop (exponential_propagator_temp.f90) on examples/rayBen/baseflow/{rayBen.re2, rayBen.ma2, BF_rayBen0.f00001}.
    'literal': the shipped base velocity is ~8e-5, so setup_nek's CFL rule takes ONE step of dt = tau (+ 2 restart steps);
    'conduction': base velocity exactly zero -> recompute_dt is disabled (neklab_nek_setup.f90:79-83) and the .par dt = 0.01 is used."""
    from neklab_b200 import api
    om, bf, kw, buoy = _rayben()
    tight = dict(vtol=1e-13, ptol=1e-12, ttol=1e-13, gmres_maxit=2000, cg_maxit=2000)
    tau = 0.2
    if mode != "literal":
        bf.v = [np.zeros_like(bf.v[0]), np.zeros_like(bf.v[1])]
    prm = StepParams(buoyancy=buoy, **kw, **tight)
    st = CPertStepper(om, prm, precond=SchwarzCoarse(om)); st.dt = 0.01
    A_or = ExptA(st, tau, bf)
    p = api.default_params(buoyancy=buoy, **{k: (int(v) if k == "ifheat" else v) for k, v in kw.items()}, **tight)
    ctx = api.Context(nlk_mesh(om), p); ctx.set_dt(0.01)
    A = api.exptA_linop(ctx, tau, _dev(ctx, bf))
    x0 = seeded_field(om, 7, True)
    tr = mode.endswith("adjoint")
    y_or = A_or._apply(x0, tr)
    y = (A.rmatvec if tr else A.matvec)(_dev(ctx, x0))
    assert A.stats()["nsteps"] == (1 if mode == "literal" else 20) == st.nsteps
    assert _wdiff(om, y, y_or, heat=True) < TOL_APPLY
    for k in (1, 2):
        assert _rst_diff(om, y, y_or, k, heat=True) < 10 * TOL_APPLY
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------- back_fstep
@pytest.mark.parametrize("transpose", [False, True])
def test_back_fstep_60_steps(nlk_lib, transpose):
    """examples/back_fstep/transient_growth (bfs.usr:18-21): 60 of the 1933 steps of one exptA (r)matvec, bdf2 + explicit
    filter + 'SYM' planes + all-Dirichlet pressure operator (mean-free solution of the bordered system)."""
    from neklab_b200 import api
    om, bf, prm, z = bfs_case()
    tight = dict(vtol=1e-13, ptol=1e-12, gmres_maxit=3000, cg_maxit=3000)
    prm = StepParams(viscosity=1 / 600.0, torder=2, filter_weight=0.01, filter_cutoff=0.84, **tight)
    ctarg = ops.compute_cfl(om, bf.v, 1.0); dt0 = 0.5 / ctarg
    tau = 59.5 * dt0
    A_or = ExptA(CPertStepper(om, prm, precond=SchwarzCoarse(om)), tau, bf)
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=1 / 600.0, torder=2, filter_weight=0.01, filter_cutoff=0.84, **tight))
    A = api.exptA_linop(ctx, tau, _dev(ctx, bf))
    x0 = seeded_field(om, 5, torder=2)
    y_or = A_or._apply(x0, transpose)
    y = (A.rmatvec if transpose else A.matvec)(_dev(ctx, x0))
    assert A.stats()["nsteps"] == 60
    assert _wdiff(om, y, y_or) < 10 * TOL_APPLY
    assert _rst_diff(om, y, y_or, 1) < 10 * TOL_APPLY
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------- synth3d
def test_synth3d_stepper_parity(nlk_lib):
    """The throughput config of bench.py (3-D extruded cylinder, lx1 = 8, lxd = 12): a window of 288 near-cylinder elements
    x 3 periodic z-layers (864 curved elements, 442 368 points).  3 time steps + 2 restart steps of exptA against the C++
    oracle -- stepper-level parity on the geometry, order and dealiasing every perf number is quoted on."""
    from neklab_b200 import api
    from tests.util import synth3d_window
    om, U = synth3d_window()
    assert om.E == 864 and om.n == 8 and om.m == 12 and om.has_outflow
    tight = dict(vtol=1e-13, ptol=1e-12, gmres_maxit=2000, cg_maxit=2000)
    prm = StepParams(viscosity=1 / 50.0, torder=3, **tight)
    bf = NekVec(om, 3); bf.v = U
    ctarg = ops.compute_cfl(om, bf.v, 1.0); tau = 2.5 * 0.5 / ctarg
    A_or = ExptA(CPertStepper(om, prm, precond=SchwarzCoarse(om)), tau, bf)
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=1 / 50.0, torder=3, **tight))
    A = api.exptA_linop(ctx, tau, _dev(ctx, bf))
    x = om.coords
    x0 = seeded_field(om, 3)
    x0.v[2] = om.vmask[2] * om.dssum(np.sin(0.5 * x[:, 0]) * np.cos(2 * np.pi * x[:, 2] / 1.5) * 0.5) * om.vmult      # genuinely 3-D
    y_or = A_or.matvec(x0)
    y = A.matvec(_dev(ctx, x0))
    assert A.stats()["nsteps"] == 3 and A.stats()["steps"] == 5
    assert _wdiff(om, y, y_or) < TOL_APPLY
    for k in (1, 2):
        assert _rst_diff(om, y, y_or, k) < TOL_APPLY
    # adjoint on the same mesh (convect_adj at lx1 = 8 / lxd = 12)
    assert _wdiff(om, A.rmatvec(_dev(ctx, x0)), A_or.rmatvec(x0)) < TOL_APPLY
    ctx.close()
