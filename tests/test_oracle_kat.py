"""Known-answer tests of the ORACLE on the reference's shipped cylinder fixture.

tests/golden/cylinder_case.npz was made from /root/reference/examples/cylinder/stability/direct/
{BF_1cyl0.f00001,1cyl.ma2,1cyl.re2} by the snippet recorded in tests/golden/README.md.
"""
import json
import os

import numpy as np
import pytest

from oracle import ops
from oracle.mesh import glo_num_from_coords, partition_rank, same_partition
from tests.util import GOLDEN, cylinder_case


@pytest.fixture(scope="module")
def cyl():
    return cylinder_case()


def test_kat1_weak_divergence_of_baseflow(cyl):
    om, bf, prm, z = cyl
    div = ops.opdiv(om, bf.v)
    assert np.abs(div / om.bm2).max() < 1e-9            # SURVEY: max 7.3e-10


def test_kat2_pressure_is_degree_lx2(cyl):
    om, bf, prm, z = cyl
    back = ops.map21(om, ops.map12(om, z["pr"]))
    assert np.abs(back - z["pr"]).max() < 1e-14


def test_kat3_steady_momentum_residual(cyl):
    om, bf, prm, z = cyl
    U = bf.v; nu = 1.0 / 50.0
    conv = [ops.convect_new(om, U[c], U) for c in range(2)]
    gp = ops.opgradt(om, bf.pr)
    res = [om.vmask[c] * om.dssum(-conv[c] + gp[c] - ops.axhelm(om, U[c], nu, 0.0)) for c in range(2)]
    cn = [om.vmask[c] * om.dssum(conv[c]) for c in range(2)]
    nrm = lambda v: np.sqrt(sum((x * x * om.vmult).sum() for x in v))
    r = nrm(res) / nrm(cn)
    assert 5e-6 < r < 1e-5                               # SURVEY: 7.3e-6 with lxd=9 dealiasing (5.4e-4 without)


def test_kat4_numbering_matches_coordinates(cyl):
    om, bf, prm, z = cyl
    assert om.nglob == 50089
    assert sorted(np.unique(om.mult_count).tolist()) == [1, 2, 4, 5]
    gc = glo_num_from_coords(om.coords, periods={1: (-16.0, 32.0)})
    assert same_partition(om.glo, gc)


def test_kat5_cfl_dt_nsteps(cyl):
    om, bf, prm, z = cyl
    ctarg = ops.compute_cfl(om, bf.v, 1.0)
    assert abs(ctarg - 49.72) < 0.01
    dt = 0.5 / ctarg
    assert int(np.ceil(1.0 / dt)) == 100                 # fld header: time 1.0, istep 101


def test_kat6_partition_balance(cyl):
    om, bf, prm, z = cyl
    assert np.bincount(partition_rank(z["pid"], 2)).tolist() == [998, 998]
    assert np.bincount(partition_rank(z["pid"], 4)).tolist() == [499] * 4
    assert np.bincount(partition_rank(z["pid"], 8)).tolist() == [249, 250] * 4


def test_golden_eigenvalue_pin():
    """The reference's only golden number: |lambda_1| = 1.0156 +- 1e-4 (test/neklabTests.py:44).

    The oracle run with restart-field arithmetic exactly as written in real_vectors.f90:186-200 gives 1.01782
    (2.2e-3 away); with consistent restart combinations it gives 1.01578 (1.8e-4 away).  Neither meets the
    reference's own 1e-4 (its harness cannot run as shipped), so parity on this number is reported as
    PARTIAL; what is asserted here is the documented distance.
    """
    lit = json.load(open(os.path.join(GOLDEN, "cylinder_eig_oracle.json")))
    assert abs(lit["modulus"][0] - 1.0156) < 3e-3
    assert lit["resid"][0] < 3.2e-8
    p = os.path.join(GOLDEN, "cylinder_eig_oracle_consistent.json")
    if os.path.exists(p):
        con = json.load(open(p))
        assert abs(con["modulus"][0] - 1.0156) < 2.5e-4


@pytest.fixture(scope="module")
def cyl_pre(cyl):
    from oracle.precond import SchwarzCoarse
    return SchwarzCoarse(cyl[0])


def test_kat7_shipped_base_flow_is_a_fixed_point_of_the_restated_nonlinear_map(cyl, cyl_pre):
    """KAT-7 (reference-PRODUCED data): `BF_1cyl0.f00001` is the output of the reference's Newton solver on
    F_tau(X) - X = 0 (src/neklab_analysis.f90:158-212; tolerance 1e-6 in the shipped Newton example,
    examples/cylinder/newton/Re40_fixed_point/1cyl.usr:29; the file header says time = 1, istep = 101: written after a
    100-step tau = 1 integration).  The restated NONLINEAR stepper (C++ oracle, tight inner tolerances) must therefore
    leave it fixed: ||F_1(X) - X||_bm1 = 3.1e-6 (6.7e-8 relative) -- the same order as the reference's Newton tolerance.
    This pins dealiased convection, Helmholtz, PN-PN-2 projection, masks and geometry against Nek5000 output."""
    from oracle.cref import CPertStepper
    from oracle.stepper import StepParams, nonlinear_map
    om, bf, prm, z = cyl
    st = CPertStepper(om, StepParams(viscosity=1 / 50.0, torder=3, vtol=1e-12, ptol=1e-11, gmres_maxit=400), precond=cyl_pre)
    r = nonlinear_map(st, bf, 1.0, 0.5)
    assert st.nsteps == 100
    assert r.norm() < 5e-6 and r.norm() / bf.norm() < 1e-7


@pytest.mark.skipif(not os.environ.get("NLK_LONG_TESTS"), reason="~1 min of CPU: NLK_LONG_TESTS=1")
def test_kat8_exptA_is_the_jacobian_of_the_nonlinear_map(cyl, cyl_pre):
    """KAT-8: exptA applied to v equals the central finite difference of the nonlinear flow map about the shipped
    (reference-produced) fixed point, same dt: 1.4e-7 relative -- the linearised stepper is the Jacobian of a nonlinear
    discretisation that reproduces Nek5000's fixed point (KAT-7)."""
    from oracle.cref import CPertStepper
    from oracle.stepper import ExptA, StepParams, nonlinear_map, seeded_field
    om, bf, prm, z = cyl
    st = CPertStepper(om, StepParams(viscosity=1 / 50.0, torder=3, vtol=1e-13, ptol=1e-12, gmres_maxit=600), precond=cyl_pre)
    A = ExptA(st, 1.0, bf)
    v = A.matvec(seeded_field(om, 11)); v.nrst = 0; v.rst = [None, None]; v.scal(1.0 / v.norm())
    Av = A.matvec(v)
    eps = 1e-4
    out = []
    for sg in (+1, -1):
        x = bf.copy(); x.axpby(sg * eps, v, 1.0)
        r = nonlinear_map(st, x, 1.0, 0.5); r.axpby(1.0, x, 1.0)
        out.append(r)
    J = out[0]; J.axpby(-1.0, out[1], 1.0); J.scal(0.5 / eps); J.axpby(-1.0, Av, 1.0)
    assert J.norm() / Av.norm() < 1e-6


def test_golden_sweep_record():
    """tests/golden/cylinder_golden_sweep_r02.json: B200 runs of examples/golden_sweep.py over the start-up / pressure
    bookkeeping degrees of freedom.  Every variant whose restart arithmetic is a fixed linear map lands in
    1.01578 ... 1.01581 (dt/2: 1.015754); the literal arithmetic gives 1.0182.  None reaches 1.0156 +- 1e-4."""
    rec = json.load(open(os.path.join(GOLDEN, "cylinder_golden_sweep_r02.json")))
    sane = [r for r in rec if r["rst_mode"] in (1, 2) and r["cfl"] == 0.5]
    assert len(sane) >= 10
    assert max(r["modulus"] for r in sane) - min(r["modulus"] for r in sane) < 4e-5
    assert all(1.7e-4 < r["modulus"] - 1.0156 < 2.1e-4 for r in sane)
    assert all(r["resid"] < 3.2e-8 for r in rec)
    lit = [r for r in rec if r["rst_mode"] == 0]
    assert all(abs(r["modulus"] - 1.0182) < 1e-4 for r in lit)


def test_direct_eigenvalue_record():
    """tests/golden/cylinder_direct_eig.json (oracle/cylinder_direct_eig.py): the leading eigenvalue of the SEMI-DISCRETE
    linearised operator on the reference's cylinder mesh and base flow by a direct sparse shift-invert solve (no time stepper).
    The time-stepped Ritz values of the sweep converge onto it -- consistent rst at CFL 0.5 is 6.5e-5 away, dt/2 closer -- and it
    sits 2.7e-5 above the golden's acceptance window: DESIGN.md 1.1 item 3a."""
    d = json.load(open(os.path.join(GOLDEN, "cylinder_direct_eig.json")))
    mu = complex(*d["exp_lambda"][0]); lam = complex(*d["lambda"][0])
    assert d["residuals"][0] < 1e-12 and abs(np.exp(lam) - mu) < 1e-14
    assert abs(abs(mu) - 1.0157273) < 1e-6
    assert all(abs(v["exp_lambda_modulus"] - abs(mu)) < 1e-6 for v in d["sensitivity"].values())      # un-dealiased convection: same value
    rec = json.load(open(os.path.join(GOLDEN, "cylinder_golden_sweep_r02.json")))
    cons = [r for r in rec if r["rst_mode"] == 1 and r["step_variant"] == 0 and r["torder"] == 3 and r["cfl"] == 0.5][0]
    assert abs(complex(cons["lam_re"], cons["lam_im"]) - mu) < 7e-5                      # whole temporal error of the stepper at CFL 0.5
    half = [r for r in rec if r["rst_mode"] == 1 and r["cfl"] == 0.25]
    assert half and all(abs(r["modulus"] - abs(mu)) < abs(cons["modulus"] - abs(mu)) for r in half)      # dt/2 converges towards it
    assert 2e-5 < abs(mu) - 1.0157 < 4e-5                                                # above the window [1.0155, 1.0157]
    lit = [r for r in rec if r["rst_mode"] == 0][0]
    assert abs(lit["modulus"] - abs(mu)) > 2e-3                                          # the literal axpby is what is far off


@pytest.mark.skipif(not os.environ.get("NLK_LONG_TESTS"), reason="100 s sparse LU of the 131 654-unknown saddle-point operator: opt-in")
def test_direct_eigenvalue_recomputed(tmp_path):
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "eig.json"
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "cylinder_direct_eig.py"), "--out", str(out)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr[-2000:]
    a = json.load(open(out)); b = json.load(open(os.path.join(GOLDEN, "cylinder_direct_eig.json")))
    assert abs(complex(*a["lambda"][0]) - complex(*b["lambda"][0])) < 1e-9


# ---- second reference fixture: examples/back_fstep/transient_growth (gmsh mesh with rotated elements, 'SYM' planes) ----
@pytest.fixture(scope="module")
def bfs():
    from tests.util import bfs_case
    return bfs_case()


def test_bfs_cfl_gives_the_reference_step_count(bfs):
    """`setup_nek` rule (src/neklab_nek_setup.f90:193-224): dt = 0.5/ctarg, nsteps = ceil(tau/dt).  SURVEY.md C4 quotes
    nsteps = 1933 for tau = 18 -- reproduced from the shipped base flow by the oracle's compute_cfl."""
    om, bf, prm, z = bfs
    ctarg = ops.compute_cfl(om, bf.v, 1.0)
    assert int(np.ceil(18.0 / (0.5 / ctarg))) == 1933


def test_bfs_numbering_partition_and_divergence(bfs):
    om, bf, prm, z = bfs
    assert om.nglob == 69696 and sorted(np.unique(om.mult_count).tolist()) == [1, 2, 4]
    assert same_partition(om.glo, glo_num_from_coords(om.coords, periods={}))        # .ma2 ids are coordinate-consistent
    assert np.bincount(partition_rank(z["pid"], 2)).tolist() == [1380, 1380]
    assert np.bincount(partition_rank(z["pid"], 8)).tolist() == [345] * 8
    # the base flow is stored in single precision: weakly divergence-free to f32 round-off
    div = ops.opdiv(om, bf.v)
    assert np.sqrt((div * div / om.bm2).sum() / om.bm2.sum()) < 1e-6
    assert not om.has_outflow                                                        # inlet and outlet are both 'v'


def test_bfs_symmetry_masks_follow_the_physical_normal(bfs):
    """The gmsh mesh rotates its elements: the free-slip planes y = 1 (x < 0) and y = 20 must mask v_y and leave v_x free
    whatever the element's reference orientation (Nek `bcmask` for 'SYM')."""
    om, bf, prm, z = bfs
    x, y = om.coords[:, 0], om.coords[:, 1]
    top = np.isclose(y, 20.0) & (x > -19.9) & (x < 99.9)
    assert top.sum() > 500 and om.vmask[1][top].max() == 0.0 and om.vmask[0][top].min() == 1.0
    wall = np.isclose(y, 0.0) & (x > 0.1) & (x < 99.9)
    assert om.vmask[0][wall].max() == 0.0 and om.vmask[1][wall].max() == 0.0
    # faces of one physical plane sit on different reference faces
    ref_faces = {f for f in range(4) if (z["cbc"][:, f] == "SYM").any()}
    assert len(ref_faces) > 1
