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
