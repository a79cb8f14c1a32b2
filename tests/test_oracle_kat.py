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
