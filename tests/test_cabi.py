"""CPU-side checks of the C-ABI library: loads, exports every symbol of include/nlk.h, host mesh setup parity with
the oracle (bit-exact numbering), LAPACK-free dense eig, partition rule, and the no-CPU-fallback contract."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.mesh import partition_rank
from tests.util import box_case, cylinder_case, nlk_mesh, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(nlk_lib):
    hdr = open(os.path.join(ROOT, "include", "nlk.h")).read()
    declared = set(re.findall(r"\b(nlk_[a-zA-Z0-9_]+)\s*\(", hdr))
    declared -= {"nlk_eigs_cb"}
    from neklab_b200 import api
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    for s in declared:
        assert hasattr(nlk_lib, s), s


@pytest.mark.parametrize("cfg", [dict(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "}),
                                 dict(ndim=2, nel=(5, 4), n=8, lxd=12, periodic=[True, False]),
                                 dict(ndim=3, nel=(3, 3, 3), n=5, lxd=8, periodic=[False, False, True]),
                                 dict(ndim=3, nel=(2, 2, 2), n=4, lxd=6, bc={"zhi": "SYM"})])
def test_host_mesh_matches_oracle(nlk_lib, cfg):
    om, _, _ = box_case(**cfg)
    m = nlk_mesh(om)
    assert np.array_equal(m.glo_num(), om.glo)
    for nm, ref in [("bm1", om.bm1), ("jac", om.jac), ("binvm1", om.binvm1), ("vmult", om.vmult), ("bm2", om.bm2), ("g11", om.G[0][0]), ("g12", om.G[0][1])]:
        assert rel(m.field(nm), ref) < 1e-12, nm
    for c in range(om.ndim):
        assert np.array_equal(m.field(f"vmask{c}"), om.vmask[c])
    assert m.info.nglob_local == om.nglob
    assert bool(m.info.has_outflow) == om.has_outflow
    for nm in ("z1", "w1", "z2", "w2", "D", "I12", "D12", "I1d", "Dd"):
        ref = getattr(om.b, nm)
        assert np.abs(m.basis(nm) - np.asarray(ref).ravel()).max() < 1e-13, nm


def test_cylinder_mesh_numbering_bit_exact(nlk_lib):
    om, bf, prm, z = cylinder_case()
    m = nlk_mesh(om)
    assert np.array_equal(m.glo_num(), om.glo)
    assert m.info.nvert == 2033 and m.info.nglob_local == 50089


def test_partition_rule(nlk_lib):
    from neklab_b200 import api
    z = cylinder_case()[3]
    for P in (1, 2, 4, 8):
        assert np.array_equal(api.partition(z["pid"], P), partition_rank(z["pid"], P))
    with pytest.raises(api.NlkError):
        api.partition(z["pid"], 3)


def test_dense_eig(nlk_lib):
    from neklab_b200 import api
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 60):
        A = rng.standard_normal((n, n))
        wr, wi, VR = api.dense_eig(A)
        lam = wr + 1j * wi
        ref = np.linalg.eigvals(A)
        assert max(np.abs(ref - l).min() for l in lam) < 1e-10
        j = 0
        while j < n:
            if wi[j] == 0:
                v = VR[:, j]; assert np.abs(A @ v - wr[j] * v).max() < 1e-9; j += 1
            else:
                v = VR[:, j] + 1j * VR[:, j + 1]; assert np.abs(A @ v - lam[j] * v).max() < 1e-9; j += 2


def test_invalid_mesh_is_rejected(nlk_lib):
    from neklab_b200 import api
    om, bm, coords = box_case(ndim=2, nel=(2, 2), n=5, lxd=8)
    bad = coords.copy(); bad[:, 0] *= -1.0                    # negative Jacobian
    with pytest.raises(api.NlkError):
        api.Mesh(bad, om.vertex, om.cbc_v, 8)


def test_no_cpu_fallback(nlk_lib):
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    from neklab_b200 import api
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    om, _, _ = box_case(ndim=2, nel=(2, 2), n=5, lxd=8)
    m = nlk_mesh(om)
    with pytest.raises(api.NlkError, match="no usable CUDA device"):
        api.Context(m, api.default_params())
