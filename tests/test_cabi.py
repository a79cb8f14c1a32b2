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


def test_bfs_mesh_numbering_and_symmetry_masks(nlk_lib):
    """Second reference mesh (examples/back_fstep, gmsh, rotated elements): numbering bit-exact, masks equal to the oracle's;
    a plain 'SYM' on a face whose physical normal is not the element's reference axis is rejected instead of mis-masked."""
    from neklab_b200 import api
    from tests.util import bfs_case
    om, bf, prm, z = bfs_case()
    m = nlk_mesh(om)                                    # api.Mesh resolves 'SYM' -> 'SYx'/'SYy' from the coordinates
    assert np.array_equal(m.glo_num(), om.glo)
    assert m.info.nglob_local == om.nglob == 69696 and not m.info.has_outflow
    for c in range(2):
        assert np.array_equal(m.field(f"vmask{c}"), om.vmask[c])
    res = api.resolve_sym(om.coords, z["cbc"])
    assert set(np.unique(res[z["cbc"] == "SYM"]).tolist()) == {"SYy"}
    # multi-rank entry (local coordinates, no auto-resolution): the resolved codes give the same masks and numbering
    gll = api.partition(z["pid"], 2); sel = np.where(gll == 0)[0]
    m0 = api.Mesh(om.coords[sel], om.vertex, res, 9, gllnid=gll, rank=0, nranks=2)
    assert np.array_equal(m0.field("vmask1"), om.vmask[1][sel]) and np.array_equal(m0.glo_num(), om.glo[sel])


def test_symmetry_code_on_a_rotated_element(nlk_lib):
    """One element turned by 90 degrees (x = 1 - s, y = r): its reference face 0 (s = -1) is the physical plane x = 1.  A plain
    'SYM' there would mask the wrong component, so the library refuses it when it cannot resolve the axis itself, and the
    resolved code 'SYx' masks v_x."""
    from neklab_b200 import api
    from neklab_b200.boxmesh import gll_points
    n = 5
    g = 0.5 * (np.asarray(gll_points(n)) + 1.0)
    coords = np.zeros((1, 2, 1, n, n))
    coords[0, 0, 0] = 1.0 - g[:, None]              # x[j, i] = 1 - s_j
    coords[0, 1, 0] = g[None, :]                    # y[j, i] = r_i
    vertex = np.array([[1, 2, 3, 4]], dtype=np.int64)
    cbc = np.array([["SYM", "W  ", "W  ", "W  "]])
    d = api.MeshDesc(2, n, 8, 1, 1, api._p(np.ascontiguousarray(coords[:, 0])), api._p(np.ascontiguousarray(coords[:, 1])), None,
                     api._p(vertex), api._p(api.cbc_bytes(cbc)), None, None, 0, 1)
    h = api.C.c_void_p()
    assert api.lib().nlk_mesh_create(api.C.byref(d), api.C.byref(h)) != 0          # raw C-ABI call with the ambiguous code
    assert b"symmetry" in api.lib().nlk_last_error()
    res = api.resolve_sym(coords, cbc)
    assert res[0, 0] == "SYx"
    m = api.Mesh(coords, vertex, cbc, 8)                                             # the Python front end resolves it
    mx, my = m.field("vmask0"), m.field("vmask1")
    face = np.isclose(coords[:, 0], 1.0)
    inner = face & (coords[:, 1] > 1e-9) & (coords[:, 1] < 1 - 1e-9)                 # away from the wall corners
    assert mx[inner].max() == 0.0 and my[inner].min() == 1.0
    from oracle.mesh import SEMesh
    om = SEMesh(coords, vertex, cbc, 8)
    assert np.array_equal(om.vmask[0], mx) and np.array_equal(om.vmask[1], my)


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


def test_fortran_shim_binds_only_exported_symbols(nlk_lib):
    """fortran/neklab_b200.f90 cannot be compiled here (no Fortran compiler in the image): check mechanically that every
    bind(C, name=...) it declares is exported by libnlk.so and declared in include/nlk.h, and that the interoperable
    derived types list the fields of the C structs in order."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "fortran", "neklab_b200.f90")).read()
    hdr = open(os.path.join(root, "include", "nlk.h")).read()
    names = sorted(set(re.findall(r'bind\(C,\s*name="(\w+)"\)', src)))
    assert len(names) >= 50
    for nm in names:
        assert hasattr(nlk_lib, nm), nm
        assert re.search(r"\b%s\s*\(" % nm, hdr), nm
    # nlk_params field order (Fortran sequence association with the C struct)
    cblk = hdr[hdr.index("typedef struct {\n  double viscosity"):hdr.index("} nlk_params;")]
    cblk = re.sub(r"/\*.*?\*/", "", cblk, flags=re.S)
    c_fields = []
    for decl in re.findall(r"\b(?:double|int32_t)\s+([^;]+);", cblk):
        c_fields += [w.strip().split("[")[0] for w in decl.split(",")]
    fblk = src[src.index("type, bind(C) :: nlk_params"):]
    fblk = fblk[:fblk.index("end type")]
    f_fields = []
    for line in fblk.splitlines()[1:]:
        if "::" in line:
            f_fields += [w.strip().split("(")[0] for w in line.split("::")[1].split(",")]
    assert f_fields == c_fields, (f_fields, c_fields)
