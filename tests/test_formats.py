"""Readers / writer of the data formats either side of the hot path (`.re2` mesh, `.ma2` vertex map,
`.f%05d` field files -- `load_fld` / `outpost` at `src/neklab_utils.f90:309`, `1cyl.usr:15`).  Round trips run everywhere;
the comparisons with the reference's shipped files run only where `/root/reference` exists (the build container)."""
import os

import numpy as np
import pytest

from neklab_b200.formats import read_fld, read_ma2, read_re2, write_fld
from tests.util import GOLDEN

REF = "/root/reference/examples"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference fixtures are only present in the build container")


@pytest.mark.parametrize("wd,ndim", [(8, 2), (4, 2), (8, 3)])
def test_fld_round_trip(tmp_path, wd, ndim):
    rng = np.random.default_rng(wd + ndim)
    nel, n = 7, 5
    shp = (nel, ndim, n if ndim == 3 else 1, n, n)
    coords, vel = rng.standard_normal(shp), rng.standard_normal(shp)
    pr, temp = rng.standard_normal(shp[:1] + shp[2:]), rng.standard_normal(shp[:1] + shp[2:])
    p = str(tmp_path / "t0.f00001")
    write_fld(p, coords=coords, vel=vel, pr=pr, temp=temp, time=1.25, istep=17, wdsize=wd)
    raw = open(p, "rb").read()
    assert raw[:4] == b"#std" and abs(np.frombuffer(raw, "<f4", 1, 132)[0] - 6.54321) < 1e-6          # 132-byte header + endian tag
    f = read_fld(p)
    assert (f.nx, f.ny, f.nz, f.nel, f.nelg, f.wdsize, f.istep, f.rdcode) == (n, n, shp[2], nel, nel, wd, 17, "XUPT")
    assert f.time == 1.25 and f.ndim == ndim
    tol = 0 if wd == 8 else 1e-6
    for a, b in ((f.coords, coords), (f.vel, vel), (f.pr, pr), (f.temp, temp)):
        assert a.shape == b.shape and np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())
    # a velocity-only file (what `outpost_dnek` writes for eigenvectors after the first one)
    write_fld(p, vel=vel, time=0.0, istep=0)
    g = read_fld(p)
    assert g.rdcode == "U" and g.coords is None and g.pr is None and np.array_equal(g.vel, vel)


def test_fld_reader_sorts_by_global_element_id(tmp_path):
    """Parallel Nek writes elements in rank order; the reader returns them by ascending global id."""
    rng = np.random.default_rng(0)
    nel, n = 5, 4
    vel = rng.standard_normal((nel, 2, 1, n, n))
    p = str(tmp_path / "s0.f00001")
    write_fld(p, vel=vel)
    raw = bytearray(open(p, "rb").read())
    perm = np.array([3, 1, 5, 2, 4], dtype="<i4")                      # file order of global ids
    raw[136:136 + 4 * nel] = perm.tobytes()
    body = np.ascontiguousarray(vel[perm - 1], dtype="<f8").tobytes()
    raw[136 + 4 * nel:] = body
    open(p, "wb").write(bytes(raw))
    assert np.array_equal(read_fld(p).vel, vel)


def test_fld_rejects_other_files(tmp_path):
    p = tmp_path / "x.f00001"; p.write_bytes(b"not a field file" * 20)
    with pytest.raises(ValueError):
        read_fld(str(p))


@needs_ref
def test_reference_cylinder_files_match_the_committed_fixture():
    ex = REF + "/cylinder/stability/direct/"
    f = read_fld(ex + "BF_1cyl0.f00001"); a = read_ma2(ex + "1cyl.ma2"); r = read_re2(ex + "1cyl.re2")
    z = np.load(os.path.join(GOLDEN, "cylinder_case.npz"))
    assert (f.nx, f.nel, f.wdsize, f.rdcode[:3]) == (6, 1996, 8, "XUP") and f.istep == 101 and abs(f.time - 1.0) < 1e-12
    assert np.array_equal(f.coords, z["coords"]) and np.array_equal(f.vel, z["vel"]) and np.array_equal(f.pr, z["pr"])
    assert np.array_equal(a.vertex, z["vertex"]) and np.array_equal(a.pid, z["pid"]) and a.vertex.max() == 2033
    assert np.array_equal(r.cbc[0], z["cbc"]) and r.nel == 1996 and r.ndim == 2
    # re2 corners are the GLL corners of the field file (preprocessor order -> lexicographic)
    from oracle.mesh import lex_corners_from_re2
    c = lex_corners_from_re2(r.xyz, 2)
    gx = f.coords[:, :, 0][:, :, [0, 0, -1, -1], [0, -1, 0, -1]]
    assert np.abs(c - gx).max() < 1e-6


@needs_ref
def test_reference_step_and_convection_cell_files():
    r = read_re2(REF + "/back_fstep/transient_growth/bfs.re2"); a = read_ma2(REF + "/back_fstep/transient_growth/bfs.ma2")
    assert r.nel == a.vertex.shape[0] == 2760 and (r.cbc[0] == "MSH").sum() == 278           # gmsh boundary faces carry ids in bc(5)
    assert sorted(np.unique(np.rint(r.bc[0][r.cbc[0] == "MSH"][:, 4])).astype(int).tolist()) == [2, 3, 4, 5]
    f = read_fld(REF + "/rayBen/baseflow/BF_rayBen0.f00001"); rb = read_re2(REF + "/rayBen/baseflow/rayBen.re2")
    assert (f.nx, f.nel, f.rdcode) == (10, 40, "XUPT") and f.temp is not None
    assert np.abs(f.temp - 2.0 * (1.0 - f.coords[:, 1])).max() < 1e-5                        # shipped state: linear conduction profile, T = 2 (1 - y)
    assert rb.cbc.shape[0] == 2 and set(np.unique(rb.cbc[1]).tolist()) == {"E  ", "P  ", "t  "}


def test_pressure_maps_between_mesh1_and_mesh2(nlk_lib):
    """`map12` / `map21` of the PRODUCT (neklab_b200.api; Nek `mappr`): mesh-2 -> mesh-1 -> mesh-2 is the identity (KAT-2: the fld
    pressure is exactly of degree lx2-1) and both agree with the oracle's maps."""
    from neklab_b200 import api
    from oracle import ops
    from tests.util import box_case, nlk_mesh
    for ndim, nel in ((2, (3, 2)), (3, (2, 2, 2))):
        om, _, _ = box_case(ndim=ndim, nel=nel, n=6, lxd=9)
        m = nlk_mesh(om)
        p2 = np.random.default_rng(ndim).standard_normal(om.bm2.shape)
        assert np.abs(api.map12(m, api.map21(m, p2)) - p2).max() < 1e-13
        assert np.abs(api.map21(m, p2) - ops.map21(om, p2)).max() < 1e-13
        assert np.abs(api.map12(m, om.bm1) - ops.map12(om, om.bm1)).max() < 1e-13


def _parse_like_the_reference(logfile):
    """Transcription of the parsing RULES of the reference harness (test/lib/neklabTestCase.py:413-455): split on blanks, skip
    lines with < 6 tokens or non-numeric columns 2-5, keep the lines whose 6th token is 'T', label them lambda_1, lambda_2, ..."""
    eigs = {}
    for line in open(logfile):
        parts = line.split()
        if len(parts) < 6:
            continue
        try:
            re_, im_, mod, res = (float(parts[i]) for i in (1, 2, 3, 4))
        except ValueError:
            continue
        if parts[5] == "T":
            eigs["lambda_%d" % (len(eigs) + 1)] = dict(Re=re_, Im=im_, modulus=mod, residual=res)
    return eigs


@pytest.mark.gpu
def test_outpost_load_round_trip_and_eigs_output_format(nlk_lib, tmp_path):
    """(f)2: `outpost_dnek` writes a field file `load_fld` (ours and Nek's layout) reads back bit-exactly, pressure mapped mesh 2 ->
    mesh 1 -> mesh 2; `linear_stability_analysis_fixed_point` writes `eigs_output.txt` / `dir_eigenspectrum.npy` that the reference's
    own parser rules accept (test/lib/neklabTestCase.py:413-455; poiseuille.usr:30-36 for the npy layout)."""
    from neklab_b200 import api
    from oracle.stepper import seeded_field
    from tests.util import box_case, nlk_mesh
    om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "})
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=0.05, torder=3, vtol=1e-11, ptol=1e-10, gmres_maxit=500))
    x0 = seeded_field(om, 3); x0.pr = np.random.default_rng(1).standard_normal(om.bm2.shape)
    v = ctx.vec(); v.upload(x0.v, x0.pr)
    path = api.outpost_dnek(v, "BF_", "box", 1, str(tmp_path), time=1.0, istep=101)
    assert os.path.basename(path) == "BF_box0.f00001"
    f = read_fld(path)
    assert f.rdcode == "XUP" and f.istep == 101 and np.array_equal(f.coords, om.coords) and np.array_equal(f.vel[:, 0], x0.v[0])
    w = api.load_fld(ctx, path)
    wv, wp, _ = w.download()
    assert np.array_equal(wv[1], x0.v[1]) and np.abs(wp - x0.pr).max() < 1e-12
    # analysis driver output files
    x = om.coords
    bf = ctx.vec(); bf.upload([1.0 + 0.3 * np.sin(0.5 * x[:, 1]), 0.2 * np.cos(0.4 * x[:, 0])])
    A = api.exptA_linop(ctx, 0.1, bf); A.init()
    r = api.linear_stability_analysis_fixed_point(A, 12, 2, outdir=str(tmp_path), tol=1e-3, case="box")
    got = _parse_like_the_reference(str(tmp_path / "eigs_output.txt"))
    assert "lambda_1" in got and abs(got["lambda_1"]["modulus"] - abs(r["lam"][0])) < 1e-9 * abs(r["lam"][0])
    spec = np.load(str(tmp_path / "dir_eigenspectrum.npy"))
    assert spec.shape == (2, 3) and np.allclose(spec[:, 0] + 1j * spec[:, 1], np.log(r["lam"]) / 0.1)
    assert os.path.exists(str(tmp_path / "dirbox0.f00001")) and read_fld(str(tmp_path / "dirbox0.f00001")).rdcode == "XUP"
    ctx.close()
