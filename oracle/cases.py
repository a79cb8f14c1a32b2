"""ORACLE (test infrastructure): builders for the reference's example configs from shipped fixtures.

Reads `/root/reference/examples/...` (only available in the build container, never on the GPU
box); used by oracle/make_golden.py to generate committed fixtures under tests/golden/.
"""
from __future__ import annotations

import os

import numpy as np

from neklab_b200.formats import read_fld, read_ma2, read_re2
from . import ops
from .mesh import SEMesh
from .stepper import NekVec, StepParams

REF = os.environ.get("NEKLAB_REFERENCE", "/root/reference")


def cylinder(adjoint_bcs: bool = False):
    """examples/cylinder/stability/direct: Re=50, lx1=6, lxd=9, bdf3 (1cyl.par, SIZE)."""
    ex = os.path.join(REF, "examples/cylinder/stability/direct/")
    f = read_fld(ex + "BF_1cyl0.f00001"); a = read_ma2(ex + "1cyl.ma2"); r = read_re2(ex + "1cyl.re2")
    cbc = r.cbc[0].copy()
    if adjoint_bcs:                                  # 1cyl.usr usrdat2: 'O' -> 'v' for the adjoint
        cbc[cbc == "O  "] = "v  "
    mesh = SEMesh(f.coords, a.vertex, cbc, 9)
    bf = NekVec(mesh, 3)
    bf.v = [f.vel[:, 0].copy(), f.vel[:, 1].copy()]
    bf.pr = ops.map12(mesh, f.pr)
    prm = StepParams(viscosity=1.0 / 50.0, torder=3, vtol=1e-9, ptol=1e-7)
    return mesh, bf, prm, a


BFS_BOUNDARY_IDS = {5: "W  ", 2: "v  ", 3: "v  ", 4: "SYM"}     # bfs.usr usrdat: setbc(5,1,'W  ') setbc(2,1,'v  ') setbc(3,1,'v  ') setbc(4,1,'SYM')


def bfs_cbc(re2):
    """gmsh-converted .re2: every boundary face carries 'MSH' and its physical-group id in bc(5); `usrdat` maps ids to codes."""
    cbc = re2.cbc[0].copy()
    ids = np.rint(re2.bc[0][:, :, 4]).astype(int)
    for bid, code in BFS_BOUNDARY_IDS.items():
        cbc[(re2.cbc[0] == "MSH") & (ids == bid)] = code
    assert not (cbc == "MSH").any()
    return cbc


def back_fstep():
    """examples/back_fstep/transient_growth: Re=600, lx1=6, lxd=9, bdf2, tau=18, explicit filter 0.01/0.84 (bfs.par, SIZE, bfs.usr)."""
    ex = os.path.join(REF, "examples/back_fstep/transient_growth/")
    f = read_fld(ex + "BF_bfs0.f00001"); a = read_ma2(ex + "bfs.ma2"); r = read_re2(ex + "bfs.re2")
    cbc = bfs_cbc(r)
    mesh = SEMesh(f.coords, a.vertex, cbc, 9)
    bf = NekVec(mesh, 2)
    bf.v = [f.vel[:, 0].copy(), f.vel[:, 1].copy()]
    bf.pr = ops.map12(mesh, f.pr)
    prm = StepParams(viscosity=1.0 / 600.0, torder=2, vtol=1e-8, ptol=1e-6, filter_weight=0.01, filter_cutoff=0.84)
    return mesh, bf, prm, a
