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
