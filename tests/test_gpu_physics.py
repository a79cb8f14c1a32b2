"""External physics anchors (not from the reference's tests; SURVEY.md §8c last row): the whole GPU path -- dealiased
linearised convection, PN-PN-2 pressure solve, BDF/EXT stepping with the rst protocol, device-resident Krylov-Schur --
must reproduce (1) the Orr-Sommerfeld eigenvalue of plane Poiseuille flow (config `examples/poiseuille`: Re = 7500,
alpha = 1, bdf2, tau = 1) computed here by an independent Chebyshev collocation solve, and (2) the sign change of the
Rayleigh-Benard growth rate around Ra_c = 1707.762 (`examples/rayBen/baseflow/rayBen.par:7-9`)."""
import numpy as np
import pytest

from neklab_b200.boxmesh import box_mesh
from oracle.mesh import coords_from_corners

pytestmark = pytest.mark.gpu


from tests.util import orr_sommerfeld_leading  # noqa: E402  (independent Chebyshev collocation solve)

import os


@pytest.mark.parametrize("rst_mode,tol_growth", [(1, 3e-4)] + ([(0, 2e-3)] if os.environ.get("NLK_LONG_TESTS") else []))
def test_poiseuille_orr_sommerfeld(nlk_lib, rst_mode, tol_growth):
    """rst_mode=1 (consistent restart-field arithmetic) must hit the Orr-Sommerfeld growth rate; rst_mode=0 (the
    reference's nek_daxpby as written, real_vectors.f90:186-200) is allowed its documented O(dt) bias."""
    from neklab_b200 import api
    Re, tau, n = 7500.0, 1.0, 10
    lam_os = orr_sommerfeld_leading(Re)
    assert 0.0015 < lam_os.real < 0.003                      # unstable TS wave (Re_c = 5772)
    # wall-refined mesh: 8 x 10 elements on [0, 2 pi] x [-1, 1]
    ny = 10
    yy = -np.cos(np.pi * np.arange(ny + 1) / ny)

    def warp(c):
        c = c.copy()
        c[:, 1] = np.interp(c[:, 1], np.linspace(-1, 1, ny + 1), yy)
        return c
    bm = box_mesh((8, ny), (0.0, -1.0), (2 * np.pi, 1.0), periodic=[True, False], warp=warp)
    coords = coords_from_corners(bm["corners"], n)
    mesh = api.Mesh(coords, bm["vertex"], bm["cbc"], 15)
    ctx = api.Context(mesh, api.default_params(viscosity=1.0 / Re, torder=2, vtol=1e-11, ptol=1e-10, gmres_maxit=400, pr_proj=20, rst_mode=rst_mode))
    bf = ctx.vec(); bf.upload([1.0 - coords[:, 1] ** 2, np.zeros_like(coords[:, 1])])
    A = api.exptA_linop(ctx, tau, bf); A.init()
    r = api.linear_stability_analysis_fixed_point(A, kdim=100, nev=2, tol=1e-7)
    lam = r["eigvals"][0]                                     # log(mu)/tau (src/neklab_analysis.f90:84)
    ctx.close()
    assert r["info"] == 0
    assert abs(lam.real - lam_os.real) < tol_growth, (lam, lam_os)
    assert abs(abs(lam.imag) - abs(lam_os.imag)) < 2e-3, (lam, lam_os)


@pytest.mark.parametrize("Ra,sign", [(1900.0, +1), (1550.0, -1)])
def test_rayleigh_benard_onset(nlk_lib, Ra, sign):
    """rayBen config (Pr = 1, periodic box of length 2.0158 x 2 cells wide, conduction state T = 1 - y):
    growth rate positive at Ra = 1900 > Ra_c = 1707.762 and negative at Ra = 1550."""
    from neklab_b200 import api
    n, Lx = 8, 2.0158 * 2
    bm = box_mesh((6, 4), (0.0, 0.0), (Lx, 1.0), periodic=[True, False])
    coords = coords_from_corners(bm["corners"], n)
    cbc_t = bm["cbc"].copy(); cbc_t[cbc_t == "W  "] = "t  "
    mesh = api.Mesh(coords, bm["vertex"], bm["cbc"], 12, cbc_t=cbc_t)
    # non-dimensionalisation of rayBen.usr/.par: viscosity = Pr = 1, conductivity = 1, buoyancy = Ra*Pr * T'
    ctx = api.Context(mesh, api.default_params(viscosity=1.0, torder=3, vtol=1e-11, ptol=1e-10, ttol=1e-11, ifheat=1, conductivity=1.0,
                                               rhocp=1.0, buoyancy=(0.0, Ra, 0.0), gmres_maxit=400, pr_proj=20))
    ctx.set_dt(2.0e-3)                                        # zero base flow: dt preset (setup_nek disables recompute_dt)
    bf = ctx.vec(); bf.upload([np.zeros_like(coords[:, 0]), np.zeros_like(coords[:, 0])], None, 1.0 - coords[:, 1])
    A = api.exptA_linop(ctx, 0.1, bf); A.init()
    r = api.linear_stability_analysis_fixed_point(A, kdim=60, nev=1, tol=1e-7)
    lam = r["eigvals"][0]
    ctx.close()
    assert r["info"] == 0
    assert abs(lam.imag) < 1e-6                               # exchange of stabilities: real leading eigenvalue
    assert np.sign(lam.real) == sign, lam


def test_poiseuille_reference_mesh_golden_apply(nlk_lib):
    """Config C1 on the reference's own mesh (tests/golden/poiseuille_case.npz, x-periodic, wall-graded, Re = 7500, bdf2):
    one GPU apply of the recorded Ritz vector's real and imaginary parts must reproduce the oracle's stored images
    (tests/golden/poiseuille_eigvec.npz) and, through them, the growth of the Tollmien-Schlichting wave that
    tests/test_oracle_poiseuille.py ties to Orr-Sommerfeld."""
    import json
    from neklab_b200 import api
    from tests.test_oracle_poiseuille import load_pair, rayleigh_quotient
    from tests.util import GOLDEN, nlk_mesh, poiseuille_case
    om, bf, prm, _ = poiseuille_case()
    rec = json.load(open(os.path.join(GOLDEN, "poiseuille_eig_oracle.json")))
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=1.0 / 7500.0, torder=2, vtol=1e-12, ptol=1e-12, gmres_maxit=3000, cg_maxit=3000, pr_proj=20))
    bd = ctx.vec(); bd.upload(bf.v, bf.pr)
    A = api.exptA_linop(ctx, 1.0, bd)
    assert A.init()["nsteps"] == rec["nsteps"] == 50
    ins, gold = load_pair(om)
    outs = []
    for v in ins:
        d = ctx.vec(); d.upload(v.v, v.pr)
        outs.append(A.matvec(d).download()[0])
    ctx.close()
    nrm = lambda f: np.sqrt(sum(float((a * a * om.bm1).sum()) for a in f))
    for o, g in zip(outs, gold):
        assert nrm([o[c] - g[c] for c in range(2)]) < 1e-9 * nrm(g)            # iterative (1e-12) vs sparse-direct inner solves
    mu = complex(rec["mu_re"], rec["mu_im"])
    assert abs(rayleigh_quotient(om, ins[0].v, ins[1].v, outs[0], outs[1]) - mu) < 2e-3
