"""Config C1 (examples/poiseuille, Re = 7500, alpha = 1, bdf2, tau = 1) on the reference's own mesh: the oracle's leading
eigenvalue against an independent Orr-Sommerfeld solve, and a one-apply check of the recorded Ritz pair.

tests/golden/poiseuille_eig_oracle.json and poiseuille_eigvec.npz were produced by `python -m oracle.make_golden
poiseuille_eig` (Krylov-Schur, kdim 100, tol 1e-8, consistent restart-field arithmetic, sparse-direct inner solves;
about 11 min of CPU)."""
import json
import os

import numpy as np
import pytest

from oracle.stepper import ExptA, NekVec, PertStepper
from tests.util import GOLDEN, orr_sommerfeld_leading, poiseuille_case


def load_pair(om):
    """Golden vectors: real / imaginary part of the recorded Ritz vector (no restart slots) and the oracle's exptA images."""
    z = np.load(os.path.join(GOLDEN, "poiseuille_eigvec.npz"))
    ins, outs = [], []
    for part in ("re", "im"):
        v = NekVec(om, 2)
        v.v = [z[part + "_v0"].copy(), z[part + "_v1"].copy()]; v.pr = z[part + "_pr"].copy()
        ins.append(v)
        outs.append([z["out_" + part + "_v0"], z["out_" + part + "_v1"]])
    return ins, outs


def rayleigh_quotient(om, vr, vi, yr, yi):
    """<v, A v> / <v, v> in the bm1 inner product for v = vr + i vi, A v = yr + i yi."""
    ip = lambda a, b: sum(float((a[c] * b[c] * om.bm1).sum()) for c in range(2))
    return ((ip(vr, yr) + ip(vi, yi)) + 1j * (ip(vr, yi) - ip(vi, yr))) / (ip(vr, vr) + ip(vi, vi))


def test_recorded_eigenvalue_matches_orr_sommerfeld():
    rec = json.load(open(os.path.join(GOLDEN, "poiseuille_eig_oracle.json")))
    lam = orr_sommerfeld_leading(7500.0)
    assert 0.0015 < lam.real < 0.003                                   # the unstable Tollmien-Schlichting wave (Re_c = 5772)
    assert abs(rec["sigma_re"] - lam.real) < 1e-4, (rec["sigma_re"], lam.real)          # growth rate: 6.0e-5 apart
    assert abs(abs(rec["sigma_im"]) - abs(lam.imag)) < 1e-3                             # frequency: 7.6e-4 apart (dt = 0.02, bdf2)
    assert rec["resid"] < 1e-8 and rec["nsteps"] == 50 and abs(rec["dt"] - 0.02) < 1e-15       # the CFL rule lands on poiseuille.par's dt = 2e-2


def test_one_apply_reproduces_the_golden_vectors_and_the_growth():
    """A live run of the oracle's time stepper on the reference mesh (2 x 50 steps): the stored images are reproduced, and the
    Rayleigh quotient of the recorded Ritz vector is mu up to the BDF1 start-up error of a restart-less apply (1.4e-3).
    (The Krylov-Schur restart rebuilds its basis with `zero` + `axpby`, which drops the restart slots, so a Ritz vector has no
    consistent slot of its own and A v = mu v cannot be checked more tightly than that with a single apply.)"""
    om, bf, prm, _ = poiseuille_case()
    rec = json.load(open(os.path.join(GOLDEN, "poiseuille_eig_oracle.json")))
    mu = complex(rec["mu_re"], rec["mu_im"])
    prm.pressure_solver = "direct"; prm.helm_solver = "direct"
    A = ExptA(PertStepper(om, prm), 1.0, bf)
    dt, ns = A.init()
    assert ns == rec["nsteps"]
    (vr, vi), (gr, gi) = load_pair(om)
    yr, yi = A.matvec(vr), A.matvec(vi)
    nrm = lambda f: np.sqrt(sum(float((a * a * om.bm1).sum()) for a in f))
    assert nrm([yr.v[c] - gr[c] for c in range(2)]) < 1e-10 * nrm(gr) and nrm([yi.v[c] - gi[c] for c in range(2)]) < 1e-10 * nrm(gi)
    assert abs(rayleigh_quotient(om, vr.v, vi.v, yr.v, yi.v) - mu) < 2e-3
