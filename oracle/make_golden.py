"""ORACLE (test infrastructure): generate committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference fixtures):
    python -m oracle.make_golden cylinder_eig      # ~30-60 min of CPU
    python -m oracle.make_golden small_cases       # seconds

`cylinder_eig` pins the oracle against the reference's ONLY golden number,
|lambda_1| = 1.0156 +- 1e-4 (`/root/reference/test/neklabTests.py:44`).
"""
from __future__ import annotations

import json
import sys
import time

import numpy as np


def cylinder_eig(out="tests/golden/cylinder_eig_oracle.json", kdim=128, nev=2, rst_quirk=True):
    from .cases import cylinder
    from .krylov import eigs, RTOL_DP
    from .stepper import ExptA, NekVec, PertStepper, seeded_field
    NekVec.RST_MODE = 0 if rst_quirk else 1
    mesh, bf, prm, _ = cylinder()
    prm.pressure_solver = "direct"; prm.helm_solver = "direct"
    st = PertStepper(mesh, prm)
    A = ExptA(st, 1.0, bf)
    dt, nsteps = A.init()
    x0 = seeded_field(mesh, 12345, torder=3)
    t0 = time.time()
    hist = []

    def log(niter, k, lam, res):
        i = int(np.argmax(np.abs(lam)))
        hist.append([niter, float(lam[i].real), float(lam[i].imag), float(abs(lam[i])), float(res[i])])
        print(f"iter {niter:4d} k {k:4d} lam1 {lam[i].real:+.10f} {lam[i].imag:+.10f} |lam| {abs(lam[i]):.10f} res {res[i]:.3e}  t={time.time()-t0:.0f}s", flush=True)

    lam, res, X, Y, k = eigs(A.matvec, x0, nev, kdim, RTOL_DP, log=log)
    rec = {"config": "examples/cylinder/stability/direct (Re=50, lx1=6, lxd=9, bdf3, tau=1)",
           "golden_reference": {"modulus": 1.0156, "tol": 1e-4, "source": "test/neklabTests.py:44"},
           "dt": dt, "nsteps": nsteps, "kdim": kdim, "nev": nev, "tol": RTOL_DP,
           "solver": "sparse-direct pressure/Helmholtz (converged limit of cggo/uzawa_gmres)",
           "lam_re": [float(x.real) for x in lam[:10]], "lam_im": [float(x.imag) for x in lam[:10]],
           "modulus": [float(abs(x)) for x in lam[:10]], "resid": [float(x) for x in res[:10]],
           "niter": len(hist), "history": hist, "seed": 12345, "rst_quirk": rst_quirk}
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote", out, "modulus", rec["modulus"][:2])


def cylinder_eig_consistent():
    cylinder_eig(out="tests/golden/cylinder_eig_oracle_consistent.json", rst_quirk=False)


def poiseuille_eig(kdim=100, nev=2, tol=1e-8):
    """Config C1 on the reference's mesh (tests/golden/poiseuille_case.npz): leading eigenpair with consistent restart-field
    arithmetic; writes poiseuille_eig_oracle.json and poiseuille_eigvec.npz (Ritz vector, real/imaginary parts + restart slot)."""
    from tests.util import poiseuille_case
    from .krylov import eigs
    from .stepper import ExptA, NekVec, PertStepper, seeded_field
    om, bf, prm, _ = poiseuille_case()
    NekVec.RST_MODE = 1
    prm.pressure_solver = "direct"; prm.helm_solver = "direct"
    A = ExptA(PertStepper(om, prm), 1.0, bf)
    dt, ns = A.init()
    x0 = seeded_field(om, 12345, torder=2)
    t0 = time.time()

    def log(niter, k, lam, res):
        i = int(np.argmax(np.abs(lam)))
        print(f"iter {niter} lam1 {lam[i].real:+.10f} {lam[i].imag:+.10f} |lam| {abs(lam[i]):.10f} res {res[i]:.2e} t={time.time()-t0:.0f}s", flush=True)

    lam, res, X, Y, k = eigs(A.matvec, x0, nev, kdim, tol, log=log)
    mu = lam[0]; s = np.log(mu) / 1.0
    y = Y[:, 0]
    vr = X[0].copy(); vr.zero(); vi = X[0].copy(); vi.zero()
    for v in (vr, vi):                          # `zero` drops the restart slots: re-open one so that axpby accumulates it
        v._rst_slot(0); v.nrst = 1
    for j in range(len(y)):
        vr.axpby(float(y[j].real), X[j], 1.0); vi.axpby(float(y[j].imag), X[j], 1.0)

    # Golden vectors: the Ritz vector's real and imaginary parts WITHOUT restart slots (the Krylov-Schur restart drops them,
    # `zero` + `axpby`, so no consistent slot exists) and the oracle's exptA images of both (BDF1 -> BDF2 start-up).
    rec = {}
    for tag, v in (("re", vr), ("im", vi)):
        v.nrst = 0; v.rst = [None] * (v.torder - 1)
        out = A.matvec(v)
        rec.update({tag + "_v0": v.v[0], tag + "_v1": v.v[1], tag + "_pr": v.pr,
                    "out_" + tag + "_v0": out.v[0], "out_" + tag + "_v1": out.v[1], "out_" + tag + "_pr": out.pr})
    np.savez_compressed("tests/golden/poiseuille_eigvec.npz", **rec)
    json.dump({"config": "examples/poiseuille/stability/direct_alpha_1 (Re=7500, lx1=8, lxd=12, bdf2, tau=1)", "mu_re": float(mu.real), "mu_im": float(mu.imag),
               "sigma_re": float(s.real), "sigma_im": float(s.imag), "resid": float(res[0]), "dt": dt, "nsteps": ns, "niter": int(k), "kdim": kdim, "tol": tol,
               "rst_mode": 1, "seed": 12345, "solver": "sparse-direct pressure/Helmholtz"}, open("tests/golden/poiseuille_eig_oracle.json", "w"), indent=1)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "cylinder_eig"
    globals()[what]()
