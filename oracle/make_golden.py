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


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "cylinder_eig"
    globals()[what]()
