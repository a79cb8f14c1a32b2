"""Sweep of the start-up / pressure-bookkeeping degrees of freedom of the perturbation step against the reference's
only golden value (test/neklabTests.py:42-46: cylinder Re=50, |lambda_1| = 1.0156 +- 1e-4).  See DESIGN.md 1.1.

    python examples/golden_sweep.py --out gpurun_out/golden_sweep.json [--only rst,variant,torder ...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neklab_b200 import api, build  # noqa: E402

GOLD = 1.0156

# (rst_mode, step_variant, torder, cfl, seed, label)
CASES = [
    (1, 0, 3, 0.5, 12345, "consistent rst"),
    (2, 0, 3, 0.5, 12345, "no rst (BDF1->3 ramp every matvec)"),
    (0, 0, 3, 0.5, 12345, "literal rst (real_vectors.f90:186-200)"),
    (0, 0, 3, 0.5, 777, "literal rst, other start vector"),
    (2, 4, 3, 0.5, 12345, "no rst, input pressure ignored"),
    (1, 4, 3, 0.5, 12345, "consistent rst, input pressure ignored"),
    (2, 3, 3, 0.5, 12345, "no rst, prlagp never updated + dp added to prp"),
    (1, 3, 3, 0.5, 12345, "consistent rst, prlagp never updated + dp added to prp"),
    (0, 3, 3, 0.5, 12345, "literal rst, prlagp never updated + dp added to prp"),
    (2, 8, 3, 0.5, 12345, "no rst, first-order pressure extrapolation"),
    (1, 8, 3, 0.5, 12345, "consistent rst, first-order pressure extrapolation"),
    (2, 12, 3, 0.5, 12345, "no rst, input pressure ignored, first-order pressure extrapolation"),
    (2, 7, 3, 0.5, 12345, "no rst, input pressure ignored, prlagp never updated + dp added to prp"),
    (2, 0, 2, 0.5, 12345, "bdf2, no rst"),
    (1, 0, 2, 0.5, 12345, "bdf2, consistent rst"),
    (1, 0, 3, 0.25, 12345, "consistent rst, dt/2"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden_sweep.json")
    ap.add_argument("--cases", default=None, help="comma-separated case indices")
    a = ap.parse_args()
    build.build()
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cylinder_case.npz"))
    mesh = api.Mesh(z["coords"], z["vertex"], z["cbc"], 9)
    I12 = mesh.basis("I12").reshape(mesh.lx1 - 2, mesh.lx1)
    p2 = np.einsum("qj,pi,ezji->ezqp", I12, I12, z["pr"])
    sel = range(len(CASES)) if a.cases is None else [int(s) for s in a.cases.split(",")]
    out = []
    for i in sel:
        rst, var, tord, cfl, seed, label = CASES[i]
        prm = api.default_params(viscosity=1.0 / 50.0, torder=tord, vtol=1e-9, ptol=1e-7, cfl_limit=cfl, rst_mode=rst, pr_proj=20, step_variant=var)
        ctx = api.Context(mesh, prm)
        bf = ctx.vec(); bf.upload([z["vel"][:, 0], z["vel"][:, 1]], p2)
        A = api.exptA_linop(ctx, 1.0, bf); A.init()
        x0 = ctx.vec(); x0.rand(True, seed)
        t0 = time.time()
        r = api.eigs(A, 2, 128, x0=x0)
        mod = float(np.abs(r["lam"][0]))
        rec = dict(case=i, rst_mode=rst, step_variant=var, torder=tord, cfl=cfl, seed=seed, label=label, modulus=mod, lam_re=float(r["lam"][0].real),
                   lam_im=float(r["lam"][0].imag), resid=float(r["resid"][0]), niter=r["niter"], info=r["info"], seconds=time.time() - t0,
                   dist_to_golden=mod - GOLD, in_band=bool(abs(mod - GOLD) < 1e-4), nsteps=A.stats()["nsteps"])
        print(json.dumps(rec), flush=True)
        out.append(rec)
        ctx.close()
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
