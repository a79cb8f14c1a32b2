"""Cylinder Re=50 leading global mode (reference: examples/cylinder/stability/direct/1cyl.usr:13-26).

    python examples/cylinder_eigs.py [--kdim 128] [--nev 2] [--out DIR]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neklab_b200 import api, build  # noqa: E402


def load_case():
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cylinder_case.npz"))
    return z


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kdim", type=int, default=128)
    ap.add_argument("--nev", type=int, default=2)
    ap.add_argument("--tau", type=float, default=1.0)
    ap.add_argument("--vtol", type=float, default=1e-9)
    ap.add_argument("--ptol", type=float, default=1e-7)
    ap.add_argument("--cfl", type=float, default=0.5)
    ap.add_argument("--maxit", type=int, default=100)
    ap.add_argument("--rst-mode", type=int, default=0)
    ap.add_argument("--proj", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    build.build()
    z = load_case()
    mesh = api.Mesh(z["coords"], z["vertex"], z["cbc"], 9)
    prm = api.default_params(viscosity=1.0 / 50.0, torder=3, vtol=a.vtol, ptol=a.ptol, gmres_maxit=a.maxit, cfl_limit=a.cfl, rst_mode=a.rst_mode, pr_proj=a.proj)
    t0 = time.time()
    ctx = api.Context(mesh, prm)
    print("setup %.2fs" % (time.time() - t0), flush=True)
    # pressure on mesh 2: evaluate the mesh-1 polynomial at the GL points (mappr inverse; exact, KAT-2)
    I12 = mesh.basis("I12").reshape(mesh.lx1 - 2, mesh.lx1)
    p2 = np.einsum("qj,pi,ezji->ezqp", I12, I12, z["pr"])
    bf = ctx.vec(); bf.upload([z["vel"][:, 0], z["vel"][:, 1]], p2)
    A = api.exptA_linop(ctx, a.tau, bf)
    print("init", A.init(), flush=True)
    t0 = time.time()

    def cb(it, k, lam, res):
        i = int(np.argmax(np.abs(lam)))
        print(f"iter {it:4d} k {k:4d} lam1 {lam[i].real:+.10f} {lam[i].imag:+.10f} |lam| {abs(lam[i]):.10f} res {res[i]:.3e} t={time.time()-t0:.1f}s", flush=True)
    r = api.linear_stability_analysis_fixed_point(A, a.kdim, a.nev, outdir=a.out)
    for h in r["history"][-1:]:
        cb(h[0], h[1], h[2] + 1j * h[3], h[4])
    print("time-to-eigs %.2fs niter %d info %d" % (time.time() - t0, r["niter"], r["info"]))
    print("lam", r["lam"], "modulus", np.abs(r["lam"]), "resid", r["resid"])
    print("sigma+i omega", r["eigvals"])
    print("last matvec stats", A.stats())
    if a.out:
        json.dump(dict(modulus=[float(x) for x in np.abs(r["lam"])], lam_re=[float(x.real) for x in r["lam"]], lam_im=[float(x.imag) for x in r["lam"]],
                       resid=[float(x) for x in r["resid"]], niter=r["niter"], seconds=time.time() - t0, stats=A.stats()), open(os.path.join(a.out, "eigs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
