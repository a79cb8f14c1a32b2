"""ORACLE (test infrastructure, CPU only; never imported by the product): the leading eigenvalue of the SEMI-DISCRETE linearised operator of the
cylinder case (examples/cylinder/stability/direct: Re = 50, lx1 = 6, lxd = 9, the shipped base flow) by a direct sparse
shift-invert eigen-solve -- no time stepper involved.

    lambda B q = -(rho C(U) + nu A) q + D^T p ,   D q = 0        (same weak forms as the stepper: oracle/ops.py)

exp(lambda * tau) is what exptA's leading Ritz value converges to as dt -> 0, so this pins the dt-converged value of the
restatement (1.01573, tests/golden/cylinder_golden_sweep_r02.json) independently of BDF/EXT, the rst protocol, the pressure
solver and the Krylov-Schur driver, and says where the reference's golden 1.0156 +- 1e-4 (test/neklabTests.py:42-46) sits
relative to it.

    python -m oracle.cylinder_direct_eig [--out tests/golden/cylinder_direct_eig.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ops  # noqa: E402
from oracle.stepper import _local_deriv_blocks  # noqa: E402
from tests.util import cylinder_case  # noqa: E402


def conv_collocated(om, u, C):
    """un-dealiased alternative (Nek `conv1`, param(99) = 0): bm1 * (C . grad u) evaluated at the GLL points"""
    g = ops.local_grad(om, u)                               # d u / d r_k
    d = om.ndim
    out = 0.0
    for c in range(d):
        dudx = sum(om.rx[k][c] * g[k] for k in range(d)) / om.jac
        out = out + C[c] * dudx
    return om.bm1 * out


def assemble(om, U, nu, rho=1.0, conv=None):
    conv = conv or ops.convect_new
    d, E = om.ndim, om.E
    nn, nq = om.n ** d, om.q ** d
    shape = om.bm1.shape
    Q = sp.csr_matrix((np.ones(E * nn), (np.arange(E * nn), om.gidx)), shape=(E * nn, om.nglob))
    mg = []
    for c in range(d):
        m = np.zeros(om.nglob); m[om.gidx] = om.vmask[c].ravel(); mg.append(m)
    free = [np.where(m > 0)[0] for m in mg]
    # element blocks of rho C(U) + nu A : blk[c][cj][e, a, b] = d r_c[a] / d u'_cj[b]
    blk = [[np.zeros((E, nn, nn)) for _ in range(d)] for _ in range(d)]
    zero = np.zeros(shape)
    for b in range(nn):
        eb = np.zeros((E, nn)); eb[:, b] = 1.0; eb = eb.reshape(shape)
        for cj in range(d):
            up = [zero] * d; up = list(up); up[cj] = eb
            for c in range(d):
                r = rho * conv(om, U[c], up)                                       # u'.grad U_c
                if c == cj:
                    r = r + rho * conv(om, eb, U) + ops.axhelm(om, eb, nu, 0.0)
                blk[c][cj][:, :, b] = r.reshape(E, nn)
    bsr = lambda B_: sp.bsr_matrix((B_, np.arange(E), np.arange(E + 1)), shape=(E * nn, E * nn)).tocsr()
    K = [[(Q.T @ bsr(blk[c][cj]) @ Q).tocsr()[free[c]][:, free[cj]] for cj in range(d)] for c in range(d)]
    Bg = np.asarray(Q.T @ om.bm1.ravel()).ravel()
    D = []
    for c in range(d):
        Dc = sp.bsr_matrix((_local_deriv_blocks(om, c), np.arange(E), np.arange(E + 1)), shape=(E * nq, E * nn)).tocsr()
        D.append((Dc @ Q).tocsr()[:, free[c]])
    return K, Bg, D, free


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--nev", type=int, default=4)
    ap.add_argument("--conv", default="dealiased", choices=["dealiased", "collocated"], help="sensitivity study: un-dealiased convective terms")
    a = ap.parse_args()
    t0 = time.time()
    om, bf, prm, _ = cylinder_case()
    K, Bg, D, free = assemble(om, bf.v, prm.viscosity, conv=conv_collocated if a.conv == "collocated" else None)
    nf = [len(f) for f in free]; nu_ = sum(nf); n2 = om.bm2.size
    Kf = sp.bmat([[K[0][0], K[0][1]], [K[1][0], K[1][1]]]).tocsr()
    Df = sp.hstack(D).tocsr()
    Bf = np.concatenate([Bg[free[0]], Bg[free[1]]])
    A = sp.bmat([[-Kf, Df.T], [Df, None]]).tocsc().astype(complex)
    M = sp.diags(np.concatenate([Bf, np.zeros(n2)])).tocsc().astype(complex)
    print("assembled: %d velocity + %d pressure unknowns, nnz %d, %.1f s" % (nu_, n2, A.nnz, time.time() - t0), flush=True)
    sigma0 = 0.02 + 0.76j                                   # near log(0.7378 + 0.7018 i): growth ~ 0.0156, St ~ 0.121
    lu = spla.splu((A - sigma0 * M).tocsc())
    print("factorised, %.1f s" % (time.time() - t0), flush=True)
    op = spla.LinearOperator(A.shape, matvec=lambda x: lu.solve(M @ x), dtype=complex)
    th, V = spla.eigs(op, k=a.nev, which="LM", tol=1e-12)
    lam = sigma0 + 1.0 / th
    order = np.argsort(-lam.real)
    lam = lam[order]; V = V[:, order]
    res = [float(np.linalg.norm(A @ V[:, i] - lam[i] * (M @ V[:, i])) / np.linalg.norm(V[:, i])) for i in range(len(lam))]
    mu = np.exp(lam)                                        # tau = 1
    out = {"case": "cylinder Re=50, lx1=6, lxd=9, shipped base flow (tests/golden/cylinder_case.npz)", "convection": a.conv, "unknowns": [int(nu_), int(n2)],
           "shift": [sigma0.real, sigma0.imag], "lambda": [[float(z.real), float(z.imag)] for z in lam], "residuals": res,
           "exp_lambda_modulus": [float(abs(z)) for z in mu], "exp_lambda": [[float(z.real), float(z.imag)] for z in mu],
           "seconds": time.time() - t0}
    print(json.dumps(out, indent=1))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
