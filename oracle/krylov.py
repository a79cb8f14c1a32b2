"""ORACLE (test infrastructure, not product): LightKrylov algorithms used by neklab.

CPU restatement of the LightKrylov routines neklab's drivers call
(`src/neklab_analysis.f90:80-81` eigs, `:136` svds, `:191-193` newton/gmres;
re-exports `src/neklab.f90:28-42`): classical Gram-Schmidt with
reorthogonalisation (`double_gram_schmidt_step`), `arnoldi`, Krylov-Schur `eigs`,
Golub-Kahan `svds`, restarted `gmres`.  LightKrylov is an un-vendored dependency
(`LightKrylov_setup.sh:55-57`, branch main, unpinned); algorithms restated from
SURVEY.md App. B.  Works on any vector type exposing the `abstract_vector_rdp`
surface (`copy/zero/scal/axpby/dot/norm`).

PARITY STATUS: parity unpinned except through the golden eigenvalue
(`test/neklabTests.py:44`).
"""
from __future__ import annotations

import math

import numpy as np

ATOL_DP = 10.0 ** (-15)            # 10**-precision(1.0_dp)
RTOL_DP = math.sqrt(ATOL_DP)       # LightKrylov rtol_dp (default eigs tolerance)


def innerprod(X, y):
    return np.array([x.dot(y) for x in X])


def double_gram_schmidt_step(y, X):
    """y <- (I - X X^T B)^2 y ; returns h = h1 + h2 (LightKrylov `double_gram_schmidt_step`)."""
    h = innerprod(X, y)
    for hi, x in zip(h, X):
        y.axpby(-hi, x, 1.0)
    h2 = innerprod(X, y)
    for hi, x in zip(h2, X):
        y.axpby(-hi, x, 1.0)
    return h + h2


def arnoldi_step(apply, X, H, k, tol=ATOL_DP):
    """One Arnoldi step: X[k+1] = A X[k], orthogonalised; fills H[:k+2, k] (0-based k)."""
    y = apply(X[k])
    h = double_gram_schmidt_step(y, X[:k + 1])
    H[:k + 1, k] = h
    beta = y.norm()
    H[k + 1, k] = beta
    if beta > tol:
        y.scal(1.0 / beta)
    if len(X) > k + 1:
        X[k + 1] = y
    else:
        X.append(y)
    return beta


def ritz_residuals(H, k):
    """eig of H[:k,:k] and LightKrylov residuals |H(k+1,k)| * |y_i(k)|."""
    lam, Y = np.linalg.eig(H[:k, :k])
    res = np.abs(H[k, k - 1] * Y[k - 1, :])
    return lam, Y, res


def krylov_schur_restart(X, H, k, select):
    """Stewart Krylov-Schur restart: keep the invariant subspace of the selected Ritz values.

    Returns the new kstart (number of kept vectors).  X[0:k+1], H[(k+1) x k] are updated in place
    so that A X[:p] = X[:p] H[:p,:p] + X[p] b^T  holds with p = number kept.
    """
    import scipy.linalg as sla
    Hk = H[:k, :k].copy()
    lam = np.linalg.eigvals(Hk)
    keep = select(lam)
    thresh = np.sort(np.abs(lam[keep]))[0] if keep.any() else np.inf
    # ordered real Schur form with the selected eigenvalues leading (conjugate pairs stay together)
    T, Z, sdim = sla.schur(Hk, output="real", sort=lambda re, im: math.hypot(re, im) >= thresh * (1 - 1e-12))
    p = int(sdim)
    b = H[k, k - 1] * Z[k - 1, :p]
    # new basis: X[:p] <- X[:k] Z[:, :p];  X[p] <- X[k]
    newX = []
    for j in range(p):
        v = X[0].copy(); v.zero()
        for i in range(k):
            v.axpby(Z[i, j], X[i], 1.0)
        newX.append(v)
    last = X[k]
    for j in range(p):
        X[j] = newX[j]
    X[p] = last
    del X[p + 1:]
    H[:, :] = 0.0
    H[:p, :p] = T[:p, :p]
    H[p, :p] = b
    return p


def select_above_median(lam):
    """LightKrylov default selector: keep Ritz values with |lambda| above the median."""
    a = np.abs(lam)
    return a > np.median(a)


def eigs(apply, x0, nev, kdim, tol=RTOL_DP, maxiter=10, log=None):
    """Krylov-Schur eigensolver (LightKrylov `eigs_rdp`), residual/convergence rule as upstream.

    Returns (lam, resid, X, Y, k): Ritz values sorted by |.| descending, with basis and Ritz
    vectors in the basis (eigvec_i = sum_j Y[j,i] X[j]).
    """
    X = [x0.copy()]
    nrm = X[0].norm()
    X[0].scal(1.0 / nrm)
    H = np.zeros((kdim + 1, kdim))
    kstart = 0
    conv = 0
    niter = 0
    lam = Y = res = None
    k = 0
    for outer in range(maxiter):
        for k in range(kstart, kdim):
            arnoldi_step(apply, X, H, k)
            lam, Y, res = ritz_residuals(H, k + 1)
            niter += 1
            conv = int((res < tol).sum())
            if log is not None:
                log(niter, k + 1, lam, res)
            if conv >= nev:
                break
        if conv >= nev:
            break
        kstart = krylov_schur_restart(X, H, kdim, select_above_median)
    order = np.argsort(-np.abs(lam))
    return lam[order], res[order], X, Y[:, order], k + 1


def svds(A, x0, nsv, kdim, tol=RTOL_DP):
    """Golub-Kahan bidiagonalisation (LightKrylov `svds`): per step one matvec + one rmatvec + DGS against U and V.

    A V_k = U_k B_k (B_k upper bidiagonal: diag alpha, superdiag beta), A^T U_k = V_k B_k^T + beta_k v_{k+1} e_k^T;
    residual of triplet i = |beta_k| * |last component of the left singular vector of B_k|.
    Returns (sigma descending, residuals, U, V, k).
    """
    U, V = [], []
    v = x0.copy(); v.scal(1.0 / v.norm())
    V.append(v)
    alpha, beta = [], []
    sig = res = None
    k = 0
    for k in range(kdim):
        u = A.matvec(V[k])
        if U:
            double_gram_schmidt_step(u, U)
        a = u.norm(); u.scal(1.0 / a)
        U.append(u); alpha.append(a)
        w = A.rmatvec(u)
        double_gram_schmidt_step(w, V)
        b = w.norm()
        if b > ATOL_DP:
            w.scal(1.0 / b)
        V.append(w); beta.append(b)
        n = k + 1
        B = np.zeros((n, n))
        for i in range(n):
            B[i, i] = alpha[i]
            if i + 1 < n:
                B[i, i + 1] = beta[i]
        P, sig, Qt = np.linalg.svd(B)
        res = np.abs(b * P[n - 1, :])
        if n >= nsv and int((res < tol).sum()) >= nsv:
            break
    svds.last_B = B                # the projected (bidiagonal) matrix of the final step, for singular-vector reconstruction
    return sig, res, U, V, k + 1


def gmres(apply, b, x0, kdim=30, atol=ATOL_DP, rtol=RTOL_DP, maxiter=10):
    """Restarted GMRES with DGS + Givens (LightKrylov `gmres_rdp`)."""
    x = x0.copy()
    bnorm = b.norm()
    tol = atol + rtol * bnorm
    for outer in range(maxiter):
        if outer == 0 and x.norm() == 0.0:
            r = b.copy(); r.nrst = 0                    # zero initial guess: no matvec (LightKrylov skips it too)
        else:
            r = apply(x); r.axpby(1.0, b, -1.0)         # r = b - A x
            r.nrst = 0                                  # a restart residual carries no rst fields (same choice as nlk_gmres)
        beta = r.norm()
        if beta < tol:
            break
        V = [r]; V[0].scal(1.0 / beta)
        H = np.zeros((kdim + 1, kdim))
        kk = 0
        for k in range(kdim):
            arnoldi_step(apply, V, H, k)
            kk = k + 1
            e1 = np.zeros(kk + 1); e1[0] = beta
            y, *_ = np.linalg.lstsq(H[:kk + 1, :kk], e1, rcond=None)
            rn = np.linalg.norm(H[:kk + 1, :kk] @ y - e1)
            if rn < tol:
                break
        for i in range(kk):
            x.axpby(y[i], V[i], 1.0)
    return x
