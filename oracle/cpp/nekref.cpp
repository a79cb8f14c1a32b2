// ORACLE (test infrastructure, NOT product): C++/OpenMP restatement of the PN-PN-2 perturbation time step that
// `exptA_matvec` drives (/root/reference/src/linops/exponential_propagator.f90:15-60 -> Nek5000 `nek_advance` with
// ifpert: fluidp/perturbv, makefp, advabp / advabp_adjoint, makextp, makebdfp, lagfieldp, cresvipp, ophinv -> hmholtz
// -> cggo, incomprp -> uzawa_gmres (+ Schwarz/coarse preconditioner of the same design as oracle/precond.py), heatp ->
// cdscalp, q_filter).  Nek5000 is an un-vendored, un-pinned dependency of the reference (Nek5000_setup.sh:56-58), so
// the routines are restated from their published algorithm; the in-tree mirrors are
// src/linops/neklab_linops.f90:285-312 (advabp ordering) and :343-362 (local_grad + metrics).
//
// Geometry, numbering, masks and the preconditioner's setup data come from the numpy oracle (oracle/mesh.py,
// oracle/precond.py); this file restates only the per-step arithmetic, threaded over elements, so that
//   (a) full-length applies on the reference's own configs finish in seconds (parity tests at 1e-10),
//   (b) bench.py's `cpu_baseline` / `--impl reference` arm can time the SAME workload as the GPU arm on the host cores.
// PARITY STATUS: checked against the numpy oracle operator by operator (tests/test_oracle_cpp.py); the numpy oracle is
// pinned by the KATs of tests/test_oracle_kat.py.  Nothing under neklab_b200/ may link or call this file.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <omp.h>

namespace {

typedef std::vector<double> vec;

struct Ref {
  int d = 2, n = 0, m = 0, q = 0; int64_t E = 0;
  int nz1 = 1, nz2 = 1, nzd = 1;
  size_t np1 = 0, np2 = 0, npd = 0, N1 = 0, N2 = 0; int ng = 3;
  vec D, Dt, I12, I12t, D12, D12t, I1d, I1dt, Dd, F1d;
  vec G, bm1, binvm1, vmult, bm2, rxw2, rxd, mask[4], diagA;   // diagA: un-assembled diag of A (h1 part), diagB = bm1
  std::vector<int64_t> gs_off, gs_idx; int64_t nglob = 0;
  int has_outflow = 0; double volvm1 = 0, volvm2 = 0;
  // preconditioner (oracle/precond.py SchwarzCoarse)
  int have_pre = 0; vec S, St, dinv, wt, A0inv, shape; std::vector<int64_t> vertex; int64_t nv = 0;
  // parameters
  double visc = 1, rho = 1, vtol = 1e-9, ptol = 1e-7, ttol = 1e-9, cond = 1, rhocp = 1, buoy[3] = {0, 0, 0}, fw = 0;
  int torder = 3, ifheat = 0, cg_maxit = 1000, gm_maxit = 100, lgmres = 30, adjoint = 0, nonlinear = 0, variant = 0;
  double dt = 0;
  // state
  vec U[3], T, vp[3], prp, tp, vlag[2][3], exx1[3], exx2[3], prlag, tlag[2], vg1, vg2, forcing[3]; int has_forcing = 0;
  long cg_iters = 0, gm_iters = 0;
  int pr_proj = 0; std::vector<vec> projX, projEX;      // pressure residual projection (Nek setrhsp/gensolnp; .par residualProj, SIZE mxprev)
};

// out[(k,j,i')] = sum_i M[i'][i] in[(k,j,i)] along axis ax (0 = x fastest, 1 = y, 2 = z); dims of `in` are (n0,n1,n2) = (x,y,z)
static void tapply(const double* M, int mo, int mi, int ax, const double* in, int n0, int n1, int n2, double* out) {
  if (ax == 0) {
    for (int k = 0; k < n2; ++k) for (int j = 0; j < n1; ++j) {
      const double* a = in + ((size_t)k * n1 + j) * n0; double* o = out + ((size_t)k * n1 + j) * mo;
      for (int r = 0; r < mo; ++r) { double s = 0; const double* mr = M + (size_t)r * mi; for (int i = 0; i < mi; ++i) s += mr[i] * a[i]; o[r] = s; }
    }
  } else if (ax == 1) {
    for (int k = 0; k < n2; ++k) for (int r = 0; r < mo; ++r) {
      double* o = out + ((size_t)k * mo + r) * n0;
      for (int i = 0; i < n0; ++i) o[i] = 0;
      for (int j = 0; j < mi; ++j) { const double c = M[(size_t)r * mi + j]; const double* a = in + ((size_t)k * n1 + j) * n0; for (int i = 0; i < n0; ++i) o[i] += c * a[i]; }
    }
  } else {
    const size_t pl = (size_t)n0 * n1;
    for (int r = 0; r < mo; ++r) {
      double* o = out + (size_t)r * pl;
      for (size_t i = 0; i < pl; ++i) o[i] = 0;
      for (int k = 0; k < mi; ++k) { const double c = M[(size_t)r * mi + k]; const double* a = in + (size_t)k * pl; for (size_t i = 0; i < pl; ++i) o[i] += c * a[i]; }
    }
  }
}
// tensor application of (possibly different) matrices along each axis: A0 along x, A1 along y, A2 along z (3-D only)
static void tensor3(const Ref& R, const double* A0, const double* A1, const double* A2, int mo, int mi, const double* in, double* out, double* t1, double* t2) {
  if (R.d == 2) { tapply(A0, mo, mi, 0, in, mi, mi, 1, t1); tapply(A1, mo, mi, 1, t1, mo, mi, 1, out); return; }
  tapply(A0, mo, mi, 0, in, mi, mi, mi, t1); tapply(A1, mo, mi, 1, t1, mo, mi, mi, t2); tapply(A2, mo, mi, 2, t2, mo, mo, mi, out);
}

static void dssum(const Ref& R, double* u) {
  const int64_t ng = R.nglob;
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < ng; ++g) {
    const int64_t a = R.gs_off[g], b = R.gs_off[g + 1];
    if (b - a < 2) continue;
    double s = 0; for (int64_t t = a; t < b; ++t) s += u[R.gs_idx[t]];
    for (int64_t t = a; t < b; ++t) u[R.gs_idx[t]] = s;
  }
}

// Nek axhelm: w = h1 * D^T G D u + h2 * bm1 * u   (element-local)
static void axhelm(const Ref& R, const double* u, double* w, double h1, double h2) {
  const int n = R.n, d = R.d, nz = R.nz1; const size_t np = R.np1;
#pragma omp parallel
  {
    vec ur(np), us(np), ut(np), t(np);
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) {
      const double* ue = u + e * np; double* we = w + e * np; const double* g = R.G.data() + (size_t)e * R.ng * np; const double* b = R.bm1.data() + e * np;
      tapply(R.D.data(), n, n, 0, ue, n, n, nz, ur.data()); tapply(R.D.data(), n, n, 1, ue, n, n, nz, us.data());
      if (d == 3) tapply(R.D.data(), n, n, 2, ue, n, n, nz, ut.data());
      if (d == 2) for (size_t i = 0; i < np; ++i) { double a = ur[i], c = us[i]; ur[i] = g[i] * a + g[2 * np + i] * c; us[i] = g[2 * np + i] * a + g[np + i] * c; }
      else for (size_t i = 0; i < np; ++i) { double a = ur[i], c = us[i], f = ut[i];
        ur[i] = g[i] * a + g[3 * np + i] * c + g[4 * np + i] * f; us[i] = g[3 * np + i] * a + g[np + i] * c + g[5 * np + i] * f; ut[i] = g[4 * np + i] * a + g[5 * np + i] * c + g[2 * np + i] * f; }
      for (size_t i = 0; i < np; ++i) we[i] = h2 * b[i] * ue[i];
      tapply(R.Dt.data(), n, n, 0, ur.data(), n, n, nz, t.data()); for (size_t i = 0; i < np; ++i) we[i] += h1 * t[i];
      tapply(R.Dt.data(), n, n, 1, us.data(), n, n, nz, t.data()); for (size_t i = 0; i < np; ++i) we[i] += h1 * t[i];
      if (d == 3) { tapply(R.Dt.data(), n, n, 2, ut.data(), n, n, nz, t.data()); for (size_t i = 0; i < np; ++i) we[i] += h1 * t[i]; }
    }
  }
}

// Nek opdiv/multd: p = scale * sum_c sum_k rxw2[k][c] * (D12 along k, I12 elsewhere) u_c ; pre[c] (optional) multiplies u_c first
static void opdiv(const Ref& R, const double* const u[3], double* p, double scale, const double* const pre[3] = nullptr, double prescale = 1.0) {
  const int n = R.n, q = R.q, d = R.d; const size_t np = R.np1, nq = R.np2;
#pragma omp parallel
  {
    vec t1(np), t2(np), o(nq), uu(np);
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) {
      double* pe = p + e * nq; for (size_t i = 0; i < nq; ++i) pe[i] = 0;
      const double* rx = R.rxw2.data() + (size_t)e * d * d * nq;
      for (int c = 0; c < d; ++c) {
        const double* ue = u[c] + e * np;
        if (pre) { const double* pc = pre[c] + e * np; for (size_t i = 0; i < np; ++i) uu[i] = ue[i] * pc[i] * prescale; ue = uu.data(); }
        for (int k = 0; k < d; ++k) {
          const double* A0 = k == 0 ? R.D12.data() : R.I12.data(); const double* A1 = k == 1 ? R.D12.data() : R.I12.data(); const double* A2 = k == 2 ? R.D12.data() : R.I12.data();
          tensor3(R, A0, A1, A2, q, n, ue, o.data(), t1.data(), t2.data());
          const double* r = rx + (size_t)(k * d + c) * nq;
          for (size_t i = 0; i < nq; ++i) pe[i] += scale * r[i] * o[i];
        }
      }
    }
  }
}
// Nek opgradt/cdtp: w_c = sum_k (D12^T along k, I12^T elsewhere) (p * rxw2[k][c])
static void opgradt(const Ref& R, const double* p, double* const w[3]) {
  const int n = R.n, q = R.q, d = R.d; const size_t np = R.np1, nq = R.np2;
#pragma omp parallel
  {
    vec t1(np), t2(np), o(np), pp(nq);
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) {
      const double* pe = p + e * nq; const double* rx = R.rxw2.data() + (size_t)e * d * d * nq;
      for (int c = 0; c < d; ++c) {
        double* we = w[c] + e * np; for (size_t i = 0; i < np; ++i) we[i] = 0;
        for (int k = 0; k < d; ++k) {
          const double* r = rx + (size_t)(k * d + c) * nq; for (size_t i = 0; i < nq; ++i) pp[i] = pe[i] * r[i];
          const double* A0 = k == 0 ? R.D12t.data() : R.I12t.data(); const double* A1 = k == 1 ? R.D12t.data() : R.I12t.data(); const double* A2 = k == 2 ? R.D12t.data() : R.I12t.data();
          tensor3(R, A0, A1, A2, n, q, pp.data(), o.data(), t1.data(), t2.data());
          for (size_t i = 0; i < np; ++i) we[i] += o[i];
        }
      }
    }
  }
}

// dealiased convection (Nek convect_new, ifcf = ifuf = .false.): for every field f of `nf` fields,
//   out_f += scale * I1d^T [ sum_k (sum_c rxd[k][c] * I1d C_c) * Dd_k (I1d u_f) ]
// and optionally the "swapped" term with a second pair (u2_f convected by C2) in the same pass.
static void convect(const Ref& R, int nf, const double* const u[4], const double* const C[3], double* const out[4], double scale,
                    const double* const u2[4] = nullptr, const double* const C2[3] = nullptr) {
  const int n = R.n, m = R.m, d = R.d, nzd = R.nzd; const size_t np = R.np1, nd = R.npd;
#pragma omp parallel
  {
    vec t1(nd), t2(nd), uf(nd), g(nd), acc(nd), o(np), tr[3], tr2[3], cf(nd);
    for (int k = 0; k < d; ++k) { tr[k].resize(nd); tr2[k].resize(nd); }
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) {
      const double* rx = R.rxd.data() + (size_t)e * d * d * nd;
      for (int pass = 0; pass < (C2 ? 2 : 1); ++pass) {
        const double* const* CC = pass ? C2 : C; vec* T = pass ? tr2 : tr;
        for (int k = 0; k < d; ++k) std::fill(T[k].begin(), T[k].end(), 0.0);
        for (int c = 0; c < d; ++c) {
          tensor3(R, R.I1d.data(), R.I1d.data(), R.I1d.data(), m, n, CC[c] + e * np, cf.data(), t1.data(), t2.data());
          for (int k = 0; k < d; ++k) { const double* r = rx + (size_t)(k * d + c) * nd; double* tk = T[k].data(); for (size_t i = 0; i < nd; ++i) tk[i] += r[i] * cf[i]; }
        }
      }
      for (int f = 0; f < nf; ++f) {
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int pass = 0; pass < (u2 ? 2 : 1); ++pass) {
          const double* src = (pass ? u2[f] : u[f]) + e * np; vec* T = pass ? tr2 : tr;
          tensor3(R, R.I1d.data(), R.I1d.data(), R.I1d.data(), m, n, src, uf.data(), t1.data(), t2.data());
          for (int k = 0; k < d; ++k) {
            tapply(R.Dd.data(), m, m, k, uf.data(), m, m, nzd, g.data());
            const double* tk = T[k].data(); for (size_t i = 0; i < nd; ++i) acc[i] += tk[i] * g[i];
          }
        }
        tensor3(R, R.I1dt.data(), R.I1dt.data(), R.I1dt.data(), n, m, acc.data(), o.data(), t1.data(), t2.data());
        double* oe = out[f] + e * np; for (size_t i = 0; i < np; ++i) oe[i] += scale * o[i];
      }
    }
  }
}
// Nek convect_adj: out_i += scale * I1d^T [ sum_j (I1d c_j) * sum_k rxd[k][i] Dd_k (I1d U_j) ]
static void convect_adj(const Ref& R, const double* const U[3], const double* const c[3], double* const out[3], double scale) {
  const int n = R.n, m = R.m, d = R.d, nzd = R.nzd; const size_t np = R.np1, nd = R.npd;
#pragma omp parallel
  {
    vec t1(nd), t2(nd), Uf(nd), cf(nd), g(nd), o(np), acc[3];
    for (int i = 0; i < d; ++i) acc[i].resize(nd);
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) {
      const double* rx = R.rxd.data() + (size_t)e * d * d * nd;
      for (int i = 0; i < d; ++i) std::fill(acc[i].begin(), acc[i].end(), 0.0);
      for (int j = 0; j < d; ++j) {
        tensor3(R, R.I1d.data(), R.I1d.data(), R.I1d.data(), m, n, U[j] + e * np, Uf.data(), t1.data(), t2.data());
        tensor3(R, R.I1d.data(), R.I1d.data(), R.I1d.data(), m, n, c[j] + e * np, cf.data(), t1.data(), t2.data());
        for (int k = 0; k < d; ++k) {
          tapply(R.Dd.data(), m, m, k, Uf.data(), m, m, nzd, g.data());
          for (int i = 0; i < d; ++i) { const double* r = rx + (size_t)(k * d + i) * nd; double* a = acc[i].data(); for (size_t t = 0; t < nd; ++t) a[t] += cf[t] * r[t] * g[t]; }
        }
      }
      for (int i = 0; i < d; ++i) {
        tensor3(R, R.I1dt.data(), R.I1dt.data(), R.I1dt.data(), n, m, acc[i].data(), o.data(), t1.data(), t2.data());
        double* oe = out[i] + e * np; for (size_t t = 0; t < np; ++t) oe[t] += scale * o[t];
      }
    }
  }
}

static double dot(const double* a, const double* b, size_t n) { double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s; }
static double dot3(const double* a, const double* b, const double* w, size_t n) { double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (size_t i = 0; i < n; ++i) s += a[i] * b[i] * w[i];
  return s; }
static void ortho(const Ref& R, double* p) {
  if (R.has_outflow) return;
  double s = 0; const size_t n = R.N2;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (size_t i = 0; i < n; ++i) s += p[i];
  s /= (double)n;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) p[i] -= s;
}

// Nek cdabdtp (intype = 1): E p = D (mask binvm1/rho dssum(D^T p))
static void apply_E(Ref& R, const double* p, double* ep, vec w[3]) {
  double* ww[3] = {w[0].data(), w[1].data(), w[2].data()};
  opgradt(R, p, ww);
  for (int c = 0; c < R.d; ++c) { dssum(R, ww[c]); double* a = ww[c]; const double* mk = R.mask[c].data(); const double* bi = R.binvm1.data(); const double ir = 1.0 / R.rho;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < R.N1; ++i) a[i] *= mk[i] * bi[i] * ir; }
  const double* cw[3] = {ww[0], ww[1], ww[2]};
  opdiv(R, cw, ep, 1.0);
}

// Nek hmholtz + cggo: x = (h1 A + h2 B)^-1 mask dssum(f)   (Jacobi-PCG, norm sqrt(sum r^2 mult binv / vol))
static int cggo(Ref& R, double* f, double h1, double h2, const double* mask, double tol, double* x) {
  const size_t N = R.N1;
  vec diag(N), r(N), p(N, 0.0), w(N), z(N);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < N; ++i) diag[i] = h1 * R.diagA[i] + h2 * R.bm1[i];
  dssum(R, diag.data());
  dssum(R, f);
  double rmax = 0;
#pragma omp parallel for reduction(max : rmax) schedule(static)
  for (size_t i = 0; i < N; ++i) { r[i] = mask[i] * f[i]; x[i] = 0; rmax = std::max(rmax, std::fabs(r[i])); }
  if (rmax == 0.0) return 0;
  double rtz1 = 1.0; int it = 0;
  for (it = 1; it <= R.cg_maxit; ++it) {
    double a = 0, b = 0;
#pragma omp parallel for reduction(+ : a, b) schedule(static)
    for (size_t i = 0; i < N; ++i) { z[i] = r[i] / diag[i]; a += z[i] * r[i] * R.vmult[i]; b += r[i] * r[i] * R.vmult[i] * R.binvm1[i]; }
    double rtz2 = rtz1; rtz1 = a;
    if (std::sqrt(b / R.volvm1) <= tol) { --it; break; }
    double beta = it == 1 ? 0.0 : rtz1 / rtz2;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; ++i) p[i] = z[i] + beta * p[i];
    axhelm(R, p.data(), w.data(), h1, h2); dssum(R, w.data());
    double rho = 0;
#pragma omp parallel for reduction(+ : rho) schedule(static)
    for (size_t i = 0; i < N; ++i) { w[i] *= mask[i]; rho += w[i] * p[i] * R.vmult[i]; }
    double alpha = rtz1 / rho;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * w[i]; }
  }
  if (it > R.cg_maxit) it = R.cg_maxit;
  R.cg_iters += it; return it;
}

// ---- preconditioner (same design as oracle/precond.py: overlapping Schwarz with FDM local solves + vertex coarse grid)
static inline size_t idx3(int n, int k, int j, int i) { return ((size_t)k * n + j) * n + i; }
// copy (dir 0: face <- inner, fwd) helpers over the tangentially-interior part of the face layer of direction k, side s
template <class F> static void for_face(const Ref& R, int k, int side, F f) {
  const int n = R.n, d = R.d;
  const int fpos = side == 0 ? 0 : n - 1, ipos = side == 0 ? 1 : n - 2;
  const int zlo = d == 3 ? 1 : 0, zhi = d == 3 ? n - 1 : 1;
  if (k == 0) { for (int kk = zlo; kk < zhi; ++kk) for (int j = 1; j < n - 1; ++j) f(idx3(n, kk, j, fpos), idx3(n, kk, j, ipos)); }
  else if (k == 1) { for (int kk = zlo; kk < zhi; ++kk) for (int i = 1; i < n - 1; ++i) f(idx3(n, kk, fpos, i), idx3(n, kk, ipos, i)); }
  else { for (int j = 1; j < n - 1; ++j) for (int i = 1; i < n - 1; ++i) f(idx3(n, fpos, j, i), idx3(n, ipos, j, i)); }
}
static void precond(Ref& R, const double* r, double* z) {
  const int n = R.n, q = R.q, d = R.d; const size_t np = R.np1, nq = R.np2; const size_t N = R.N1;
  if (!R.have_pre) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < R.N2; ++i) z[i] = r[i] / R.bm2[i];
    return;
  }
  vec w(N, 0.0), s(N), zz(N), t(N, 0.0);
  const int zq = d == 3 ? q : 1;
  // embed + exchange_fwd
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < R.E; ++e) {
    double* we = w.data() + e * np; const double* re = r + e * nq;
    for (int kk = 0; kk < zq; ++kk) for (int j = 0; j < q; ++j) for (int i = 0; i < q; ++i) we[idx3(n, d == 3 ? kk + 1 : 0, j + 1, i + 1)] = re[((size_t)kk * q + j) * q + i];
    for (int k = 0; k < d; ++k) for (int sd = 0; sd < 2; ++sd) for_face(R, k, sd, [&](size_t f, size_t in) { we[f] = we[in]; });
  }
  s = w; dssum(R, s.data());
#pragma omp parallel
  {
    vec a(np), b(np);
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) {
      double* we = w.data() + e * np; const double* se = s.data() + e * np;
      for (int k = 0; k < d; ++k) for (int sd = 0; sd < 2; ++sd) for_face(R, k, sd, [&](size_t f, size_t in) { we[f] = se[f] - we[in]; });
      // FDM: S^T along every direction, scale, S along every direction
      const double* Se = R.S.data() + (size_t)e * d * n * n; const double* Ste = R.St.data() + (size_t)e * d * n * n;
      const double* cur = we; double* bufs[2] = {a.data(), b.data()}; int wbuf = 0;
      for (int k = 0; k < d; ++k) { tapply(Ste + (size_t)k * n * n, n, n, k, cur, n, n, R.nz1, bufs[wbuf]); cur = bufs[wbuf]; wbuf ^= 1; }
      double* c2 = const_cast<double*>(cur); const double* di = R.dinv.data() + e * np;
      for (size_t i = 0; i < np; ++i) c2[i] *= di[i];
      for (int k = 0; k < d; ++k) { tapply(Se + (size_t)k * n * n, n, n, k, cur, n, n, R.nz1, bufs[wbuf]); cur = bufs[wbuf]; wbuf ^= 1; }
      double* ze = zz.data() + e * np; double* te = t.data() + e * np;
      for (size_t i = 0; i < np; ++i) ze[i] = cur[i];
      for (int k = 0; k < d; ++k) for (int sd = 0; sd < 2; ++sd) for_face(R, k, sd, [&](size_t f, size_t) { te[f] = ze[f]; });
    }
  }
  s = t; dssum(R, s.data());
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < R.E; ++e) {
    double* ze = zz.data() + e * np; const double* se = s.data() + e * np; const double* te = t.data() + e * np;
    for (int k = 0; k < d; ++k) for (int sd = 0; sd < 2; ++sd) for_face(R, k, sd, [&](size_t f, size_t in) { ze[in] += se[f] - te[f]; });
    double* oe = z + e * nq; const double* wt = R.wt.data() + e * nq;
    for (int kk = 0; kk < zq; ++kk) for (int j = 0; j < q; ++j) for (int i = 0; i < q; ++i) { size_t o = ((size_t)kk * q + j) * q + i; oe[o] = ze[idx3(n, d == 3 ? kk + 1 : 0, j + 1, i + 1)] * wt[o]; }
  }
  if (R.nv > 0) {
    const int nc = 1 << d; vec rc((size_t)R.nv, 0.0), yc((size_t)R.nv);
    for (int64_t e = 0; e < R.E; ++e) for (int c = 0; c < nc; ++c) {
      const double* sh = R.shape.data() + (size_t)c * nq; const double* re = r + e * nq; double sacc = 0;
      for (size_t i = 0; i < nq; ++i) sacc += re[i] * sh[i];
      rc[R.vertex[e * nc + c] - 1] += sacc;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < R.nv; ++i) { double sacc = 0; const double* row = R.A0inv.data() + (size_t)i * R.nv; for (int64_t j = 0; j < R.nv; ++j) sacc += row[j] * rc[j]; yc[i] = sacc; }
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) for (int c = 0; c < nc; ++c) {
      const double v = yc[R.vertex[e * nc + c] - 1]; const double* sh = R.shape.data() + (size_t)c * nq; double* oe = z + e * nq;
      for (size_t i = 0; i < nq; ++i) oe[i] += v * sh[i];
    }
  }
}

// Nek uzawa_gmres (gmres.f): right-preconditioned, mass-scaled restarted GMRES(lgmres), classical Gram-Schmidt
static int uzawa_gmres(Ref& R, const double* res, double tol, double* x) {
  const size_t N = R.N2; const int m = R.lgmres;
  vec ml(N), mu(N), r(N), w(N), tmp(N), wk[3];
  for (int c = 0; c < R.d; ++c) wk[c].resize(R.N1);
  for (size_t i = 0; i < N; ++i) { ml[i] = std::sqrt(1.0 / R.bm2[i]); mu[i] = std::sqrt(R.bm2[i]); }
  const double norm_fac = 1.0 / std::sqrt(R.volvm2);
  std::fill(x, x + N, 0.0);
  std::vector<vec> V, Z; vec H((size_t)(m + 1) * m), cs(m), sn(m), gamma(m + 2), cc(m);
  int it = 0; bool conv = false;
  while (!conv && it < R.gm_maxit) {
    if (it == 0) for (size_t i = 0; i < N; ++i) r[i] = ml[i] * res[i];
    else { apply_E(R, x, tmp.data(), wk); for (size_t i = 0; i < N; ++i) r[i] = ml[i] * (res[i] - tmp[i]); }
    gamma[0] = std::sqrt(dot(r.data(), r.data(), N));
    if (gamma[0] == 0.0) break;
    V.assign(1, r); for (size_t i = 0; i < N; ++i) V[0][i] /= gamma[0];
    Z.clear(); std::fill(H.begin(), H.end(), 0.0);
    int j = 0;
    for (j = 1; j <= m; ++j) {
      ++it;
      for (size_t i = 0; i < N; ++i) tmp[i] = mu[i] * V[j - 1][i];
      vec z(N); precond(R, tmp.data(), z.data()); ortho(R, z.data());
      Z.push_back(z);
      apply_E(R, Z.back().data(), w.data(), wk);
      for (size_t i = 0; i < N; ++i) w[i] *= ml[i];
      vec h(j);
      for (int i2 = 0; i2 < j; ++i2) h[i2] = dot(w.data(), V[i2].data(), N);
      for (int i2 = 0; i2 < j; ++i2) { const double hh = h[i2]; const double* v = V[i2].data();
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < N; ++i) w[i] -= hh * v[i]; }
      for (int i2 = 0; i2 < j; ++i2) H[(size_t)i2 * m + j - 1] = h[i2];
      for (int i2 = 0; i2 < j - 1; ++i2) { double t = H[(size_t)i2 * m + j - 1]; H[(size_t)i2 * m + j - 1] = cs[i2] * t + sn[i2] * H[(size_t)(i2 + 1) * m + j - 1]; H[(size_t)(i2 + 1) * m + j - 1] = -sn[i2] * t + cs[i2] * H[(size_t)(i2 + 1) * m + j - 1]; }
      double alpha = std::sqrt(dot(w.data(), w.data(), N));
      if (alpha == 0.0) { conv = true; break; }
      double hjj = H[(size_t)(j - 1) * m + j - 1], l = std::sqrt(hjj * hjj + alpha * alpha);
      cs[j - 1] = hjj / l; sn[j - 1] = alpha / l; H[(size_t)(j - 1) * m + j - 1] = l;
      gamma[j] = -sn[j - 1] * gamma[j - 1]; gamma[j - 1] = cs[j - 1] * gamma[j - 1];
      if (std::fabs(gamma[j]) * norm_fac < tol) { conv = true; break; }
      if (j == m || it >= R.gm_maxit) break;
      vec v(N); for (size_t i = 0; i < N; ++i) v[i] = w[i] / alpha; V.push_back(v);
    }
    if (j > m) j = m;
    for (int k = j - 1; k >= 0; --k) { double t = gamma[k]; for (int i2 = j - 1; i2 > k; --i2) t -= H[(size_t)k * m + i2] * cc[i2]; cc[k] = t / H[(size_t)k * m + k]; }
    for (int i2 = 0; i2 < j; ++i2) { const double c = cc[i2]; const double* zv = Z[i2].data();
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < N; ++i) x[i] += c * zv[i]; }
  }
  ortho(R, x);
  R.gm_iters += it; return it;
}

// Pressure residual projection (Nek `setrhsp`/`gensolnp`): the rhs is projected onto the span of up to pr_proj previous
// solutions kept E-orthonormal; same restatement as the device path (nlk_solver.cu pressure_solve_projected).
static int pressure_projected(Ref& R, double* rhs, double tol, double* x) {
  const size_t N = R.N2; const int mx = R.pr_proj;
  if (mx <= 0) return uzawa_gmres(R, rhs, tol, x);
  const int m = (int)R.projX.size();
  vec xbar(N, 0.0), w(N), wk[3]; for (int c = 0; c < R.d; ++c) wk[c].resize(R.N1);
  for (int i = 0; i < m; ++i) { const double al = dot(R.projX[i].data(), rhs, N); const double* X = R.projX[i].data(); const double* EX = R.projEX[i].data();
#pragma omp parallel for schedule(static)
    for (size_t t = 0; t < N; ++t) { xbar[t] += al * X[t]; rhs[t] -= al * EX[t]; } }
  int it = uzawa_gmres(R, rhs, tol, x);
  if (m == mx) {
    for (size_t t = 0; t < N; ++t) x[t] += xbar[t];
    apply_E(R, x, w.data(), wk);
    const double nr = 1.0 / std::sqrt(dot(x, w.data(), N));
    R.projX.assign(1, vec(x, x + N)); R.projEX.assign(1, w);
    for (size_t t = 0; t < N; ++t) { R.projX[0][t] *= nr; R.projEX[0][t] *= nr; }
    return it;
  }
  apply_E(R, x, w.data(), wk);
  vec dl(x, x + N);
  for (int i = 0; i < m; ++i) { const double be = dot(R.projX[i].data(), w.data(), N); const double* X = R.projX[i].data(); const double* EX = R.projEX[i].data();
    for (size_t t = 0; t < N; ++t) { dl[t] -= be * X[t]; w[t] -= be * EX[t]; } }
  const double nn = dot(dl.data(), w.data(), N);
  if (nn > 0) { const double nr = 1.0 / std::sqrt(nn); for (size_t t = 0; t < N; ++t) { dl[t] *= nr; w[t] *= nr; } R.projX.push_back(dl); R.projEX.push_back(w); }
  if (m > 0) for (size_t t = 0; t < N; ++t) x[t] += xbar[t];
  return it;
}

static void bdf_coeffs(int nbd, double* bd) { bd[0] = bd[1] = bd[2] = bd[3] = 0; if (nbd == 1) { bd[0] = 1; bd[1] = 1; } else if (nbd == 2) { bd[0] = 1.5; bd[1] = 2; bd[2] = -0.5; } else { bd[0] = 11.0 / 6.0; bd[1] = 3; bd[2] = -1.5; bd[3] = 1.0 / 3.0; } }
static void ab_coeffs(int nab, int nbd, double* ab) {      // Nek setabbd with constant dt
  ab[0] = 1; ab[1] = ab[2] = 0;
  if (nab == 1) return;
  if (nab == 2) { if (nbd <= 2) { ab[0] = 1.5; ab[1] = -0.5; } else { ab[0] = 2; ab[1] = -1; } return; }
  if (nbd == 1) { ab[2] = 0.5 * (0.5 + 1.0 / 3.0); ab[1] = -0.5 - 2 * ab[2]; ab[0] = 1 - ab[1] - ab[2]; }
  else if (nbd == 2) { ab[2] = 2.0 / 3.0; ab[1] = -1 - 2 * ab[2]; ab[0] = 1 - ab[1] - ab[2]; }
  else { ab[0] = 3; ab[1] = -3; ab[2] = 1; }
}

static void filter(const Ref& R, double* u) {
  const int n = R.n; const size_t np = R.np1;
#pragma omp parallel
  { vec t1(np), t2(np), o(np);
#pragma omp for schedule(static)
    for (int64_t e = 0; e < R.E; ++e) { tensor3(R, R.F1d.data(), R.F1d.data(), R.F1d.data(), n, n, u + e * np, o.data(), t1.data(), t2.data()); std::copy(o.begin(), o.end(), u + e * np); } }
}

// one Nek `nek_advance` in perturbation mode
static void advance(Ref& R, int istep) {
  const int d = R.d; const size_t N1 = R.N1, N2 = R.N2; const double dt = R.dt, rho = R.rho;
  const int nbd = std::min(istep, R.torder), nab = std::min(istep, 3);
  double bd[4], ab[3]; bdf_coeffs(nbd, bd); ab_coeffs(nab, nbd, ab);
  vec bf[3], bq;
  for (int c = 0; c < d; ++c) { bf[c].assign(N1, 0.0);
    for (size_t i = 0; i < N1; ++i) { double f = 0; if (R.ifheat && R.buoy[c] != 0.0 && !R.adjoint) f += R.buoy[c] * R.tp[i]; if (R.has_forcing) f += R.forcing[c][i]; bf[c][i] = f * R.bm1[i]; } }
  const double* Ub[3] = {R.U[0].data(), R.U[1].data(), R.U[2].data()}; const double* up[3] = {R.vp[0].data(), R.vp[1].data(), R.vp[2].data()};
  double* bfp[4] = {bf[0].data(), bf[1].data(), bf[2].data(), nullptr};
  const double* Ub4[4] = {Ub[0], Ub[1], Ub[2], nullptr}; const double* up4[4] = {up[0], up[1], up[2], nullptr};
  if (R.nonlinear) convect(R, d, up4, up, bfp, -rho);
  else if (!R.adjoint) convect(R, d, Ub4, up, bfp, -rho, up4, Ub);                     // u'.grad U + U.grad u'
  else {
    double* b3[3] = {bfp[0], bfp[1], bfp[2]};
    convect_adj(R, Ub, up, b3, -rho); convect(R, d, up4, Ub, bfp, +rho);
    if (R.ifheat) {                                                                    // - theta' grad(T_base)   (adjoint of u'.grad T)
      // strong gradient of the base temperature on the fine mesh, weak form like convect_adj with c = theta'
      const double* Tb[3] = {R.T.data(), R.T.data(), R.T.data()}; (void)Tb;
      const int n = R.n, m = R.m; const size_t np = R.np1, nd = R.npd;
#pragma omp parallel
      { vec t1(nd), t2(nd), Tf(nd), cf(nd), g(nd), o(np), acc(nd);
#pragma omp for schedule(static)
        for (int64_t e = 0; e < R.E; ++e) {
          const double* rx = R.rxd.data() + (size_t)e * d * d * nd;
          tensor3(R, R.I1d.data(), R.I1d.data(), R.I1d.data(), m, n, R.T.data() + e * np, Tf.data(), t1.data(), t2.data());
          tensor3(R, R.I1d.data(), R.I1d.data(), R.I1d.data(), m, n, R.tp.data() + e * np, cf.data(), t1.data(), t2.data());
          for (int i = 0; i < d; ++i) {
            std::fill(acc.begin(), acc.end(), 0.0);
            for (int k = 0; k < d; ++k) { tapply(R.Dd.data(), m, m, k, Tf.data(), m, m, R.nzd, g.data()); const double* r = rx + (size_t)(k * d + i) * nd; for (size_t t = 0; t < nd; ++t) acc[t] += cf[t] * r[t] * g[t]; }
            tensor3(R, R.I1dt.data(), R.I1dt.data(), R.I1dt.data(), n, m, acc.data(), o.data(), t1.data(), t2.data());
            double* oe = bf[i].data() + e * np; for (size_t t = 0; t < np; ++t) oe[t] -= R.rhocp * o[t];
          }
        } }
    }
  }
  // makextp, makebdfp, lagfieldp
  for (int c = 0; c < d; ++c) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N1; ++i) {
      double ta = ab[1] * R.exx1[c][i] + ab[2] * R.exx2[c][i];
      R.exx2[c][i] = R.exx1[c][i]; R.exx1[c][i] = bf[c][i];
      double tb = bd[1] * R.vp[c][i] + bd[2] * R.vlag[0][c][i] + bd[3] * R.vlag[1][c][i];
      bf[c][i] = ab[0] * bf[c][i] + ta + tb * R.bm1[i] * (rho / dt);
      R.vlag[1][c][i] = R.vlag[0][c][i]; R.vlag[0][c][i] = R.vp[c][i];
    }
  }
  if (R.ifheat) {
    bq.assign(N1, 0.0);
    double* bq4[4] = {bq.data(), nullptr, nullptr, nullptr}; const double* T4[4] = {R.T.data(), nullptr, nullptr, nullptr}; const double* tp4[4] = {R.tp.data(), nullptr, nullptr, nullptr};
    if (!R.adjoint) convect(R, 1, T4, up, bq4, -R.rhocp, tp4, Ub);                       // u'.grad T + U.grad T'
    else {
      convect(R, 1, tp4, Ub, bq4, +R.rhocp);                                            // adjoint: + U.grad theta'
      for (int c = 0; c < d; ++c) if (R.buoy[c] != 0.0) for (size_t i = 0; i < N1; ++i) bq[i] += R.buoy[c] * R.vp[c][i] * R.bm1[i] * rho;   // + buoyancy . u'
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N1; ++i) {
      double ta = ab[1] * R.vg1[i] + ab[2] * R.vg2[i];
      R.vg2[i] = R.vg1[i]; R.vg1[i] = bq[i];
      double tb = bd[1] * R.tp[i] + bd[2] * R.tlag[0][i] + bd[3] * R.tlag[1][i];
      bq[i] = ab[0] * bq[i] + ta + tb * R.bm1[i] * (R.rhocp / dt);
      R.tlag[1][i] = R.tlag[0][i]; R.tlag[0][i] = R.tp[i];
    }
  }
  // igeom = 2: velocity
  const double h1 = R.visc, h2 = rho * bd[0] / dt;
  if (!R.nonlinear) for (int c = 0; c < d; ++c) for (size_t i = 0; i < N1; ++i) R.vp[c][i] *= R.mask[c][i];
  vec pext(N2);
  if (nbd == 3 && !(R.variant & 8)) for (size_t i = 0; i < N2; ++i) pext[i] = 2.0 * R.prp[i] - R.prlag[i]; else pext = R.prp;
  vec gp[3], hw(N1), dv(N1);
  for (int c = 0; c < d; ++c) gp[c].resize(N1);
  { double* g3[3] = {gp[0].data(), gp[1].data(), gp[2].data()}; opgradt(R, pext.data(), g3); }
  for (int c = 0; c < d; ++c) {
    axhelm(R, R.vp[c].data(), hw.data(), h1, h2);
    for (size_t i = 0; i < N1; ++i) gp[c][i] += bf[c][i] - hw[i];
    cggo(R, gp[c].data(), h1, h2, R.mask[c].data(), R.vtol, dv.data());
    for (size_t i = 0; i < N1; ++i) R.vp[c][i] += dv[i];
  }
  // incomprp
  vec rhs(N2), xs(N2);
  { const double* u3[3] = {R.vp[0].data(), R.vp[1].data(), R.vp[2].data()}; opdiv(R, u3, rhs.data(), -1.0); }
  ortho(R, rhs.data());
  pressure_projected(R, rhs.data(), R.ptol, xs.data());
  if (!(R.variant & 1)) R.prlag = R.prp;
  if (!(R.variant & 2)) R.prp = pext;
  for (size_t i = 0; i < N2; ++i) R.prp[i] += (bd[0] / dt) * xs[i];
  { vec w[3]; double* w3[3]; for (int c = 0; c < d; ++c) { w[c].resize(N1); w3[c] = w[c].data(); }
    opgradt(R, xs.data(), w3);
    for (int c = 0; c < d; ++c) { dssum(R, w3[c]); for (size_t i = 0; i < N1; ++i) R.vp[c][i] += w[c][i] * R.mask[c][i] * R.binvm1[i] / rho; } }
  if (R.ifheat) {
    const double h1t = R.cond, h2t = R.rhocp * bd[0] / dt;
    for (size_t i = 0; i < N1; ++i) R.tp[i] *= R.mask[3][i];
    axhelm(R, R.tp.data(), hw.data(), h1t, h2t);
    for (size_t i = 0; i < N1; ++i) bq[i] -= hw[i];
    cggo(R, bq.data(), h1t, h2t, R.mask[3].data(), R.ttol, dv.data());
    for (size_t i = 0; i < N1; ++i) R.tp[i] += dv[i];
  }
  if (R.fw > 0) { for (int c = 0; c < d; ++c) filter(R, R.vp[c].data()); if (R.ifheat) filter(R, R.tp.data()); }
}

static void transpose(const vec& M, int r, int c, vec& T) { T.resize(M.size()); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) T[(size_t)j * r + i] = M[(size_t)i * c + j]; }

}  // namespace

extern "C" {

struct nekref_desc {
  int32_t ndim, n, m; int64_t E, nglob;
  const double *D, *I12, *D12, *I1d, *Dd, *F1d;
  const double *G, *bm1, *binvm1, *vmult, *bm2, *rxw2, *rxd, *mask0, *mask1, *mask2, *mask3, *diagA;
  const int64_t* gidx;
  int32_t has_outflow; double volvm1, volvm2;
  // preconditioner (NULL S = mass-scaled identity)
  const double *S, *dinv, *wt, *A0inv, *shape; const int64_t* vertex; int64_t nv;
};

void* nekref_create(const nekref_desc* d) {
  Ref* R = new Ref();
  R->d = d->ndim; R->n = d->n; R->m = d->m; R->q = d->n - 2; R->E = d->E; R->nglob = d->nglob;
  const int n = R->n, m = R->m, q = R->q, dd = R->d;
  R->nz1 = dd == 3 ? n : 1; R->nz2 = dd == 3 ? q : 1; R->nzd = dd == 3 ? m : 1;
  R->np1 = (size_t)n * n * R->nz1; R->np2 = (size_t)q * q * R->nz2; R->npd = (size_t)m * m * R->nzd; R->N1 = R->np1 * R->E; R->N2 = R->np2 * R->E; R->ng = dd == 3 ? 6 : 3;
  R->D.assign(d->D, d->D + n * n); transpose(R->D, n, n, R->Dt);
  R->I12.assign(d->I12, d->I12 + q * n); transpose(R->I12, q, n, R->I12t);
  R->D12.assign(d->D12, d->D12 + q * n); transpose(R->D12, q, n, R->D12t);
  R->I1d.assign(d->I1d, d->I1d + m * n); transpose(R->I1d, m, n, R->I1dt);
  R->Dd.assign(d->Dd, d->Dd + m * m); R->F1d.assign(d->F1d, d->F1d + n * n);
  R->G.assign(d->G, d->G + (size_t)R->ng * R->N1); R->bm1.assign(d->bm1, d->bm1 + R->N1); R->binvm1.assign(d->binvm1, d->binvm1 + R->N1);
  R->vmult.assign(d->vmult, d->vmult + R->N1); R->bm2.assign(d->bm2, d->bm2 + R->N2); R->diagA.assign(d->diagA, d->diagA + R->N1);
  R->rxw2.assign(d->rxw2, d->rxw2 + (size_t)dd * dd * R->N2); R->rxd.assign(d->rxd, d->rxd + (size_t)dd * dd * R->npd * R->E);
  const double* mk[4] = {d->mask0, d->mask1, d->mask2, d->mask3};
  for (int k = 0; k < 4; ++k) if (mk[k]) R->mask[k].assign(mk[k], mk[k] + R->N1);
  R->has_outflow = d->has_outflow; R->volvm1 = d->volvm1; R->volvm2 = d->volvm2;
  // CSR global -> local copies
  R->gs_off.assign(R->nglob + 1, 0);
  for (size_t i = 0; i < R->N1; ++i) ++R->gs_off[d->gidx[i] + 1];
  for (int64_t g = 0; g < R->nglob; ++g) R->gs_off[g + 1] += R->gs_off[g];
  R->gs_idx.resize(R->N1); { std::vector<int64_t> pos(R->gs_off.begin(), R->gs_off.end() - 1); for (size_t i = 0; i < R->N1; ++i) R->gs_idx[pos[d->gidx[i]]++] = (int64_t)i; }
  if (d->S) {
    R->have_pre = 1; R->S.assign(d->S, d->S + (size_t)R->E * dd * n * n); R->St.resize(R->S.size());
    for (int64_t e = 0; e < R->E * dd; ++e) for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) R->St[(size_t)e * n * n + j * n + i] = R->S[(size_t)e * n * n + i * n + j];
    R->dinv.assign(d->dinv, d->dinv + R->N1); R->wt.assign(d->wt, d->wt + R->N2);
    R->nv = d->A0inv ? d->nv : 0;
    if (R->nv) { R->A0inv.assign(d->A0inv, d->A0inv + (size_t)R->nv * R->nv); R->shape.assign(d->shape, d->shape + (size_t)(1 << dd) * R->np2); R->vertex.assign(d->vertex, d->vertex + (size_t)R->E * (1 << dd)); }
  }
  for (int c = 0; c < 3; ++c) { R->U[c].assign(R->N1, 0.0); R->vp[c].assign(R->N1, 0.0); R->exx1[c].assign(R->N1, 0.0); R->exx2[c].assign(R->N1, 0.0); R->vlag[0][c].assign(R->N1, 0.0); R->vlag[1][c].assign(R->N1, 0.0); }
  R->T.assign(R->N1, 0.0); R->tp.assign(R->N1, 0.0); R->prp.assign(R->N2, 0.0); R->prlag.assign(R->N2, 0.0); R->tlag[0].assign(R->N1, 0.0); R->tlag[1].assign(R->N1, 0.0); R->vg1.assign(R->N1, 0.0); R->vg2.assign(R->N1, 0.0);
  return R;
}
void nekref_destroy(void* h) { delete (Ref*)h; }
void nekref_set_params(void* h, double visc, double rho, int torder, double vtol, double ptol, int ifheat, double cond, double rhocp, double ttol, const double* buoy,
                       double fw, int cg_maxit, int gm_maxit, int lgmres, int variant) {
  Ref* R = (Ref*)h; R->visc = visc; R->rho = rho; R->torder = torder; R->vtol = vtol; R->ptol = ptol; R->ifheat = ifheat; R->cond = cond; R->rhocp = rhocp; R->ttol = ttol;
  for (int c = 0; c < 3; ++c) R->buoy[c] = buoy[c]; R->fw = fw; R->cg_maxit = cg_maxit; R->gm_maxit = gm_maxit; R->lgmres = lgmres; R->variant = variant;
}
void nekref_set_proj(void* h, int pr_proj) { ((Ref*)h)->pr_proj = pr_proj; }
void nekref_set_mode(void* h, double dt, int adjoint, int nonlinear) { Ref* R = (Ref*)h; R->dt = dt; R->adjoint = adjoint; R->nonlinear = nonlinear; }
void nekref_set_base(void* h, const double* ux, const double* uy, const double* uz, const double* T) {
  Ref* R = (Ref*)h; const double* u[3] = {ux, uy, uz};
  for (int c = 0; c < R->d; ++c) R->U[c].assign(u[c], u[c] + R->N1);
  if (T) R->T.assign(T, T + R->N1);
}
void nekref_set_forcing(void* h, const double* fx, const double* fy, const double* fz) {
  Ref* R = (Ref*)h; const double* f[3] = {fx, fy, fz}; R->has_forcing = fx != nullptr;
  for (int c = 0; c < R->d && fx; ++c) R->forcing[c].assign(f[c], f[c] + R->N1);
}
void nekref_set_state(void* h, const double* vx, const double* vy, const double* vz, const double* pr, const double* tp) {
  Ref* R = (Ref*)h; const double* v[3] = {vx, vy, vz};
  for (int c = 0; c < R->d; ++c) R->vp[c].assign(v[c], v[c] + R->N1);
  R->prp.assign(pr, pr + R->N2); if (tp) R->tp.assign(tp, tp + R->N1);
}
void nekref_get_state(void* h, double* vx, double* vy, double* vz, double* pr, double* tp) {
  Ref* R = (Ref*)h; double* v[3] = {vx, vy, vz};
  for (int c = 0; c < R->d; ++c) std::copy(R->vp[c].begin(), R->vp[c].end(), v[c]);
  std::copy(R->prp.begin(), R->prp.end(), pr); if (tp) std::copy(R->tp.begin(), R->tp.end(), tp);
}
void nekref_reset_history(void* h) {
  Ref* R = (Ref*)h;
  for (int c = 0; c < 3; ++c) { std::fill(R->exx1[c].begin(), R->exx1[c].end(), 0.0); std::fill(R->exx2[c].begin(), R->exx2[c].end(), 0.0); for (int l = 0; l < 2; ++l) std::fill(R->vlag[l][c].begin(), R->vlag[l][c].end(), 0.0); }
  std::fill(R->prlag.begin(), R->prlag.end(), 0.0); for (int l = 0; l < 2; ++l) std::fill(R->tlag[l].begin(), R->tlag[l].end(), 0.0);
  std::fill(R->vg1.begin(), R->vg1.end(), 0.0); std::fill(R->vg2.begin(), R->vg2.end(), 0.0);
  R->projX.clear(); R->projEX.clear();
}
void nekref_advance(void* h, int istep) { advance(*(Ref*)h, istep); }
void nekref_counters(void* h, int64_t* cg, int64_t* gm) { Ref* R = (Ref*)h; *cg = R->cg_iters; *gm = R->gm_iters; }
int nekref_threads() { return omp_get_max_threads(); }
void nekref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }     // torchrun exports OMP_NUM_THREADS=1 to its workers
// operator-level hooks (validated against the numpy oracle in tests/test_oracle_cpp.py)
void nekref_axhelm(void* h, const double* u, double h1, double h2, double* w) { axhelm(*(Ref*)h, u, w, h1, h2); }
void nekref_dssum(void* h, double* u) { dssum(*(Ref*)h, u); }
void nekref_opdiv(void* h, const double* ux, const double* uy, const double* uz, double* p) { const double* u[3] = {ux, uy, uz}; opdiv(*(Ref*)h, u, p, 1.0); }
void nekref_opgradt(void* h, const double* p, double* wx, double* wy, double* wz) { double* w[3] = {wx, wy, wz}; opgradt(*(Ref*)h, p, w); }
void nekref_convect(void* h, const double* u, const double* cx, const double* cy, const double* cz, double* out) {
  Ref* R = (Ref*)h; const double* u4[4] = {u, nullptr, nullptr, nullptr}; const double* C[3] = {cx, cy, cz}; double* o4[4] = {out, nullptr, nullptr, nullptr};
  std::fill(out, out + R->N1, 0.0); convect(*R, 1, u4, C, o4, 1.0);
}
void nekref_convect_adj(void* h, const double* const* U, const double* const* c, double* const* out) {
  Ref* R = (Ref*)h; for (int i = 0; i < R->d; ++i) std::fill(out[i], out[i] + R->N1, 0.0);
  const double* U3[3] = {U[0], U[1], R->d == 3 ? U[2] : nullptr}; const double* c3[3] = {c[0], c[1], R->d == 3 ? c[2] : nullptr}; double* o3[3] = {out[0], out[1], R->d == 3 ? out[2] : nullptr};
  convect_adj(*R, U3, c3, o3, 1.0);
}
void nekref_cdabdtp(void* h, const double* p, double* ep) { Ref* R = (Ref*)h; vec w[3]; for (int c = 0; c < R->d; ++c) w[c].resize(R->N1); apply_E(*R, p, ep, w); }
void nekref_precond(void* h, const double* r, double* z) { precond(*(Ref*)h, r, z); }
int nekref_helmholtz(void* h, const double* f, double h1, double h2, int comp, double tol, double* x) { Ref* R = (Ref*)h; vec ff(f, f + R->N1); return cggo(*R, ff.data(), h1, h2, R->mask[comp].data(), tol, x); }
int nekref_pressure(void* h, const double* rhs, double tol, double* x) { return uzawa_gmres(*(Ref*)h, rhs, tol, x); }

}  // extern "C"
