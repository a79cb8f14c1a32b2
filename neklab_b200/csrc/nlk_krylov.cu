// Device-resident Krylov layer: block Gram-Schmidt kernels (K13) and the LightKrylov drivers neklab calls
// (eigs = Krylov-Schur Arnoldi, svds = Golub-Kahan, gmres) -- SURVEY.md App. B; reference call sites
// src/neklab_analysis.f90:80-81 (eigs), :136 (svds), :191-193 (newton/gmres).  LightKrylov is un-vendored
// (LightKrylov_setup.sh:55-57); algorithms restated.  Small dense algebra runs on the host (nlk_dense.cpp).
#include "nlk_ctx.hpp"
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>

using namespace nlk;

namespace nlk {

static const int KB = 8;
struct BasisPtrs { const double* x[KB][4]; double c[KB]; };

// h[j] = sum_f sum_i X_j.f[i] * y.f[i] * bm1[i]  for up to KB basis vectors per launch (y and bm1 streamed once)
__global__ void __launch_bounds__(256)
k_basis_dot(BasisPtrs P, int nb, CPtr4 y, int nf, const double* __restrict__ bm1, size_t n, double* h, Reducer red) {
  double v[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) v[j] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double b = bm1[i];
    for (int f = 0; f < nf; ++f) {
      double yb = y.p[f][i] * b;
#pragma unroll
      for (int j = 0; j < KB; ++j) if (j < nb) v[j] += P.x[j][f][i] * yb;
    }
  }
  // reuse the generic deterministic reducer through a local copy (KB values)
  __shared__ double s_part[KB][32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < KB; ++k) { double t = v[k]; for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o); if (lane == 0) s_part[k][wid] = t; }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < KB; ++k) { double t = lane < nw ? s_part[k][lane] : 0.0; for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o); if (lane == 0) red.partial[(size_t)k * red.maxblocks + blockIdx.x] = t; }
  }
  if (threadIdx.x == 0) { __threadfence(); unsigned int tk = atomicAdd(red.counter, 1u); s_last = (tk == gridDim.x - 1); }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int k = 0; k < KB; ++k) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) acc += __ldcg(&red.partial[(size_t)k * red.maxblocks + b]);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __syncthreads();
    if (lane == 0) s_part[k][wid] = acc;
    __syncthreads();
    if (wid == 0) { double t = lane < nw ? s_part[k][lane] : 0.0; for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o); if (lane == 0 && k < nb) h[k] = t; }
  }
  if (threadIdx.x == 0) *red.counter = 0u;
}

// y[i] += sum_j c_j * X_j[i]   (one field; up to KB vectors per launch)
struct AxpyPtrs { const double* x[KB]; double c[KB]; };
__global__ void k_basis_axpy(double* __restrict__ y, AxpyPtrs P, int nb, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double s = y[i];
#pragma unroll
    for (int j = 0; j < KB; ++j) if (j < nb) s += P.c[j] * P.x[j][i];
    y[i] = s;
  }
}

static int basis_innerprod(nlk_ctx* c, nlk_vec* const* X, int k, const nlk_vec* y, double* h) {
  const DevMesh& dm = c->dm;
  if (k > 480) { set_error("basis_innerprod: k too large"); return 1; }
  CPtr4 Y{}; int nf = 0;
  for (int f = 0; f < dm.ndim; ++f) Y.p[nf++] = y->v[f];
  if (c->prm.ifheat) Y.p[nf++] = y->theta;
  int grid = std::min((int)((dm.N1 + 255) / 256), 592);
  for (int k0 = 0; k0 < k; k0 += KB) {
    BasisPtrs P{}; int nb = std::min(KB, k - k0);
    for (int j = 0; j < nb; ++j) { int g = 0; for (int f = 0; f < dm.ndim; ++f) P.x[j][g++] = X[k0 + j]->v[f]; if (c->prm.ifheat) P.x[j][g++] = X[k0 + j]->theta; }
    k_basis_dot<<<grid, 256, 0, c->st>>>(P, nb, Y, nf, dm.bm1, dm.N1, c->d_red + k0, c->red); ++g_launches;
  }
  if (ctx_allreduce(c, c->d_red, k, false)) return 1;
  if (ctx_read_scalars(c, k)) return 1;
  for (int j = 0; j < k; ++j) h[j] = c->h_red[j];
  return 0;
}

static void axpy_field(nlk_ctx* c, double* y, nlk_vec* const* X, int k, const double* coef, size_t n, int which /*0..2 vel, 3 pr, 4 theta*/, int slot = -1) {
  int grid = std::min((int)((n + 255) / 256), 148 * 8);
  for (int k0 = 0; k0 < k; k0 += KB) {
    AxpyPtrs P{}; int nb = std::min(KB, k - k0);
    for (int j = 0; j < nb; ++j) {
      const nlk_vec* x = X[k0 + j];
      if (slot >= 0 && x->nrst > slot) P.x[j] = which < 3 ? x->rv[slot][which] : (which == 3 ? x->rpr[slot] : x->rth[slot]);
      else P.x[j] = which < 3 ? x->v[which] : (which == 3 ? x->pr : x->theta);
      P.c[j] = coef[k0 + j];
    }
    k_basis_axpy<<<grid, 256, 0, c->st>>>(y, P, nb, n); ++g_launches;
  }
}

// y += X c on every field (pressure included, like nek_daxpby); the rst slots of y receive the same combination
// of the X's CURRENT fields (quirk of real_vectors.f90:186-200 replicated through LightKrylov's axpby loop).
static int basis_axpy(nlk_ctx* c, nlk_vec* y, nlk_vec* const* X, int k, const double* coef) {
  const DevMesh& dm = c->dm;
  for (int f = 0; f < dm.ndim; ++f) axpy_field(c, y->v[f], X, k, coef, dm.N1, f);
  axpy_field(c, y->pr, X, k, coef, dm.N2, 3);
  if (c->prm.ifheat) axpy_field(c, y->theta, X, k, coef, dm.N1, 4);
  const int sl = c->prm.rst_mode == 1 ? 1 : 0;    // rst_mode 1 (non-reference): combine the X's rst fields instead
  for (int s = 0; s < y->nrst; ++s) {
    for (int f = 0; f < dm.ndim; ++f) axpy_field(c, y->rv[s][f], X, k, coef, dm.N1, f, sl ? s : -1);
    axpy_field(c, y->rpr[s], X, k, coef, dm.N2, 3, sl ? s : -1);
    if (c->prm.ifheat) axpy_field(c, y->rth[s], X, k, coef, dm.N1, 4, sl ? s : -1);
  }
  return 0;
}

// LightKrylov double_gram_schmidt_step: h = X^T B y; y -= X h; twice; returns h1+h2
static int basis_dgs(nlk_ctx* c, nlk_vec* y, nlk_vec* const* X, int k, double* h) {
  std::vector<double> h1(k), h2(k), neg(k);
  if (basis_innerprod(c, X, k, y, h1.data())) return 1;
  for (int j = 0; j < k; ++j) neg[j] = -h1[j];
  if (basis_axpy(c, y, X, k, neg.data())) return 1;
  if (basis_innerprod(c, X, k, y, h2.data())) return 1;
  for (int j = 0; j < k; ++j) neg[j] = -h2[j];
  if (basis_axpy(c, y, X, k, neg.data())) return 1;
  for (int j = 0; j < k; ++j) h[j] = h1[j] + h2[j];
  return 0;
}

static const double ATOL_DP = 1e-15;

// owns the work vectors of a Krylov driver: every early `return 1` releases them (a retrying caller -- Newton -- must not grow HBM)
struct VecGuard {
  std::vector<std::vector<nlk_vec*>*> lists; std::vector<nlk_vec**> singles;
  ~VecGuard() {
    for (auto* l : lists) for (auto& x : *l) { nlk_vec_destroy(x); x = nullptr; }
    for (auto** p : singles) { nlk_vec_destroy(*p); *p = nullptr; }
  }
};

}  // namespace nlk

extern "C" {

int nlk_basis_innerprod(nlk_vec* const* X, int32_t k, const nlk_vec* y, double* h) { return k > 0 ? basis_innerprod(y->c, X, k, y, h) : 0; }
int nlk_basis_axpy(nlk_vec* y, nlk_vec* const* X, int32_t k, const double* coef) { return k > 0 ? basis_axpy(y->c, y, X, k, coef) : 0; }
int nlk_basis_dgs(nlk_vec* y, nlk_vec* const* X, int32_t k, double* h, double* norm_out) {
  if (k > 0 && basis_dgs(y->c, y, X, k, h)) return 1;
  if (norm_out) return nlk_vec_norm(y, norm_out);
  return 0;
}

// Krylov-Schur eigensolver (LightKrylov eigs_rdp as driven by linear_stability_analysis_fixed_point,
// src/neklab_analysis.f90:38-105): residual_i = |H(k+1,k)| |y_i(k)|, converged when count(res < tol) >= nev,
// restart keeps Ritz values with |lambda| above the median.  Returns the nev leading Ritz values (|.| descending).
int nlk_eigs(nlk_op* op, int32_t nev, int32_t kdim, double tol, int32_t transpose, const nlk_vec* x0, double* lam_re, double* lam_im,
             double* resid, nlk_vec** eigvecs, int32_t* niter_out, nlk_eigs_cb cb, void* user, int32_t* info) {
  nlk_ctx* c = op->c;
  if (nev < 1 || kdim < nev + 1) { set_error("eigs: need kdim > nev >= 1"); return 1; }
  if (tol <= 0) tol = std::sqrt(ATOL_DP);
  std::vector<nlk_vec*> X(kdim + 1, nullptr);
  VecGuard guard; guard.lists.push_back(&X);
  if (nlk_vec_create(c, &X[0])) return 1;                 // the basis grows with the iteration: at 100k elements a vector is 1.4 GB
  if (x0) { if (nlk_vec_copy(X[0], x0)) return 1; } else { if (nlk_vec_rand(X[0], 0, 12345)) return 1; }
  double nr; if (nlk_vec_norm(X[0], &nr)) return 1;
  if (nr == 0) { set_error("eigs: zero starting vector"); return 1; }
  if (nlk_vec_scal(X[0], 1.0 / nr)) return 1;
  const int ldh = kdim;                                   // H is (kdim+1) x kdim, row-major
  std::vector<double> H((size_t)(kdim + 1) * kdim, 0.0);
  std::vector<double> wr(kdim), wi(kdim), VR((size_t)kdim * kdim), res(kdim), Hk;
  int kstart = 0, conv = 0, niter = 0, kcur = 0;
  const int maxrestart = 20;
  *info = 0;
  auto ritz = [&](int k) -> int {
    Hk.assign((size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) Hk[(size_t)i * k + j] = H[(size_t)i * ldh + j];
    if (dense_eig(k, Hk.data(), wr.data(), wi.data(), VR.data())) return 1;
    double beta = H[(size_t)k * ldh + (k - 1)];
    for (int j = 0; j < k; ++j) {
      double y;
      if (wi[j] == 0.0) y = std::fabs(VR[(size_t)(k - 1) * k + j]);
      else if (wi[j] > 0) y = std::hypot(VR[(size_t)(k - 1) * k + j], VR[(size_t)(k - 1) * k + j + 1]);
      else y = std::hypot(VR[(size_t)(k - 1) * k + j - 1], VR[(size_t)(k - 1) * k + j]);
      res[j] = std::fabs(beta) * y;
    }
    return 0;
  };
  for (int outer = 0; outer <= maxrestart && conv < nev; ++outer) {
    for (int k = kstart; k < kdim; ++k) {
      // Arnoldi step: X[k+1] = A X[k]; DGS; normalise
      if (!X[k + 1] && nlk_vec_create(c, &X[k + 1])) return 1;
      if (exptA_apply(op, X[k], X[k + 1], transpose != 0)) return 1;
      // rst fields are read only from the vector a matvec starts from: with the reference's axpby (rst slots receive the
      // CURRENT fields of the other vector, real_vectors.f90:186-200) nothing ever reads X[k]'s again -- free them, so that a
      // basis vector costs d fields + pressure instead of three times that
      if (c->prm.rst_mode != 1) vec_release_rst(X[k]);
      std::vector<double> h(k + 1);
      if (basis_dgs(c, X[k + 1], X.data(), k + 1, h.data())) return 1;
      for (int i = 0; i <= k; ++i) H[(size_t)i * ldh + k] = h[i];
      double beta; if (nlk_vec_norm(X[k + 1], &beta)) return 1;
      H[(size_t)(k + 1) * ldh + k] = beta;
      if (beta > ATOL_DP) { if (nlk_vec_scal(X[k + 1], 1.0 / beta)) return 1; }
      kcur = k + 1; ++niter;
      if (ritz(kcur)) return 1;
      conv = 0; for (int j = 0; j < kcur; ++j) if (res[j] < tol) ++conv;
      if (cb) cb(niter, kcur, wr.data(), wi.data(), res.data(), user);
      if (conv >= nev) break;
    }
    if (conv >= nev) break;
    if (outer == maxrestart) { *info = 1; break; }
    // ---- Krylov-Schur restart: orthonormal basis Q of the invariant subspace of the selected Ritz values
    const int k = kdim;
    std::vector<double> mod(k); for (int j = 0; j < k; ++j) mod[j] = std::hypot(wr[j], wi[j]);
    std::vector<double> srt(mod); std::sort(srt.begin(), srt.end());
    double median = (k % 2) ? srt[k / 2] : 0.5 * (srt[k / 2 - 1] + srt[k / 2]);
    std::vector<int> sel;
    for (int j = 0; j < k; ++j) if (mod[j] > median) sel.push_back(j);
    // keep complex pairs together
    std::vector<char> in(k, 0); for (int j : sel) in[j] = 1;
    for (int j = 0; j < k; ++j) if (wi[j] > 0 && j + 1 < k && (in[j] != in[j + 1])) in[j] = in[j + 1] = 1;
    std::vector<std::vector<double>> Q;
    for (int j = 0; j < k; ++j) if (in[j]) {
      std::vector<double> col(k);
      for (int i = 0; i < k; ++i) col[i] = VR[(size_t)i * k + j];     // real vec, or Re / Im parts of a pair
      for (int pass = 0; pass < 2; ++pass) for (auto& qv : Q) { double s = 0; for (int i = 0; i < k; ++i) s += qv[i] * col[i]; for (int i = 0; i < k; ++i) col[i] -= s * qv[i]; }
      double nn = 0; for (int i = 0; i < k; ++i) nn += col[i] * col[i]; nn = std::sqrt(nn);
      if (nn < 1e-10) continue;
      for (int i = 0; i < k; ++i) col[i] /= nn;
      Q.push_back(col);
    }
    const int p = (int)Q.size();
    if (p == 0 || p >= k) { *info = 2; break; }
    // new Rayleigh quotient T = Q^T H Q, coupling row b = H(k+1,k) * Q(k,:)
    std::vector<double> HQ((size_t)k * p, 0.0), T((size_t)p * p, 0.0), brow(p);
    for (int i = 0; i < k; ++i) for (int l = 0; l < k; ++l) { double hv = H[(size_t)i * ldh + l]; if (hv != 0) for (int j = 0; j < p; ++j) HQ[(size_t)i * p + j] += hv * Q[j][l]; }
    for (int a = 0; a < p; ++a) for (int j = 0; j < p; ++j) { double s = 0; for (int i = 0; i < k; ++i) s += Q[a][i] * HQ[(size_t)i * p + j]; T[(size_t)a * p + j] = s; }
    for (int j = 0; j < p; ++j) brow[j] = H[(size_t)k * ldh + (k - 1)] * Q[j][k - 1];
    // X[:p] <- X[:k] Q (LightKrylov linear_combination = zero + axpby, so rst follows the axpby quirk: nrst = 0 after zero)
    std::vector<nlk_vec*> NX(p, nullptr);
    VecGuard gnx; gnx.lists.push_back(&NX);
    for (int j = 0; j < p; ++j) { if (nlk_vec_create(c, &NX[j])) return 1; if (nlk_vec_zero(NX[j])) return 1; if (basis_axpy(c, NX[j], X.data(), k, Q[j].data())) return 1; }
    for (int j = 0; j < p; ++j) { std::swap(X[j], NX[j]); }
    std::swap(X[p], X[k]);
    std::fill(H.begin(), H.end(), 0.0);
    for (int a = 0; a < p; ++a) for (int j = 0; j < p; ++j) H[(size_t)a * ldh + j] = T[(size_t)a * p + j];
    for (int j = 0; j < p; ++j) H[(size_t)p * ldh + j] = brow[j];
    kstart = p;
  }
  // sort by modulus descending, report nev
  std::vector<int> ord(kcur); for (int j = 0; j < kcur; ++j) ord[j] = j;
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return std::hypot(wr[a], wi[a]) > std::hypot(wr[b], wi[b]); });
  for (int t = 0; t < nev && t < kcur; ++t) {
    int j = ord[t];
    lam_re[t] = wr[j]; lam_im[t] = wi[j]; resid[t] = res[j];
    if (eigvecs) {
      std::vector<double> cre(kcur), cim(kcur, 0.0);
      if (wi[j] == 0.0) for (int i = 0; i < kcur; ++i) cre[i] = VR[(size_t)i * kcur + j];
      else { int jr = wi[j] > 0 ? j : j - 1; double sg = wi[j] > 0 ? 1.0 : -1.0; for (int i = 0; i < kcur; ++i) { cre[i] = VR[(size_t)i * kcur + jr]; cim[i] = sg * VR[(size_t)i * kcur + jr + 1]; } }
      if (nlk_vec_zero(eigvecs[2 * t]) || nlk_vec_zero(eigvecs[2 * t + 1])) return 1;
      if (basis_axpy(c, eigvecs[2 * t], X.data(), kcur, cre.data()) || basis_axpy(c, eigvecs[2 * t + 1], X.data(), kcur, cim.data())) return 1;
    }
  }
  if (niter_out) *niter_out = niter;
  return 0;
}

// Golub-Kahan bidiagonalisation (LightKrylov svds as driven by transient_growth_analysis_fixed_point,
// src/neklab_analysis.f90:107-156): per step one matvec + one rmatvec + DGS against U and V.
int nlk_svds(nlk_op* op, int32_t nsv, int32_t kdim, double tol, const nlk_vec* x0, double* sigma, double* resid, nlk_vec** Uout, nlk_vec** Vout,
             int32_t* niter_out, int32_t* info) {
  nlk_ctx* c = op->c;
  if (tol <= 0) tol = std::sqrt(ATOL_DP);
  std::vector<nlk_vec*> U(kdim, nullptr), V(kdim + 1, nullptr);
  VecGuard guard; guard.lists.push_back(&U); guard.lists.push_back(&V);
  for (auto& x : U) if (nlk_vec_create(c, &x)) return 1;
  for (auto& x : V) if (nlk_vec_create(c, &x)) return 1;
  if (x0) { if (nlk_vec_copy(V[0], x0)) return 1; } else { if (nlk_vec_rand(V[0], 0, 12345)) return 1; }
  double nr; if (nlk_vec_norm(V[0], &nr)) return 1; if (nlk_vec_scal(V[0], 1.0 / nr)) return 1;
  std::vector<double> alpha(kdim, 0.0), beta(kdim + 1, 0.0), sv, res;
  int k = 0, conv = 0; *info = 0;
  for (k = 0; k < kdim; ++k) {
    if (exptA_apply(op, V[k], U[k], false)) return 1;
    if (k > 0) { std::vector<double> h(k); if (basis_dgs(c, U[k], U.data(), k, h.data())) return 1; }
    if (nlk_vec_norm(U[k], &alpha[k])) return 1; if (nlk_vec_scal(U[k], 1.0 / alpha[k])) return 1;
    if (exptA_apply(op, U[k], V[k + 1], true)) return 1;
    { std::vector<double> h(k + 1); if (basis_dgs(c, V[k + 1], V.data(), k + 1, h.data())) return 1; }
    if (nlk_vec_norm(V[k + 1], &beta[k + 1])) return 1;
    if (beta[k + 1] > ATOL_DP) { if (nlk_vec_scal(V[k + 1], 1.0 / beta[k + 1])) return 1; }
    // B_k (upper bidiagonal: diag alpha, superdiag beta): singular values via eig of B^T B (symmetric tridiagonal)
    const int n = k + 1;
    std::vector<double> BtB((size_t)n * n, 0.0), w(n), Z((size_t)n * n);
    // A V_k = U_k B_k with B(i,i)=alpha_i, B(i,i+1)=beta_{i+1};  B^T B tridiagonal
    for (int i = 0; i < n; ++i) { BtB[(size_t)i * n + i] = alpha[i] * alpha[i] + (i > 0 ? beta[i] * beta[i] : 0.0); if (i + 1 < n) { BtB[(size_t)i * n + i + 1] = alpha[i] * beta[i + 1]; BtB[(size_t)(i + 1) * n + i] = alpha[i] * beta[i + 1]; } }
    sym_eig_jacobi(n, BtB.data(), w.data(), Z.data());
    sv.assign(n, 0.0); res.assign(n, 0.0);
    for (int j = 0; j < n; ++j) {                         // descending
      int src = n - 1 - j; sv[j] = std::sqrt(std::max(w[src], 0.0));
      // residual |beta_{k+1} * (left singular vector)_last| ; u = B v / sigma, last component = alpha_k v_k / sigma
      double vlast = Z[(size_t)(n - 1) * n + src];
      res[j] = sv[j] > 0 ? std::fabs(beta[k + 1] * alpha[k] * vlast / sv[j]) : 0.0;
    }
    conv = 0; for (int j = 0; j < n; ++j) if (res[j] < tol) ++conv;
    if (getenv("NLK_DEBUG_SVDS")) { printf("svds k=%d alpha=%.6e beta=%.6e :", k, alpha[k], beta[k + 1]); for (int j = 0; j < n; ++j) printf(" (%.6e, %.2e)", sv[j], res[j]); printf("\n"); }
    if (n >= nsv && conv >= nsv) { ++k; break; }
  }
  const int n = std::min(k, kdim);
  for (int j = 0; j < nsv && j < (int)sv.size(); ++j) { sigma[j] = sv[j]; resid[j] = res[j]; }
  // singular vectors: right v_j = V_n z_j (z_j eigenvector of B^T B), left u_j = U_n (B z_j) / sigma_j
  if ((Uout || Vout) && n > 0) {
    std::vector<double> BtB((size_t)n * n, 0.0), w(n), Z((size_t)n * n);
    for (int i = 0; i < n; ++i) { BtB[(size_t)i * n + i] = alpha[i] * alpha[i] + (i > 0 ? beta[i] * beta[i] : 0.0); if (i + 1 < n) { BtB[(size_t)i * n + i + 1] = alpha[i] * beta[i + 1]; BtB[(size_t)(i + 1) * n + i] = alpha[i] * beta[i + 1]; } }
    sym_eig_jacobi(n, BtB.data(), w.data(), Z.data());
    for (int j = 0; j < nsv && j < n; ++j) {
      const int src = n - 1 - j; const double sg = std::sqrt(std::max(w[src], 0.0));
      std::vector<double> zc(n), uc(n);
      for (int i = 0; i < n; ++i) zc[i] = Z[(size_t)i * n + src];
      for (int i = 0; i < n; ++i) uc[i] = (alpha[i] * zc[i] + (i + 1 < n ? beta[i + 1] * zc[i + 1] : 0.0)) / (sg > 0 ? sg : 1.0);
      if (Vout) { if (nlk_vec_zero(Vout[j]) || basis_axpy(c, Vout[j], V.data(), n, zc.data())) return 1; }
      if (Uout) { if (nlk_vec_zero(Uout[j]) || basis_axpy(c, Uout[j], U.data(), n, uc.data())) return 1; }
    }
  }
  if (niter_out) *niter_out = n;
  if (conv < nsv) *info = 1;
  return 0;
}

// Restarted GMRES on A (minus_identity: on A - I, the fixed-point Jacobian of src/systems/fixed_point.f90:42-96)
int nlk_gmres(nlk_op* op, int32_t minus_identity, const nlk_vec* b, nlk_vec* x, int32_t kdim, double atol, double rtol, int32_t maxiter,
              int32_t transpose, int32_t* info) {
  nlk_ctx* c = op->c;
  std::vector<nlk_vec*> V(kdim + 1, nullptr);
  nlk_vec* r = nullptr;
  VecGuard guard; guard.lists.push_back(&V); guard.singles.push_back(&r);
  if (nlk_vec_create(c, &V[0]) || nlk_vec_create(c, &r)) return 1;      // the basis grows with the iteration (kdim 64 in the resolvent)
  auto apply = [&](const nlk_vec* in, nlk_vec* out) -> int {
    if (exptA_apply(op, in, out, transpose != 0)) return 1;
    if (minus_identity) { if (nlk_vec_axpby(-1.0, in, 1.0, out)) return 1; }
    return 0;
  };
  double bnorm; if (nlk_vec_norm(b, &bnorm)) return 1;
  const double tol = atol + rtol * bnorm;
  *info = 1;
  for (int outer = 0; outer <= maxiter; ++outer) {
    double xn = 1.0;
    if (outer == 0) { if (nlk_vec_norm(x, &xn)) return 1; }
    if (xn == 0.0) { if (nlk_vec_copy(r, b)) return 1; }        // zero initial guess: r = b without a matvec (LightKrylov does the same)
    else {
      if (apply(x, r)) return 1;
      if (nlk_vec_axpby(1.0, b, -1.0, r)) return 1;             // r = b - A x
    }
    r->nrst = 0;
    double beta; if (nlk_vec_norm(r, &beta)) return 1;
    if (beta < tol) { *info = 0; break; }
    if (outer == maxiter) break;
    if (nlk_vec_copy(V[0], r)) return 1; if (nlk_vec_scal(V[0], 1.0 / beta)) return 1;
    std::vector<double> H((size_t)(kdim + 1) * kdim, 0.0), cs(kdim), sn(kdim), g(kdim + 1, 0.0), y(kdim);
    g[0] = beta; int kk = 0;
    for (int k = 0; k < kdim; ++k) {
      if (!V[k + 1] && nlk_vec_create(c, &V[k + 1])) return 1;
      if (apply(V[k], V[k + 1])) return 1;
      if (c->prm.rst_mode != 1) vec_release_rst(V[k]);            // as in nlk_eigs: only the newest vector's rst fields are ever read
      std::vector<double> h(k + 1);
      if (basis_dgs(c, V[k + 1], V.data(), k + 1, h.data())) return 1;
      double hn; if (nlk_vec_norm(V[k + 1], &hn)) return 1;
      if (hn > ATOL_DP) { if (nlk_vec_scal(V[k + 1], 1.0 / hn)) return 1; }
      for (int i = 0; i <= k; ++i) H[(size_t)i * kdim + k] = h[i];
      for (int i = 0; i < k; ++i) { double t = H[(size_t)i * kdim + k]; H[(size_t)i * kdim + k] = cs[i] * t + sn[i] * H[(size_t)(i + 1) * kdim + k]; H[(size_t)(i + 1) * kdim + k] = -sn[i] * t + cs[i] * H[(size_t)(i + 1) * kdim + k]; }
      double hkk = H[(size_t)k * kdim + k], l = std::hypot(hkk, hn);
      cs[k] = hkk / l; sn[k] = hn / l; H[(size_t)k * kdim + k] = l;
      g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
      kk = k + 1;
      if (std::fabs(g[k + 1]) < tol) break;
    }
    for (int i = kk - 1; i >= 0; --i) { double t = g[i]; for (int j = i + 1; j < kk; ++j) t -= H[(size_t)i * kdim + j] * y[j]; y[i] = t / H[(size_t)i * kdim + i]; }
    int nr = x->nrst; x->nrst = 0;                               // solution update touches current fields only
    if (basis_axpy(c, x, V.data(), kk, y.data())) return 1;
    x->nrst = nr;
  }
  return 0;
}

}  // extern "C"
