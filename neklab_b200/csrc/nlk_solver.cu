// Solver context, gather-scatter with NCCL interface exchange, Helmholtz Jacobi-PCG, PN-PN-2 pressure FGMRES with
// Schwarz + coarse preconditioner, the perturbation time step and the exptA matvec.
// Restates (un-vendored Nek5000): hmholtz/cggo (hmholtz.f), uzawa_gmres (gmres.f), cdabdtp,
// hsmg_solve structure (hsmg.f), perturbv/makefp/advabp/makextp/makebdfp/lagfieldp/cresvipp/incomprp (perturb.f),
// compute_cfl; and in-tree: setup_nek dt/nsteps rule (src/neklab_nek_setup.f90:193-224),
// exptA_matvec / compute_rst / get_rst (src/linops/exponential_propagator.f90:15-142).
#include "nlk_ctx.hpp"
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace nlk {

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
int nccl_load(Nccl& n) {
  if (n.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !n.lib; ++i) n.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!n.lib) { set_error("cannot dlopen libnccl.so.2 (needed for multi-GPU runs)"); return 1; }
#define NLK_SYM(field, name) *(void**)(&n.field) = dlsym(n.lib, name); if (!n.field) { set_error(std::string("missing NCCL symbol ") + name); return 1; }
  NLK_SYM(GetUniqueId, "ncclGetUniqueId") NLK_SYM(CommInitRank, "ncclCommInitRank") NLK_SYM(CommDestroy, "ncclCommDestroy")
  NLK_SYM(AllReduce, "ncclAllReduce") NLK_SYM(Send, "ncclSend") NLK_SYM(Recv, "ncclRecv") NLK_SYM(GroupStart, "ncclGroupStart")
  NLK_SYM(GroupEnd, "ncclGroupEnd") NLK_SYM(GetErrorString, "ncclGetErrorString")
#undef NLK_SYM
  *(void**)(&n.CommSplit) = dlsym(n.lib, "ncclCommSplit");       // optional (NCCL >= 2.18)
  return 0;
}
#define NLK_NCCL(c, call) do { int _r = (call); if (_r != 0) { set_error(std::string("NCCL error: ") + (c)->nccl.GetErrorString(_r)); return 1; } } while (0)
static const int NCCL_DOUBLE = 8, NCCL_SUM = 0, NCCL_MAX = 2;

int ctx_allreduce(nlk_ctx* c, double* d_ptr, int count, bool maxop) {
  if (c->nccl.nranks <= 1) return 0;
  void* comm = (c->st == c->st2 && c->nccl.comm2) ? c->nccl.comm2 : c->nccl.comm;      // side stream -> its own communicator
  NLK_NCCL(c, c->nccl.AllReduce(d_ptr, d_ptr, (size_t)count, NCCL_DOUBLE, maxop ? NCCL_MAX : NCCL_SUM, comm, c->st));
  return 0;
}

// full direct-stiffness sum: local segmented reduction, then pack -> ncclSend/Recv with every neighbour -> unpack-add
int ctx_gs(nlk_ctx* c, Ptr3 f, int nf) {
  launch_gs(c->dm, f, nf, c->st);
  if (c->nccl.nranks <= 1 || c->neigh.empty()) return 0;
  for (auto& nb : c->neigh) for (int k = 0; k < nf; ++k) launch_pack(f.p[k], nb.rep, nb.cnt, nb.sendbuf + (size_t)k * nb.cnt, c->st);
  NLK_NCCL(c, c->nccl.GroupStart());
  for (auto& nb : c->neigh) {
    NLK_NCCL(c, c->nccl.Send(nb.sendbuf, (size_t)nf * nb.cnt, NCCL_DOUBLE, nb.rank, c->nccl.comm, c->st));
    NLK_NCCL(c, c->nccl.Recv(nb.recvbuf, (size_t)nf * nb.cnt, NCCL_DOUBLE, nb.rank, c->nccl.comm, c->st));
  }
  NLK_NCCL(c, c->nccl.GroupEnd());
  for (auto& nb : c->neigh) for (int k = 0; k < nf; ++k) launch_unpack_add(f.p[k], nb.cp_off, nb.cp_idx, nb.cnt, nb.recvbuf + (size_t)k * nb.cnt, c->st);
  return 0;
}

// Host read-back of a few device scalars (FGMRES Hessenberg column, dots, CFL).  Default: a one-block kernel writes them straight
// into pinned host memory (UVA zero-copy), fences, and publishes a sequence number the host spins on -- no cudaMemcpyAsync +
// cudaStreamSynchronize round trip on the critical path of every FGMRES iteration (r01 verdict: ~0.13 of 0.85 ms per iteration on
// the launch-bound configs).  A stream synchronize remains the fallback (NLK_SYNC_READBACK=1, or if the spin times out).
int ctx_read_scalars(nlk_ctx* c, int count) {
  static const bool sync_mode = getenv("NLK_SYNC_READBACK") != nullptr;
  if (!sync_mode && c->h_seq) {
    const unsigned int seq = ++c->seq;
    launch_publish(c->d_red, count, c->h_red, c->h_seq, seq, c->st);
    volatile unsigned int* flag = c->h_seq;
    for (long spin = 0; *flag != seq; ++spin) {
      if ((spin & 0xfffff) == 0xfffff) {                     // every ~1M polls: has the stream died, or is it just a long kernel?
        cudaError_t q = cudaStreamQuery(c->st);
        if (q != cudaErrorNotReady) {                        // finished (flag will be visible after the sync) or failed
          NLK_CUDA(cudaStreamSynchronize(c->st)); NLK_CUDA(cudaGetLastError());
          if (*flag != seq) { set_error("scalar read-back: stream finished without publishing"); return 1; }
          break;
        }
      }
    }
    return 0;
  }
  NLK_CUDA(cudaMemcpyAsync(c->h_red, c->d_red, sizeof(double) * count, cudaMemcpyDeviceToHost, c->st));
  NLK_CUDA(cudaStreamSynchronize(c->st));
  NLK_CUDA(cudaGetLastError());               // launch-configuration errors are not sticky: surface them here
  return 0;
}

static int global_dot(nlk_ctx* c, size_t n, const double* a, const double* b, const double* w, double* d_out) {
  CPtr4 A{{a, nullptr, nullptr, nullptr}}, B{{b, nullptr, nullptr, nullptr}};
  launch_dot(n, A, B, 1, w, d_out, c->red, c->st);
  return ctx_allreduce(c, d_out, 1, false);
}

// ------------------------------------------------------------------------------------------------ FDM setup (host)
enum { BC_OVERLAP = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2 };
// 1-D generalized eigenpairs of (linear-FE stiffness, GL-quadrature mass) on the extended point set.
static void fdm_1d(const Basis& b, double lm, double ll, double lr, int bcl, int bcr, double* S /*n*n*/, double* lam /*n*/, int* nact_out) {
  const int q = b.q, n = q + 2;
  std::vector<double> pts, mass; std::vector<int> slot;
  const double g0 = 0.5 * (1.0 + b.z2[0]);
  const double gap = 0.5 * (b.z2[1] - b.z2[0]);
  if (bcl == BC_OVERLAP) { pts.push_back(-ll * (g0 + gap)); mass.push_back(0); slot.push_back(-1); pts.push_back(-ll * g0); mass.push_back(0.5 * ll * b.w2[0]); slot.push_back(0); }
  else if (bcl == BC_DIRICHLET) { pts.push_back(0.0); mass.push_back(0); slot.push_back(-1); }
  for (int i = 0; i < q; ++i) { pts.push_back(0.5 * lm * (b.z2[i] + 1.0)); mass.push_back(0.5 * lm * b.w2[i]); slot.push_back(i + 1); }
  if (bcr == BC_OVERLAP) { pts.push_back(lm + lr * g0); mass.push_back(0.5 * lr * b.w2[0]); slot.push_back(n - 1); pts.push_back(lm + lr * (g0 + gap)); mass.push_back(0); slot.push_back(-1); }
  else if (bcr == BC_DIRICHLET) { pts.push_back(lm); mass.push_back(0); slot.push_back(-1); }
  const int N = (int)pts.size();
  std::vector<double> A((size_t)N * N, 0.0);
  for (int i = 0; i + 1 < N; ++i) { double h = pts[i + 1] - pts[i]; A[i * N + i] += 1 / h; A[(i + 1) * N + i + 1] += 1 / h; A[i * N + i + 1] -= 1 / h; A[(i + 1) * N + i] -= 1 / h; }
  std::vector<int> act; for (int i = 0; i < N; ++i) if (slot[i] >= 0) act.push_back(i);
  const int na = (int)act.size();
  std::vector<double> Aa((size_t)na * na), bi(na), w(na), V((size_t)na * na);
  for (int i = 0; i < na; ++i) bi[i] = 1.0 / std::sqrt(mass[act[i]]);
  for (int i = 0; i < na; ++i) for (int j = 0; j < na; ++j) Aa[i * na + j] = A[act[i] * N + act[j]] * bi[i] * bi[j];
  sym_eig_jacobi(na, Aa.data(), w.data(), V.data());
  std::fill(S, S + n * n, 0.0); std::fill(lam, lam + n, 1.0);
  for (int jj = 0; jj < na; ++jj) { lam[jj] = w[jj]; for (int i = 0; i < na; ++i) S[slot[act[i]] * n + jj] = V[i * na + jj] * bi[i]; }
  *nact_out = na;
}

static double diag_local(const HostMesh& hm, int64_t e, int p) {
  const int n = hm.n, d = hm.ndim, np1 = hm.np1;
  const double* D = hm.b.D.data();
  int i = p % n, j = (p / n) % n, k = d == 3 ? p / (n * n) : 0;
  size_t eb = (size_t)e * np1;
  double s = 0;
  for (int l = 0; l < n; ++l) {
    s += hm.G[0][eb + (k * n + j) * n + l] * D[l * n + i] * D[l * n + i];
    s += hm.G[1][eb + (k * n + l) * n + i] * D[l * n + j] * D[l * n + j];
    if (d == 3) s += hm.G[2][eb + (l * n + j) * n + i] * D[l * n + k] * D[l * n + k];
  }
  s += 2.0 * hm.G[3][eb + p] * D[i * n + i] * D[j * n + j];
  if (d == 3) { s += 2.0 * hm.G[4][eb + p] * D[i * n + i] * D[k * n + k]; s += 2.0 * hm.G[5][eb + p] * D[j * n + j] * D[k * n + k]; }
  return s;
}

static std::vector<double> transpose(const std::vector<double>& M, int r, int cdim) {
  std::vector<double> T((size_t)r * cdim);
  for (int i = 0; i < r; ++i) for (int j = 0; j < cdim; ++j) T[(size_t)j * r + i] = M[(size_t)i * cdim + j];
  return T;
}

// Nek q_filter 1-D operator: F = V diag(sigma) V^-1 on the bubble basis phi_k = L_k - L_{k-2}
static void filter_matrix(const Basis& b, double wght, double cutoff, std::vector<double>& F) {
  const int n = b.n;
  F.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) F[i * n + i] = 1.0;
  int ncut = cutoff < 1.0 ? n - (int)(n * cutoff + 0.01) : 0;
  if (ncut <= 0 || wght <= 0) return;
  std::vector<double> V((size_t)n * n), Vi((size_t)n * n, 0.0), L((size_t)n * n);
  for (int i = 0; i < n; ++i) {
    double x = b.z1[i]; double p0 = 1, p1 = x;
    L[i * n + 0] = 1; if (n > 1) L[i * n + 1] = x;
    for (int k = 1; k + 1 < n; ++k) { double p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1); L[i * n + k + 1] = p2; p0 = p1; p1 = p2; }
  }
  for (int i = 0; i < n; ++i) for (int k = 0; k < n; ++k) V[i * n + k] = L[i * n + k] - (k >= 2 ? L[i * n + k - 2] : 0.0);
  // invert V by Gauss-Jordan
  std::vector<double> A(V);
  for (int i = 0; i < n; ++i) Vi[i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c; for (int r = c + 1; r < n; ++r) if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    for (int k = 0; k < n; ++k) { std::swap(A[c * n + k], A[piv * n + k]); std::swap(Vi[c * n + k], Vi[piv * n + k]); }
    double dinv = 1.0 / A[c * n + c];
    for (int k = 0; k < n; ++k) { A[c * n + k] *= dinv; Vi[c * n + k] *= dinv; }
    for (int r = 0; r < n; ++r) if (r != c) { double f = A[r * n + c]; if (f != 0) for (int k = 0; k < n; ++k) { A[r * n + k] -= f * A[c * n + k]; Vi[r * n + k] -= f * Vi[c * n + k]; } }
  }
  std::vector<double> sig(n, 1.0);
  int k0 = n - ncut;
  for (int k = k0; k < n; ++k) { int kk = k + 1 - k0; sig[k] = 1.0 - wght * (double)(kk * kk) / (double)(ncut * ncut); }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += V[i * n + k] * sig[k] * Vi[k * n + j]; F[i * n + j] = s; }
}

// ------------------------------------------------------------------------------------------------ operators
int apply_E(nlk_ctx* c, const double* p, double* ep, const double* out_mul) {       // Nek cdabdtp, intype = 1
  const DevMesh& dm = c->dm; const int d = dm.ndim;
  Ptr3 w{{c->wk[0], c->wk[1], c->wk[2]}};
  launch_opgradt(dm, p, w, c->st);
  if (ctx_gs(c, w, d)) return 1;
  CPtr3 cw{{w.p[0], w.p[1], w.p[2]}};
  launch_opdiv_fused(dm, cw, ep, 1.0 / c->prm.density, dm.binvm1, out_mul, c->st);      // opbinv (binvm1*mask/rho) fused into the load
  return 0;
}

int ortho(nlk_ctx* c, double* p) {                           // Nek ortho: remove the mean if E is singular
  if (c->dm.has_outflow) return 0;
  if (global_dot(c, c->dm.N2, p, c->ones2, nullptr, c->d_red + 300)) return 1;
  launch_sub_mean(p, c->dm.N2, c->d_red + 300, 1.0 / (double)c->dm.N2_global, c->st);
  return 0;
}

int apply_precond(nlk_ctx* c, const double* r, double* z, const double* in_mul) {     // z = M^-1 (in_mul * r)
  const DevMesh& dm = c->dm;
  if (!c->have_schwarz) {                                   // mass-scaled identity
    launch_lin(z, dm.N2, 1.0, r, 0, nullptr, 0, nullptr, 0, nullptr, c->pw[5], c->st);   // pw[5] holds 1/bm2
    if (in_mul) launch_lin(z, dm.N2, 1.0, z, 0, nullptr, 0, nullptr, 0, nullptr, in_mul, c->st);
    return 0;
  }
  // single rank: the coarse branch (restrict -> A0^-1 -> ...) runs on a side stream concurrently with the Schwarz branch
  // multi-rank: the same, with the allreduce of the coarse right-hand side on a split communicator so that it cannot interleave
  // with the Schwarz ghost exchange of the main stream
  const bool side = c->have_coarse && c->st2 && (c->nccl.nranks <= 1 || (c->nccl.comm2 && c->swf));
  if (c->swf && c->prm.precond != 3) {                      // fused Schwarz (pull tables): 2 kernels, prolongation folded into the second
    auto coarse_chain = [&]() -> int {
      launch_coarse_part_w(dm, r, in_mul, c->crs_part, c->st);
      launch_vert_gather(dm, c->crs_part, c->crs_r, c->st);
      if (ctx_allreduce(c, c->crs_r, (int)dm.nvert, false)) return 1;
      if (c->coarse_sparse) return coarse_solve_sparse(c, c->crs_r, c->crs_y);
      launch_gemv(dm.A0inv, c->crs_r, c->crs_y, (int)dm.nvert, c->st);
      return 0;
    };
    static const bool swf_first = getenv("NLK_SWF_FIRST") != nullptr;      // A/B: enqueue the Schwarz kernel before the coarse chain
    if (side) {
      NLK_CUDA(cudaEventRecord(c->ev_in, c->st));
      NLK_CUDA(cudaStreamWaitEvent(c->st2, c->ev_in, 0));
      if (swf_first && swf_apply(c, r, in_mul, nullptr, z, nullptr, 1)) return 1;
      cudaStream_t main_st = c->st; c->st = c->st2;
      int rc = coarse_chain();
      c->st = main_st;
      if (rc) return 1;
      NLK_CUDA(cudaEventRecord(c->ev_crs, c->st2));
      if (swf_first) return swf_apply(c, r, in_mul, c->crs_y, z, c->ev_crs, 2);
    } else if (c->have_coarse) { if (coarse_chain()) return 1; }
    return swf_apply(c, r, in_mul, c->have_coarse ? c->crs_y : nullptr, z, side ? c->ev_crs : nullptr);    // kernel B waits for the coarse solve
  }
  if (side) {
    NLK_CUDA(cudaEventRecord(c->ev_in, c->st));
    NLK_CUDA(cudaStreamWaitEvent(c->st2, c->ev_in, 0));
    cudaStream_t main_st = c->st; c->st = c->st2;
    launch_coarse_restrict(dm, r, in_mul, c->crs_part, c->crs_r, c->st);
    int rc = 0;
    if (c->coarse_sparse) rc = coarse_solve_sparse(c, c->crs_r, c->crs_y);
    else launch_gemv(dm.A0inv, c->crs_r, c->crs_y, (int)dm.nvert, c->st);
    c->st = main_st;
    if (rc) return 1;
    NLK_CUDA(cudaEventRecord(c->ev_crs, c->st2));
  }
  launch_schwarz_embed(dm, r, in_mul, c->sw_w, c->st);
  if (ctx_gs(c, Ptr3{{c->sw_w, nullptr, nullptr}}, 1)) return 1;
  launch_schwarz_fdm(dm, c->sw_w, c->sw_z, c->sw_t, c->st);
  if (ctx_gs(c, Ptr3{{c->sw_t, nullptr, nullptr}}, 1)) return 1;
  launch_schwarz_gather(dm, c->sw_z, c->sw_t, z, c->st);
  if (c->prm.precond == 3) launch_fill(z, dm.N2, 0.0, c->st);     // debug: coarse term only
  if (side) {
    NLK_CUDA(cudaStreamWaitEvent(c->st, c->ev_crs, 0));
    launch_coarse_prolong_add(dm, c->crs_y, z, 1, c->st);
  } else if (c->have_coarse) {
    launch_coarse_restrict(dm, r, in_mul, c->crs_part, c->crs_r, c->st);
    if (ctx_allreduce(c, c->crs_r, (int)dm.nvert, false)) return 1;
    if (c->coarse_sparse) { if (coarse_solve_sparse(c, c->crs_r, c->crs_y)) return 1; }
    else launch_gemv(dm.A0inv, c->crs_r, c->crs_y, (int)dm.nvert, c->st);
    launch_coarse_prolong_add(dm, c->crs_y, z, 1, c->st);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ Helmholtz PCG (hmholtz + cggo)
// weights of (mask, h1, h2) for the streamed PCG: rebuilt only when the Helmholtz coefficients change (the BDF order ramp of
// each matvec); one slot per mesh mask (velocity components 0..2, temperature 3)
int cg_weights(nlk_ctx* c, const double* mask, double h1, double h2, int* slot_out) {
  const DevMesh& dm = c->dm;
  int slot = -1;
  for (int k = 0; k < 4; ++k) if (mask == dm.mask[k]) slot = k;
  if (slot < 0) { set_error("helmholtz_solve: mask is not one of the mesh masks"); return 1; }
  if (!c->cg_hd[slot]) { if (dev_alloc(c, &c->cg_hd[slot], dm.N1) || dev_alloc(c, &c->cg_wa[slot], dm.N1) || dev_alloc(c, &c->cg_wb[slot], dm.N1)) return 1; c->cg_key_h1[slot] = c->cg_key_h2[slot] = -1.0; }
  if (c->cg_key_h1[slot] != h1 || c->cg_key_h2[slot] != h2) {
    launch_cg_weights(dm, mask, h1, h2, c->cg_hd[slot], c->cg_wa[slot], c->cg_wb[slot], c->st);
    c->cg_key_h1[slot] = h1; c->cg_key_h2[slot] = h2;
  }
  *slot_out = slot;
  return 0;
}
int helmholtz_solve(nlk_ctx* c, double* rhs, double h1, double h2, const double* mask, double tol, double* x, int* iters) {
  const DevMesh& dm = c->dm;
  const bool multi = c->nccl.nranks > 1;
  int slot = -1;
  if (cg_weights(c, mask, h1, h2, &slot)) return 1;
  const double* hd = c->cg_hd[slot]; const double* wa = c->cg_wa[slot]; const double* wb = c->cg_wb[slot];
  if (ctx_gs(c, Ptr3{{rhs, nullptr, nullptr}}, 1)) return 1;
  launch_lin(c->cg_r, dm.N1, 1.0, rhs, 0, nullptr, 0, nullptr, 0, nullptr, mask, c->st);
  launch_fill(x, dm.N1, 0.0, c->st);
  launch_fill(c->cg_p, dm.N1, 0.0, c->st);
  launch_cg_init(dm, c->d_sc, tol, c->prm.cg_maxit, c->st);
  auto reduce_zr = [&](int first) -> int {
    launch_cg_update_reduce(dm, x, c->cg_r, c->cg_p, c->cg_w, wa, wb, c->d_sc, c->red, first, multi ? 1 : 0, c->st);
    if (multi) { if (ctx_allreduce(c, c->d_sc->red, 2, false)) return 1; launch_cg_finalize(c->d_sc, 0, dm.volvm1, c->st); }
    return 0;
  };
  if (reduce_zr(1)) return 1;
  const int batch = 4;
  int done = 0, it = 0;
  while (!done) {
    for (int b = 0; b < batch; ++b) {
      launch_axhelm_cg(dm, c->cg_p, c->cg_r, c->cg_w, hd, h1, h2, c->d_sc, c->cg_pap_partial, c->cg_pap_counter, multi ? 1 : 0, c->st);   // w = H p, p.Hp
      if (ctx_gs(c, Ptr3{{c->cg_w, nullptr, nullptr}}, 1)) return 1;
      if (multi) { if (ctx_allreduce(c, c->d_sc->red + 2, 1, false)) return 1; launch_cg_finalize(c->d_sc, 1, dm.volvm1, c->st); }
      if (reduce_zr(0)) return 1;
    }
    NLK_CUDA(cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(SolverScal), cudaMemcpyDeviceToHost, c->st));
    NLK_CUDA(cudaStreamSynchronize(c->st));
    done = c->h_sc->done; it = c->h_sc->iter;
    if (it >= c->prm.cg_maxit) break;
  }
  if (iters) *iters = it;
  c->cg_iters += it;
  return 0;
}

// all components of one Helmholtz system at once: persistent cooperative kernel when the problem is small and single-rank,
// else the streamed per-component solver.  sol[k] += (h1 A + h2 B)^-1 mask dssum(rhs[k]); rhs arrays are overwritten.
int helmholtz_solve_multi(nlk_ctx* c, int nf, double* const* rhs, double h1, double h2, const double* const* masks, double tol, double* const* sol) {
  const DevMesh& dm = c->dm;
  if (c->use_cgp && c->nccl.nranks <= 1) {
    CgArgs a{};
    for (int k = 0; k < nf; ++k) a.f[k] = CgField{c->cgm_x[k], rhs[k], c->cgm_p[k], c->cgm_w[k], masks[k], sol[k]};
    a.nf = nf; a.G = dm.G; a.bm1 = dm.bm1; a.D = dm.D; a.diagA = dm.diagA; a.diagB = dm.diagB; a.mult = dm.vmult; a.binv = dm.binvm1;
    a.gs_off = dm.gs_off; a.gs_idx = dm.gs_idx; a.ngs = dm.ngs; a.E = dm.E; a.h1 = h1; a.h2 = h2; a.tol = tol; a.vol = dm.volvm1;
    a.maxit = c->prm.cg_maxit; a.partial = c->red.partial; a.iters_out = c->d_cg_iters; a.iters_total = c->d_cg_total;
    if (launch_cg_persistent(dm, a, c->st)) return 0;
    c->use_cgp = false;                       // cooperative launch unavailable: fall back for good
  }
  for (int k = 0; k < nf; ++k) {
    if (helmholtz_solve(c, rhs[k], h1, h2, masks[k], tol, c->cg_x, nullptr)) return 1;
    launch_lin(sol[k], dm.N1, 1.0, sol[k], 1.0, c->cg_x, 0, nullptr, 0, nullptr, nullptr, c->st);
  }
  return 0;
}
// fold the device-side iteration counter of the persistent PCG into the host statistics (one small D2H copy)
int sync_cg_counter(nlk_ctx* c) {
  if (!c->d_cg_total) return 0;
  unsigned long long v = 0;
  NLK_CUDA(cudaMemcpyAsync(&v, c->d_cg_total, sizeof(v), cudaMemcpyDeviceToHost, c->st));
  NLK_CUDA(cudaStreamSynchronize(c->st));
  c->cg_iters += (long)(v - c->cg_total_seen); c->cg_total_seen = v;
  return 0;
}

// ------------------------------------------------------------------------------------------------ pressure FGMRES (uzawa_gmres)
int pressure_solve(nlk_ctx* c, const double* rhs, double tol, double* x, int* iters) {
  const DevMesh& dm = c->dm;
  const int m = c->prm.lgmres, maxit = c->prm.gmres_maxit;
  const size_t N2 = dm.N2;
  double* r = c->pw[0]; double* w = c->pw[1];
  const double norm_fac = 1.0 / std::sqrt(dm.volvm2);
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gamma(m + 2), cc(m);
  launch_fill(x, N2, 0.0, c->st);
  int iter = 0; bool conv = false;
  double* dh = c->d_red;          // device scalars: h[0..m), alpha2 at [m]
  while (!conv && iter < maxit) {
    if (iter == 0) launch_lin(r, N2, 1.0, rhs, 0, nullptr, 0, nullptr, 0, nullptr, dm.ml, c->st);
    else { if (apply_E(c, x, w, nullptr)) return 1; launch_lin(r, N2, 1.0, rhs, -1.0, w, 0, nullptr, 0, nullptr, dm.ml, c->st); }
    if (global_dot(c, N2, r, r, nullptr, dh)) return 1;
    if (ctx_read_scalars(c, 1)) return 1;
    gamma[0] = std::sqrt(c->h_red[0]);
    if (gamma[0] == 0.0) break;
    if (iter == 0 && gamma[0] * norm_fac < tol) { break; }
    launch_lin(c->gm_V, N2, 1.0 / gamma[0], r, 0, nullptr, 0, nullptr, 0, nullptr, nullptr, c->st);
    int j = 0;
    for (j = 1; j <= m; ++j) {
      ++iter;
      double* vj = c->gm_V + (size_t)(j - 1) * N2; double* zj = c->gm_Z + (size_t)(j - 1) * N2;
      { PhaseScope ps(c->ph, PH_PRECOND, c->st);
        if (apply_precond(c, vj, zj, dm.mu)) return 1;                        // z = M^-1 (mu v)
        if (ortho(c, zj)) return 1; }
      { PhaseScope ps(c->ph, PH_EAPPLY, c->st);
        if (apply_E(c, zj, w, dm.ml)) return 1; }                             // w = ml * E z
      { PhaseScope ps(c->ph, PH_ORTH, c->st);
        launch_multidot(c->gm_V, N2, j, w, N2, dh, c->red, c->st);
        if (ctx_allreduce(c, dh, j, false)) return 1;
        launch_multiaxpy_norm(w, c->gm_V, N2, j, dh, -1.0, N2, dh + m, c->red, c->st);   // w -= V h ; |w|^2
        if (ctx_allreduce(c, dh + m, 1, false)) return 1; }
      if (ctx_read_scalars(c, m + 1)) return 1;
      for (int i = 0; i < j; ++i) H[(size_t)i * m + (j - 1)] = c->h_red[i];
      for (int i = 0; i < j - 1; ++i) {
        double t = H[(size_t)i * m + j - 1];
        H[(size_t)i * m + j - 1] = cs[i] * t + sn[i] * H[(size_t)(i + 1) * m + j - 1];
        H[(size_t)(i + 1) * m + j - 1] = -sn[i] * t + cs[i] * H[(size_t)(i + 1) * m + j - 1];
      }
      double alpha = std::sqrt(c->h_red[m]);
      if (alpha == 0.0) { conv = true; break; }
      double hjj = H[(size_t)(j - 1) * m + j - 1];
      double l = std::sqrt(hjj * hjj + alpha * alpha);
      cs[j - 1] = hjj / l; sn[j - 1] = alpha / l; H[(size_t)(j - 1) * m + j - 1] = l;
      gamma[j] = -sn[j - 1] * gamma[j - 1]; gamma[j - 1] = cs[j - 1] * gamma[j - 1];
      double rnorm = std::fabs(gamma[j]) * norm_fac;
      if (rnorm < tol) { conv = true; break; }
      if (j == m || iter >= maxit) break;
      launch_lin(c->gm_V + (size_t)j * N2, N2, 1.0 / alpha, w, 0, nullptr, 0, nullptr, 0, nullptr, nullptr, c->st);
    }
    if (j > m) j = m;
    for (int k = j - 1; k >= 0; --k) {
      double t = gamma[k];
      for (int i = j - 1; i > k; --i) t -= H[(size_t)k * m + i] * cc[i];
      cc[k] = t / H[(size_t)k * m + k];
    }
    NLK_CUDA(cudaMemcpyAsync(dh, cc.data(), sizeof(double) * j, cudaMemcpyHostToDevice, c->st));
    NLK_CUDA(cudaStreamSynchronize(c->st));     // cc is a host temporary
    launch_multiaxpy(x, c->gm_Z, N2, j, dh, 1.0, N2, c->st);
  }
  if (ortho(c, x)) return 1;
  if (iters) *iters = iter;
  c->gmres_iters += iter;
  return 0;
}

// Pressure residual projection (Nek `setrhsp`/`gensolnp`, .par residualProj = yes, SIZE mxprev): the rhs is projected onto
// the span of up to pr_proj previous solutions kept E-orthonormal (X^T E X = I); both X and E X are stored so the
// projection costs two tall-skinny passes and the basis update one extra E-apply.  Reset at every matvec start.
int pressure_solve_projected(nlk_ctx* c, double* rhs, double tol, double* x, int* iters) {
  const DevMesh& dm = c->dm; const size_t N2 = dm.N2; const int mx = c->prm.pr_proj;
  if (mx <= 0) return pressure_solve(c, rhs, tol, x, iters);
  double* al = c->d_red + 400; double* be = c->d_red + 432; double* nr = c->d_red + 464;
  int m = c->nproj;
  if (m > 0) {
    launch_multidot(c->proj_X, N2, m, rhs, N2, al, c->red, c->st);
    if (ctx_allreduce(c, al, m, false)) return 1;
    launch_fill(c->proj_xbar, N2, 0.0, c->st);
    launch_multiaxpy(c->proj_xbar, c->proj_X, N2, m, al, 1.0, N2, c->st);
    launch_multiaxpy(rhs, c->proj_EX, N2, m, al, -1.0, N2, c->st);
  }
  if (pressure_solve(c, rhs, tol, x, iters)) return 1;                      // x = delta
  double* w = c->proj_w;
  if (m == mx) {                                                            // restart the space with the full solution
    launch_lin(x, N2, 1.0, x, 1.0, c->proj_xbar, 0, nullptr, 0, nullptr, nullptr, c->st);
    if (apply_E(c, x, w, nullptr)) return 1;
    if (global_dot(c, N2, x, w, nullptr, nr)) return 1;
    launch_scale_rsqrt(c->proj_X, x, N2, nr, c->st);
    launch_scale_rsqrt(c->proj_EX, w, N2, nr, c->st);
    c->nproj = 1;
    return 0;
  }
  // E-orthonormalise delta against X and append
  double* dl = c->proj_X + (size_t)m * N2; double* edl = c->proj_EX + (size_t)m * N2;
  if (apply_E(c, x, w, nullptr)) return 1;
  NLK_CUDA(cudaMemcpyAsync(dl, x, N2 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  if (m > 0) {
    launch_multidot(c->proj_X, N2, m, w, N2, be, c->red, c->st);
    if (ctx_allreduce(c, be, m, false)) return 1;
    launch_multiaxpy(dl, c->proj_X, N2, m, be, -1.0, N2, c->st);
    launch_multiaxpy(w, c->proj_EX, N2, m, be, -1.0, N2, c->st);
  }
  if (global_dot(c, N2, dl, w, nullptr, nr)) return 1;
  launch_scale_rsqrt(dl, dl, N2, nr, c->st);
  launch_scale_rsqrt(edl, w, N2, nr, c->st);
  c->nproj = m + 1;
  if (m > 0) launch_lin(x, N2, 1.0, x, 1.0, c->proj_xbar, 0, nullptr, 0, nullptr, nullptr, c->st);    // full solution
  return 0;
}

// ------------------------------------------------------------------------------------------------ time stepping
static void bdf_coeffs(int nbd, double* bd) {
  bd[0] = bd[1] = bd[2] = bd[3] = 0;
  if (nbd == 1) { bd[0] = 1; bd[1] = 1; }
  else if (nbd == 2) { bd[0] = 1.5; bd[1] = 2; bd[2] = -0.5; }
  else { bd[0] = 11.0 / 6.0; bd[1] = 3; bd[2] = -1.5; bd[3] = 1.0 / 3.0; }
}
static void ab_coeffs(int nab, int nbd, double* ab) {       // Nek setabbd, constant dt
  ab[0] = 1; ab[1] = ab[2] = 0;
  if (nab == 1) return;
  if (nab == 2) { if (nbd <= 2) { ab[0] = 1.5; ab[1] = -0.5; } else { ab[0] = 2; ab[1] = -1; } return; }
  if (nbd == 1) { ab[2] = 0.5 * (0.5 + 1.0 / 3.0); ab[1] = -0.5 - 2 * ab[2]; ab[0] = 1 - ab[1] - ab[2]; }
  else if (nbd == 2) { ab[2] = 2.0 / 3.0; ab[1] = -1 - 2 * ab[2]; ab[0] = 1 - ab[1] - ab[2]; }
  else { ab[0] = 3; ab[1] = -3; ab[2] = 1; }
}

// setup_nek (src/neklab_nek_setup.f90:193-224): dt = cfl_limit/ctarg, nsteps = ceiling(tau/dt), dt = tau/nsteps
int step_setup(nlk_ctx* c, double tau, bool transpose) {
  c->adjoint = transpose;
  return step_setup_cfl(c, tau, c->prm.cfl_limit, CPtr3{{c->U[0], c->U[1], c->U[2]}});
}
int step_setup_cfl(nlk_ctx* c, double tau, double cfl_limit, CPtr3 u) {
  const DevMesh& dm = c->dm;
  if (tau <= 0) { set_error("invalid endtime (tau <= 0)"); return 1; }
  launch_cfl(dm, u, c->d_red, c->red, c->st);
  if (ctx_allreduce(c, c->d_red, 1, true)) return 1;
  if (ctx_read_scalars(c, 1)) return 1;
  double ctarg = c->h_red[0];
  if (ctarg == 0.0) {
    if (c->dt <= 0) { set_error("zero base flow and no previous dt: cannot choose a time step"); return 1; }
    c->nsteps = (int)std::ceil(tau / c->dt - 1e-12);
  } else {
    double dt = cfl_limit / ctarg;
    c->nsteps = (int)std::ceil(tau / dt);
    c->dt = tau / c->nsteps;
  }
  return 0;
}

static int reset_history(nlk_ctx* c) {
  const DevMesh& dm = c->dm;
  for (int k = 0; k < dm.ndim; ++k) {
    for (int l = 0; l < 2; ++l) NLK_CUDA(cudaMemsetAsync(c->vlag[l][k], 0, dm.N1 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(c->exx1[k], 0, dm.N1 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(c->exx2[k], 0, dm.N1 * sizeof(double), c->st));
  }
  NLK_CUDA(cudaMemsetAsync(c->prlag, 0, dm.N2 * sizeof(double), c->st));
  if (c->prm.ifheat) {
    for (int l = 0; l < 2; ++l) NLK_CUDA(cudaMemsetAsync(c->tlag[l], 0, dm.N1 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(c->vgradt1, 0, dm.N1 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(c->vgradt2, 0, dm.N1 * sizeof(double), c->st));
  }
  c->nproj = 0;
  return 0;
}

// one Nek `nek_advance` in perturbation mode (fluidp + heatp [+ q_filter])
int step_advance(nlk_ctx* c, int istep) {
  const DevMesh& dm = c->dm; const int d = dm.ndim; const nlk_params& P = c->prm;
  const double dt = c->dt, rho = P.density;
  const int nbd = std::min(istep, P.torder), nab = std::min(istep, 3);
  double bd[4], ab[3]; bdf_coeffs(nbd, bd); ab_coeffs(nab, nbd, ab);
  cudaStream_t st = c->st;
  int ph = c->ph.begin(PH_MAKEF, st);
  // ---- igeom = 1: makefp = makeufp + advabp(_adjoint) + makextp + makebdfp ; lagfieldp
  for (int k = 0; k < d; ++k) {
    const double* f0 = (P.ifheat && P.buoyancy[k] != 0.0 && !c->adjoint) ? c->tp : nullptr;      // adjoint: the buoyancy coupling moves to the temperature equation
    const int fslot = c->nonlinear ? 0 : 1;                                 // neklab_forcing(..., ipert = jp): neklab_nek_forcing.f90:96-114
    const double* f1 = c->has_forcing[fslot] ? c->forcing[fslot][k] : nullptr;
    if (f0 || f1) launch_lin(c->bf[k], dm.N1, P.buoyancy[k], f0, 1.0, f1, 0, nullptr, 0, nullptr, dm.bm1, st);
    else NLK_CUDA(cudaMemsetAsync(c->bf[k], 0, dm.N1 * sizeof(double), st));
  }
  CPtr3 Ub{{c->U[0], c->U[1], c->U[2]}}, up{{c->vp[0], c->vp[1], c->vp[2]}};
  Ptr4 bf4{{c->bf[0], c->bf[1], c->bf[2], nullptr}};
  if (c->nonlinear) {
    launch_convect(dm, CPtr4{{c->vp[0], c->vp[1], c->vp[2], nullptr}}, d, up, bf4, -rho, 1, st);  // advab: u.grad u (Nek plan3/makef)
  } else if (!c->adjoint) {
    launch_convect(dm, CPtr4{{c->U[0], c->U[1], c->U[2], nullptr}}, d, up, bf4, -rho, 1, st);     // u'.grad U
    launch_convect(dm, CPtr4{{c->vp[0], c->vp[1], c->vp[2], nullptr}}, d, Ub, bf4, -rho, 1, st);  // U.grad u'
  } else {
    launch_convect_adj(dm, Ub, up, Ptr3{{c->bf[0], c->bf[1], c->bf[2]}}, -rho, 1, st);            // (grad U)^T u'
    launch_convect(dm, CPtr4{{c->vp[0], c->vp[1], c->vp[2], nullptr}}, d, Ub, bf4, +rho, 1, st);
    // exptA_temp_linop%rmatvec (exponential_propagator_temp.f90:62-107): - theta' grad(T_base), the transpose of u'.grad T
    if (P.ifheat) launch_convect_adj(dm, CPtr3{{c->T, nullptr, nullptr}}, CPtr3{{c->tp, nullptr, nullptr}}, Ptr3{{c->bf[0], c->bf[1], c->bf[2]}}, -P.rhocp, 1, st, 1);
  }
  RhsTail t{};
  int nf = d;
  for (int k = 0; k < d; ++k) { t.bf[k] = c->bf[k]; t.e1[k] = c->exx1[k]; t.e2[k] = c->exx2[k]; t.u[k] = c->vp[k]; t.lag1[k] = c->vlag[0][k]; t.lag2[k] = c->vlag[1][k]; t.coef[k] = rho / dt; }
  if (P.ifheat) {
    if (c->nonlinear) { set_error("nonlinear Boussinesq stepper (nek_system_temp) is out of scope this round"); return 1; }
    NLK_CUDA(cudaMemsetAsync(c->bq, 0, dm.N1 * sizeof(double), st));
    Ptr4 bq4{{c->bq, nullptr, nullptr, nullptr}};
    if (!c->adjoint) {
      launch_convect(dm, CPtr4{{c->T, nullptr, nullptr, nullptr}}, 1, up, bq4, -P.rhocp, 1, st);    // u'.grad T
      launch_convect(dm, CPtr4{{c->tp, nullptr, nullptr, nullptr}}, 1, Ub, bq4, -P.rhocp, 1, st);   // U.grad T'
    } else {                                                                                        // adjoint: + U.grad theta' + buoyancy . u'
      launch_convect(dm, CPtr4{{c->tp, nullptr, nullptr, nullptr}}, 1, Ub, bq4, +P.rhocp, 1, st);
      for (int k = 0; k < d; ++k) if (P.buoyancy[k] != 0.0) launch_axpy_mm(c->bq, dm.N1, c->bq, P.buoyancy[k], c->vp[k], dm.bm1, nullptr, st);
    }
    t.bf[nf] = c->bq; t.e1[nf] = c->vgradt1; t.e2[nf] = c->vgradt2; t.u[nf] = c->tp; t.lag1[nf] = c->tlag[0]; t.lag2[nf] = c->tlag[1]; t.coef[nf] = P.rhocp / dt;
    ++nf;
  }
  launch_rhs_tail(dm, t, nf, ab[0], ab[1], ab[2], bd[1], bd[2], bd[3], st);
  c->ph.end(ph, st); ph = c->ph.begin(PH_VRES, st);
  // ---- igeom = 2: velocity.  cresvipp
  const double h1 = P.viscosity, h2 = rho * bd[0] / dt;
  if (!c->nonlinear)    // bcdirvc: homogeneous for perturbations; the nonlinear state keeps its (inflow) boundary values
    for (int k = 0; k < d; ++k) launch_lin(c->vp[k], dm.N1, 1.0, c->vp[k], 0, nullptr, 0, nullptr, 0, nullptr, dm.mask[k], st);
  double* pext = c->pw[3];
  const int variant = P.step_variant;
  if (nbd == 3 && !(variant & 8)) launch_lin(pext, dm.N2, 2.0, c->prp, -1.0, c->prlag, 0, nullptr, 0, nullptr, nullptr, st);    // extrapprp
  else NLK_CUDA(cudaMemcpyAsync(pext, c->prp, dm.N2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  Ptr3 gp{{c->wk[3], c->wk[4], c->wk[5]}};
  launch_opgradt(dm, pext, gp, st);
  for (int k = 0; k < d; ++k) {
    launch_axhelm(dm, c->vp[k], c->wk[6], h1, h2, st);
    launch_lin(gp.p[k], dm.N1, 1.0, gp.p[k], 1.0, c->bf[k], -1.0, c->wk[6], 0, nullptr, nullptr, st);                           // res = Dtp* + bf - H u
  }
  c->ph.end(ph, st); ph = c->ph.begin(PH_HELM, st);
  // ophinv: Jacobi-PCG on every component (one persistent launch for all of them when the problem is small)
  {
    double* rhsv[3] = {gp.p[0], gp.p[1], gp.p[2]}; const double* mk[3] = {dm.mask[0], dm.mask[1], dm.mask[2]}; double* solv[3] = {c->vp[0], c->vp[1], c->vp[2]};
    if (helmholtz_solve_multi(c, d, rhsv, h1, h2, mk, P.vtol, solv)) return 1;
  }
  c->ph.end(ph, st); ph = c->ph.begin(PH_PRHS, st);
  // incomprp: E dp = -(bd1/dt) D u*, solved in the dt/bd1-scaled form (rhs = -D u*, dp = x * bd1/dt)
  double* rhs = c->pw[4];
  launch_opdiv(dm, CPtr3{{c->vp[0], c->vp[1], c->vp[2]}}, rhs, -1.0, st);
  if (ortho(c, rhs)) return 1;
  if (!(variant & 1)) NLK_CUDA(cudaMemcpyAsync(c->prlag, c->prp, dm.N2 * sizeof(double), cudaMemcpyDeviceToDevice, st));        // lagpresp
  if (!(variant & 2)) NLK_CUDA(cudaMemcpyAsync(c->prp, pext, dm.N2 * sizeof(double), cudaMemcpyDeviceToDevice, st));            // up = prextr (+ dp below)
  double* xs = c->pw[3];                                                                                                        // pext is dead from here
  c->ph.end(ph, st); ph = c->ph.begin(PH_PRES, st);
  if (pressure_solve_projected(c, rhs, P.ptol, xs, nullptr)) return 1;
  c->ph.end(ph, st); ph = c->ph.begin(PH_CORR, st);
  launch_lin(c->prp, dm.N2, 1.0, c->prp, bd[0] / dt, xs, 0, nullptr, 0, nullptr, nullptr, st);                                  // add3(up, prextr, dp)
  Ptr3 w{{c->wk[0], c->wk[1], c->wk[2]}};
  launch_opgradt(dm, xs, w, st);
  if (ctx_gs(c, w, d)) return 1;
  for (int k = 0; k < d; ++k) launch_axpy_mm(c->vp[k], dm.N1, c->vp[k], 1.0 / rho, w.p[k], dm.binvm1, dm.mask[k], st);          // opbinv + opadd2cm
  c->ph.end(ph, st);
  // ---- igeom = 2: temperature (cdscalp)
  if (P.ifheat) {
    PhaseScope ps(c->ph, PH_HEAT, st);
    const double h1t = P.conductivity, h2t = P.rhocp * bd[0] / dt;
    launch_lin(c->tp, dm.N1, 1.0, c->tp, 0, nullptr, 0, nullptr, 0, nullptr, dm.mask[3], st);                                   // bcdirsc (homogeneous)
    launch_axhelm(dm, c->tp, c->wk[6], h1t, h2t, st);
    launch_lin(c->wk[7], dm.N1, 1.0, c->bq, -1.0, c->wk[6], 0, nullptr, 0, nullptr, nullptr, st);
    { double* rhst[1] = {c->wk[7]}; const double* mk[1] = {dm.mask[3]}; double* solt[1] = {c->tp};
      if (helmholtz_solve_multi(c, 1, rhst, h1t, h2t, mk, P.ttol, solt)) return 1; }
  }
  // ---- q_filter (param(103) > 0)
  if (P.filter_weight > 0 && c->filterF) {
    Ptr4 fu{{c->vp[0], c->vp[1], d == 3 ? c->vp[2] : c->tp, c->tp}};
    int nfl = d + (P.ifheat ? 1 : 0);
    PhaseScope ps(c->ph, PH_FILTER, st);
    launch_filter(dm, c->filterF, fu, nfl, st);
  }
  c->ph.resolve(st);
  c->steps += 1;
  return 0;
}

int reset_history_pub(nlk_ctx* c) { return reset_history(c); }

// ------------------------------------------------------------------------------------------------ coupled base flow + perturbation
int bank2_ensure(nlk_ctx* c) {
  if (c->bank2) return 0;
  if (c->prm.ifheat) { set_error("coupled base-flow/perturbation stepping with temperature is out of scope (nonlinear Boussinesq stepper)"); return 1; }
  const DevMesh& dm = c->dm;
  StateBank* b = new StateBank();
  for (int k = 0; k < dm.ndim; ++k) {
    if (dev_alloc(c, &b->vp[k], dm.N1) || dev_alloc(c, &b->vlag[0][k], dm.N1) || dev_alloc(c, &b->vlag[1][k], dm.N1) ||
        dev_alloc(c, &b->exx1[k], dm.N1) || dev_alloc(c, &b->exx2[k], dm.N1)) { delete b; return 1; }
  }
  if (dev_alloc(c, &b->prp, dm.N2) || dev_alloc(c, &b->prlag, dm.N2) || dev_alloc(c, &b->tp, dm.N1)) { delete b; return 1; }
  if (c->prm.pr_proj > 0) {
    if (dev_alloc(c, &b->proj_X, (size_t)c->prm.pr_proj * dm.N2) || dev_alloc(c, &b->proj_EX, (size_t)c->prm.pr_proj * dm.N2)) { delete b; return 1; }
  }
  c->bank2 = b;
  return 0;
}
void bank_swap(nlk_ctx* c) {
  StateBank* b = c->bank2;
  for (int k = 0; k < 3; ++k) {
    std::swap(c->vp[k], b->vp[k]); std::swap(c->vlag[0][k], b->vlag[0][k]); std::swap(c->vlag[1][k], b->vlag[1][k]);
    std::swap(c->exx1[k], b->exx1[k]); std::swap(c->exx2[k], b->exx2[k]);
  }
  std::swap(c->prp, b->prp); std::swap(c->tp, b->tp); std::swap(c->prlag, b->prlag);
  std::swap(c->proj_X, b->proj_X); std::swap(c->proj_EX, b->proj_EX); std::swap(c->nproj, b->nproj);
}
// Nek `nek_advance` with ifpert and ifbase (SURVEY.md call stack: [ifbase -> fluid] fluidp per igeom): the explicit terms of
// the perturbation (igeom = 1) see the base flow of the previous time level, which is then advanced by the nonlinear step.
// c->U must alias the base state held in bank2 (the caller sets that up).
int coupled_advance(nlk_ctx* c, int istep) {
  if (step_advance(c, istep)) return 1;
  const bool adj = c->adjoint;
  bank_swap(c); c->nonlinear = true; c->adjoint = false;
  const int rc = step_advance(c, istep);
  c->nonlinear = false; c->adjoint = adj; bank_swap(c);
  c->steps -= 1;                                      // one nek_advance
  return rc;
}
void make_filter_matrix(const Basis& b, double w, double cutoff, std::vector<double>& F) { filter_matrix(b, w, cutoff, F); }
void make_fdm_1d(const Basis& b, double lm, double ll, double lr, int bcl, int bcr, double* S, double* lam, int* nact) { fdm_1d(b, lm, ll, lr, bcl, bcr, S, lam, nact); }
double mesh_diag_local(const HostMesh& hm, int64_t e, int p) { return diag_local(hm, e, p); }
std::vector<double> mat_transpose(const std::vector<double>& M, int r, int cdim) { return transpose(M, r, cdim); }

}  // namespace nlk
