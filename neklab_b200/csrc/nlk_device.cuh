// Device-side structures and kernel launch wrappers of libnlk (sm_100a, FP64).
// Kernel numbering (K1..K16) follows SURVEY.md §2.3; each wrapper cites the Nek5000 routine it restates.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "nlk_host.hpp"

namespace nlk {

#define NLK_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      nlk::set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                     std::to_string(__LINE__));                                                     \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

struct Ptr3 { double* p[3]; };
struct CPtr3 { const double* p[3]; };
struct Ptr4 { double* p[4]; };
struct CPtr4 { const double* p[4]; };

// all read-only mesh data in HBM
struct DevMesh {
  int ndim, n, m, q, np1, np2, npd, ng;
  int64_t E;
  size_t N1, N2, Nd;
  // 1-D operators (row-major) and their transposes
  double *D, *Dt, *I12, *I12t, *D12, *D12t, *I1d, *I1dt, *Dd, *Ddt, *w2;
  // geometry
  double* G;        // [E][ng][np1]
  double* bm1;      // [E][np1]
  double* binvm1;
  double* vmult;
  double* mask[4];  // velocity comps 0..2, temperature 3
  double* rxw2;     // [E][d*d][np2]   W2 * rx on mesh 2
  double* bm2;      // [E][np2]
  double* ml;       // sqrt(1/bm2)
  double* mu;       // sqrt(bm2)
  double* rxd;      // [E][d*d][npd]   dealias metrics (weights folded)
  double* rxj;      // [E][d*d][np1]   rx / J  (for compute_cfl)
  double* dri;      // [n] inverse GLL spacing (compute_cfl)
  double* diagA;    // dssum(diag(D^T G D))
  double* diagB;    // dssum(bm1)
  // gather-scatter
  int32_t *gs_off, *gs_idx;
  int ngs;
  // the same groups split by multiplicity for launch_gs: dense index pairs / quadruples (no offset array, one vector load per
  // group) and a CSR for the rest; each list keeps the owner-address order of the full CSR
  int32_t *gs2 = nullptr, *gs4 = nullptr, *gsr_off = nullptr, *gsr_idx = nullptr;
  int ngs2 = 0, ngs4 = 0, ngsr = 0;
  // Schwarz / coarse
  double* fdmS;     // [E][d][n*n]
  double* fdmSt;    // transposes
  double* fdmDinv;  // [E][np1]
  double* swt;      // [E][np2]
  int64_t* vertex;  // [E][nv] (1-based)
  int32_t *vert_off, *vert_ec;
  int64_t nvert;
  double* A0inv;    // [nvert][nvert]
  double volvm1, volvm2;    // GLOBAL volumes
  int has_outflow;
  int64_t N2_global;
};

// device scalar block shared by the solvers (lives in HBM; mirrored to pinned host memory on demand)
struct SolverScal {
  double rtz1, rtz2, rbn2, rbn0, pap, alpha, beta, tol;
  double red[8];
  int iter, done, maxit, pad;
};

struct Reducer {          // scratch for deterministic two-stage reductions
  double* partial;        // [nslots][maxblocks]
  unsigned int* counter;
  int maxblocks;
};

// persistent cooperative PCG (nlk_cgp.cu)
struct CgField { double* x; double* r; double* p; double* w; const double* mask; double* sol; };
struct CgArgs {
  CgField f[3]; int nf;
  const double *G, *bm1, *D, *diagA, *diagB, *mult, *binv;
  const int32_t *gs_off, *gs_idx; int ngs; int64_t E;
  double h1, h2, tol, vol; int maxit;
  double* partial; int* iters_out; unsigned long long* iters_total;
};
bool launch_cg_persistent(const DevMesh& dm, CgArgs& a, cudaStream_t st);

// ---- launch wrappers (all asynchronous on `st`) -------------------------------------------------------
// K1  axhelm (hmholtz.f): w = h1*(D^T G D)u + h2*B u, element-local.  If pz != nullptr the CG direction update
//     p = r*dinv + beta*p is fused in front (u is then p, read-modify-write), with dinv = 1/(h1*diagA+h2*diagB).
void launch_axhelm(const DevMesh& dm, const double* u, double* w, double h1, double h2, cudaStream_t st);
// pap_partial: one double per launched block (>= number of elements); with defer != 0 only sc->red[2] is stored (multi-rank)
void launch_axhelm_cg(const DevMesh& dm, double* p, const double* r, double* w, const double* hd, double h1, double h2, SolverScal* sc,
                      double* pap_partial, unsigned int* pap_counter, int defer, cudaStream_t st);
// per-(h1,h2) PCG weights: hd = mask/(h1 diagA + h2 diagB), wa = hd*mult, wb = mask*mult*binv
void launch_cg_weights(const DevMesh& dm, const double* mask, double h1, double h2, double* hd, double* wa, double* wb, cudaStream_t st);
// K2  dssum (gs_op add) on up to 3 fields; local part.
void launch_gs(const DevMesh& dm, Ptr3 f, int nf, cudaStream_t st);
// generic pointwise: out = (a0*x0 + a1*x1 + a2*x2 + a3*x3) * (mul ? mul : 1)
void launch_lin(double* out, size_t n, double a0, const double* x0, double a1, const double* x1, double a2, const double* x2,
                double a3, const double* x3, const double* mul, cudaStream_t st);
void launch_fill(double* out, size_t n, double v, cudaStream_t st);
// K5  opdiv / opgradt (navier1.f multd, cdtp)
void launch_opdiv(const DevMesh& dm, CPtr3 u, double* p, double scale, cudaStream_t st);
void launch_opdiv_fused(const DevMesh& dm, CPtr3 u, double* p, double scale, const double* in_mul, const double* out_mul, cudaStream_t st);
void launch_opgradt(const DevMesh& dm, const double* p, Ptr3 w, cudaStream_t st);
// K3/K4 convection (convect.f convect_new / convect_adj): out_f (+)= alpha * J^T W[(J C . rx) . grad J u_f]
void launch_convect(const DevMesh& dm, CPtr4 u, int nf, CPtr3 C, Ptr4 out, double alpha, int accumulate, cudaStream_t st);
void launch_convect_adj(const DevMesh& dm, CPtr3 U, CPtr3 c, Ptr3 out, double alpha, int accumulate, cudaStream_t st, int nj = -1);   // nj pairs (c_j, U_j) summed (default ndim)
// K7  fused rhs tail: makextp + makebdfp + lagfieldp for ncomp fields
struct RhsTail { double* bf[4]; double* e1[4]; double* e2[4]; const double* u[4]; double* lag1[4]; double* lag2[4]; double coef[4]; };
void launch_rhs_tail(const DevMesh& dm, const RhsTail& t, int nf, double ab0, double ab1, double ab2, double bd1, double bd2, double bd3,
                     cudaStream_t st);
// K8  CG vector phases (cggo)
void launch_cg_init(const DevMesh& dm, SolverScal* sc, double tol, int maxit, cudaStream_t st);
// defer != 0: only store the local sums in sc->red[] (multi-rank: allreduce, then launch_cg_finalize)
void launch_cg_update_reduce(const DevMesh& dm, double* x, double* r, const double* p, const double* w, const double* wa, const double* wb,
                             SolverScal* sc, Reducer red, int first, int defer, cudaStream_t st);
// reductions: out[0..nout) = sum_i a_i*b_i*(c_i)  for up to 4 (a,b) pairs sharing weight c (nullable)
void launch_dot(size_t n, CPtr4 a, CPtr4 b, int npairs, const double* c, double* out, Reducer red, cudaStream_t st);
// K9  multi-dot h[j] = sum_i V[j][i]*w[i], j<k (uzawa_gmres CGS) and w -= sum_j h[j] V[j]
void launch_multidot(const double* V, size_t ld, int k, const double* w, size_t n, double* h, Reducer red, cudaStream_t st);
void launch_multiaxpy(double* w, const double* V, size_t ld, int k, const double* h, double sign, size_t n, cudaStream_t st);
// K10 Schwarz smoother pieces (hsmg_schwarz / hsmg_fdm / extrude analogues)
void launch_schwarz_embed(const DevMesh& dm, const double* r, const double* mul, double* w, cudaStream_t st);
void launch_schwarz_fdm(const DevMesh& dm, const double* w, double* z, double* t, cudaStream_t st);
void launch_schwarz_gather(const DevMesh& dm, const double* z, const double* t, double* out, cudaStream_t st);
// K11 coarse grid (crs_solve analogue)
void launch_coarse_restrict(const DevMesh& dm, const double* r, const double* mul, double* part, double* rc, cudaStream_t st);
void launch_gemv(const double* A, const double* x, double* y, int n, cudaStream_t st);
void launch_coarse_prolong_add(const DevMesh& dm, const double* c, double* z, int accumulate, cudaStream_t st);
// K14 compute_cfl
void launch_cfl(const DevMesh& dm, CPtr3 u, double* out, Reducer red, cudaStream_t st);
// K15 q_filter: u <- (F (x) F (x) F) u  for nf fields in place
void launch_filter(const DevMesh& dm, const double* F, Ptr4 u, int nf, cudaStream_t st);
// pressure mean removal (ortho)
void launch_sub_mean(double* p, size_t n, const double* sum, double inv_count, cudaStream_t st);
// interface pack/unpack for the multi-rank gather-scatter
void launch_pack(const double* u, const int32_t* rep, int cnt, double* buf, cudaStream_t st);
void launch_unpack_add(double* u, const int32_t* off, const int32_t* idx, int cnt, const double* buf, cudaStream_t st);
// seeded C0 field generator (nek_drand analogue)

// out = (x0 ? x0 : 0) + a * x1 * m1 * (m2 ? m2 : 1)
void launch_axpy_mm(double* out, size_t n, const double* x0, double a, const double* x1, const double* m1, const double* m2, cudaStream_t st);
void launch_multiaxpy_norm(double* w, const double* V, size_t ld, int k, const double* h, double sign, size_t n, double* out, Reducer red, cudaStream_t st);
void launch_scale_rsqrt(double* out, const double* in, size_t n, const double* s, cudaStream_t st);
void launch_recip(double* out, const double* x, size_t n, cudaStream_t st);
void launch_schwarz_count(const DevMesh& dm, const double* w, double* z, double* t, cudaStream_t st);
void launch_schwarz_gather_nowt(const DevMesh& dm, const double* z, const double* t, double* out, cudaStream_t st);
void launch_rand_field_impl(double* out, const double* x, const double* y, const double* z, const int64_t* lglel, int np1, size_t N1,
                            uint64_t seed, int comp, cudaStream_t st);
void launch_cg_finalize(SolverScal* sc, int which, double vol, cudaStream_t st);
// compile-time-sized versions (nlk_kernels_tp.cu); return false when (lx1, lxd) has no instantiation -> runtime fallback
bool tp_opdiv(const DevMesh& dm, CPtr3 u, double* p, double scale, const double* in_mul, const double* out_mul, cudaStream_t st);
bool tp_opgradt(const DevMesh& dm, const double* p, Ptr3 w, cudaStream_t st);
bool tp_schwarz_fdm(const DevMesh& dm, const double* w, double* z, double* t, cudaStream_t st);
// fused Schwarz with pull tables (nlk_kernels_tp.cu; orchestration in nlk_schwarz.cu)
bool tp_swf_a(const DevMesh& dm, const double* r, const double* mul, const int32_t* t1, const double* ghost, double* zint, double* ZF, cudaStream_t st);
bool tp_swf_b(const DevMesh& dm, const double* zint, const double* ZF, const int32_t* t2, const double* ghost, const double* yc, double* out, cudaStream_t st);
void launch_coarse_part_w(const DevMesh& dm, const double* r, const double* mul, double* part, cudaStream_t st);
void launch_swf_pack(const double* src, const double* mul, const int32_t* idx, int n, double* out, cudaStream_t st);
void launch_vert_gather(const DevMesh& dm, const double* part, double* rc, cudaStream_t st);
// zero-copy scalar read-back: one thread copies `count` doubles into pinned host memory, fences, then publishes `seq`
void launch_publish(const double* d_src, int count, double* h_dst, unsigned int* h_seq, unsigned int seq, cudaStream_t st);
// exptA_proj_linop `proj_alpha`: per plane group, a_c = <2 u cv>, a_s = <2 u sv> (bm1-weighted); then u = cv a_c + sv a_s
void launch_planar_proj(double* u, const double* bm1, const double* cv, const double* sv, const int32_t* off, const int32_t* idx, const int32_t* gid,
                        int64_t ngroups, size_t N1, double* coef, cudaStream_t st);
bool tp_convect(const DevMesh& dm, CPtr4 u, int nf, CPtr3 C, Ptr4 out, double alpha, int accumulate, cudaStream_t st);
bool tp_convect_adj(const DevMesh& dm, CPtr3 U, CPtr3 c, Ptr3 out, double alpha, int accumulate, cudaStream_t st, int nj);
extern thread_local long g_launches;   // counts kernel launches issued through these wrappers

}  // namespace nlk
