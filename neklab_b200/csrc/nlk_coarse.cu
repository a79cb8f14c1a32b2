// Sparse coarse-grid level for large vertex counts (3-D meshes): the exact Galerkin operator A0 = R0 E R0^T is probed on
// the device with a distance-2 colouring of its sparsity graph, stored in CSR, and solved approximately by a fixed number
// of Jacobi-PCG iterations whose scalars never leave the device (FGMRES outside tolerates the inexact solve).
// Role in the reference: Nek5000 `crs_solve` (XXT/AMG coarse solve inside hsmg_solve; un-vendored).
#include "nlk_ctx.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace nlk {

__global__ void k_indicator(double* x, const int32_t* __restrict__ color, int c, int64_t n) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) x[v] = color[v] == c ? 1.0 : 0.0;
}

__device__ __forceinline__ double warp_sum_c(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Reductions without atomics or "last block" tails: a kernel leaves one partial sum per block, and every WARP of the next
// kernel re-sums those partials itself in a fixed order (bit-identical alpha / beta everywhere, deterministic, and no
// block barrier: the partial loads are issued first and overlap the kernel's own gather chain -- these kernels are pure
// latency chains of a few microseconds).  rz partials are double-buffered by iteration parity (beta needs the current
// and the previous r.z).
constexpr int CRS_MAXB_RZ = 296, CRS_MAXB_PQ = 1184;       // blocks per kernel = partials per buffer
template <int MAXB>
struct CrsPartials {                                        // the calling lane's share of the partials, kept in registers
  double v[(MAXB + 31) / 32];
  __device__ __forceinline__ void load(const double* __restrict__ partial, int nb, int lane) {
#pragma unroll
    for (int k = 0; k < (MAXB + 31) / 32; ++k) { const int b = lane + 32 * k; v[k] = b < nb ? partial[b] : 0.0; }
  }
  __device__ __forceinline__ double total() const {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < (MAXB + 31) / 32; ++k) t += v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  }
};
__device__ __forceinline__ void crs_block_partial(double acc, double* __restrict__ partial) {
  __shared__ double sp[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  acc = warp_sum_c(acc);
  if (lane == 0) sp[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sp[i]; partial[blockIdx.x] = t; }
}

// Search-direction update folded into the SpMV (one kernel less per iteration): with s = A z,
//   p <- z + beta p ,  q = A p <- s + beta q ,  pq partial = p.q       (beta = rz / rz_prev, 0 in the first iteration)
// Every row's p and q are touched by its own warp only; the gather reads z, which this kernel does not write.
__global__ void __launch_bounds__(256)
k_crs_spmv_pq(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val, const double* __restrict__ z,
              double* __restrict__ p, double* __restrict__ q, int n, const double* __restrict__ rz_cur, const double* __restrict__ rz_prev,
              int nb_rz, int first, double* __restrict__ pq_partial) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  CrsPartials<CRS_MAXB_RZ> pa, pb;
  if (!first) { pa.load(rz_cur, nb_rz, lane); pb.load(rz_prev, nb_rz, lane); }
  double acc = 0.0, beta = 0.0; bool have_beta = first;
  for (int row = blockIdx.x * wpb + wib; row < n; row += gridDim.x * wpb) {
    double s = 0.0;
    for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) s += val[k] * z[col[k]];
    s = warp_sum_c(s);
    if (!have_beta) { const double a = pa.total(), b = pb.total(); beta = b != 0.0 ? a / b : 0.0; have_beta = true; }
    if (lane == 0) {
      const double pn = z[row] + beta * p[row], qn = s + beta * q[row];
      p[row] = pn; q[row] = qn; acc += pn * qn;
    }
  }
  crs_block_partial(lane == 0 ? acc : 0.0, pq_partial);
}

// first: x = 0, p = q = 0, r = rhs ; else x += alpha p, r -= alpha q (alpha = rz / pq) ; then z = dinv r and the r.z partial
__global__ void __launch_bounds__(256)
k_crs_update(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ q,
             const double* __restrict__ dinv, const double* __restrict__ rhs, int n, const double* __restrict__ rz_cur, int nb_rz,
             const double* __restrict__ pq_partial, int nb_pq, int first, double* __restrict__ rz_out) {
  const int lane = threadIdx.x & 31;
  double alpha = 0.0;
  if (!first) {
    CrsPartials<CRS_MAXB_RZ> pa; CrsPartials<CRS_MAXB_PQ> pb;
    pa.load(rz_cur, nb_rz, lane); pb.load(pq_partial, nb_pq, lane);
    const double a = pa.total(), b = pb.total();
    alpha = b != 0.0 ? a / b : 0.0;
  }
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double ri;
    if (first) { ri = rhs[i]; x[i] = 0.0; p[i] = 0.0; q[i] = 0.0; }
    else { x[i] += alpha * p[i]; ri = r[i] - alpha * q[i]; }
    r[i] = ri;
    const double zi = ri * dinv[i]; z[i] = zi; acc += ri * zi;
  }
  crs_block_partial(acc, rz_out);
}

// The solve is a fixed sequence of 1 + 2*iters tiny kernels (a few microseconds each); issued one by one it is bound by
// launch latency (phase timing, 7 984 elements: the coarse branch made the preconditioner 0.5 ms per application against
// 0.16 ms for the Schwarz branch).  The sequence is therefore captured once into a CUDA graph and replayed.
static int coarse_issue(nlk_ctx* c, const double* rc, double* yc, cudaStream_t st) {
  const int n = (int)c->dm.nvert;
  const int gu = std::min((n + 255) / 256, CRS_MAXB_RZ), gs = std::min((n + 7) / 8, CRS_MAXB_PQ);     // SpMV: one row per warp while it lasts
  double* rzb[2] = {c->crs_partial, c->crs_partial + CRS_MAXB_RZ}; double* pqb = c->crs_partial + 2 * CRS_MAXB_RZ;
  k_crs_update<<<gu, 256, 0, st>>>(yc, c->crs_rr, c->crs_z, c->crs_p, c->crs_q, c->crs_dinv, rc, n, nullptr, 0, nullptr, 0, 1, rzb[0]);
  for (int it = 0; it < c->crs_iters; ++it) {
    const double* cur = rzb[it & 1]; double* nxt = rzb[(it + 1) & 1];
    k_crs_spmv_pq<<<gs, 256, 0, st>>>(c->crs_rowptr, c->crs_col, c->crs_val, c->crs_z, c->crs_p, c->crs_q, n, cur, nxt, gu, it == 0, pqb);
    k_crs_update<<<gu, 256, 0, st>>>(yc, c->crs_rr, c->crs_z, c->crs_p, c->crs_q, c->crs_dinv, rc, n, cur, gu, pqb, gs, 0, nxt);
  }
  NLK_CUDA(cudaGetLastError());
  return 0;
}

int coarse_solve_sparse(nlk_ctx* c, const double* rc, double* yc) {
  cudaStream_t st = c->st;
  static const bool no_graph = getenv("NLK_NO_GRAPH") != nullptr;
  if (no_graph) { g_launches += 1 + 2 * c->crs_iters; return coarse_issue(c, rc, yc, st); }
  if (!c->crs_graph || c->crs_graph_in != rc || c->crs_graph_out != yc) {
    if (c->crs_graph) { cudaGraphExecDestroy(c->crs_graph); c->crs_graph = nullptr; }
    cudaGraph_t g = nullptr;
    NLK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    int rc_issue = coarse_issue(c, rc, yc, st);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc_issue || e != cudaSuccess) { if (g) cudaGraphDestroy(g); set_error("coarse solve: graph capture failed"); return 1; }
    NLK_CUDA(cudaGraphInstantiate(&c->crs_graph, g, 0));
    cudaGraphDestroy(g);
    c->crs_graph_in = rc; c->crs_graph_out = yc;
  }
  NLK_CUDA(cudaGraphLaunch(c->crs_graph, st));
  g_launches += 1 + 2 * c->crs_iters;           // kernels executed by the graph
  return 0;
}

int coarse_setup_sparse(nlk_ctx* c) {
  const HostMesh& hm = c->mesh->hm; const DevMesh& dm = c->dm;
  const int nv = hm.nv; const int64_t nvt = hm.nvert, Eg = hm.Eg;
  // ---- S1: vertices sharing an element (incl. self), from the GLOBAL element->vertex table
  std::vector<std::vector<int32_t>> S1(nvt);
  for (int64_t e = 0; e < Eg; ++e) for (int a = 0; a < nv; ++a) { auto& l = S1[hm.vertex_all[e * nv + a] - 1]; for (int b = 0; b < nv; ++b) l.push_back((int32_t)(hm.vertex_all[e * nv + b] - 1)); }
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t v = 0; v < nvt; ++v) { auto& l = S1[v]; std::sort(l.begin(), l.end()); l.erase(std::unique(l.begin(), l.end()), l.end()); }
  // ---- S2 = S1 o S1: sparsity pattern of A0 (E couples elements that share a velocity node)
  std::vector<std::vector<int32_t>> S2(nvt);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t v = 0; v < nvt; ++v) { auto& l = S2[v]; for (int32_t w : S1[v]) l.insert(l.end(), S1[w].begin(), S1[w].end()); std::sort(l.begin(), l.end()); l.erase(std::unique(l.begin(), l.end()), l.end()); }
  // ---- greedy distance-2 colouring of the S2 graph: two vertices of one colour have no common S2 neighbour
  std::vector<int32_t> color(nvt, -1);
  int ncolors = 0;
  {
    std::vector<int32_t> stamp;            // stamp[col] == v  <=> colour forbidden for v
    for (int64_t v = 0; v < nvt; ++v) {
      for (int32_t w : S2[v]) for (int32_t u : S2[w]) { int32_t cu = color[u]; if (cu >= 0) { if ((size_t)cu >= stamp.size()) stamp.resize(cu + 1, -1); stamp[cu] = (int32_t)v; } }
      int cc = 0; while ((size_t)cc < stamp.size() && stamp[cc] == (int32_t)v) ++cc;
      color[v] = cc; ncolors = std::max(ncolors, cc + 1);
      if ((size_t)cc >= stamp.size()) stamp.resize(cc + 1, -1);
    }
  }
  // ---- CSR pattern and the (row, colour) -> position map
  std::vector<int32_t> rowptr(nvt + 1, 0);
  for (int64_t v = 0; v < nvt; ++v) rowptr[v + 1] = rowptr[v] + (int32_t)S2[v].size();
  const int64_t nnz = rowptr[nvt];
  std::vector<int32_t> col(nnz); std::vector<double> val(nnz, 0.0);
  for (int64_t v = 0; v < nvt; ++v) std::copy(S2[v].begin(), S2[v].end(), col.begin() + rowptr[v]);
  // ---- probing: one E-apply per colour
  int32_t* dcolor = nullptr; if (dev_upload(c, &dcolor, color)) return 1;
  if (dev_alloc(c, &c->crs_part, (size_t)hm.E << hm.ndim) || dev_alloc(c, &c->crs_r, nvt) || dev_alloc(c, &c->crs_y, nvt)) return 1;
  std::vector<double> r(nvt);
  for (int cc = 0; cc < ncolors; ++cc) {
    k_indicator<<<(unsigned)((nvt + 255) / 256), 256, 0, c->st>>>(c->crs_y, dcolor, cc, nvt);
    launch_coarse_prolong_add(dm, c->crs_y, c->pw[0], 0, c->st);
    if (apply_E(c, c->pw[0], c->pw[1], nullptr)) return 1;
    launch_coarse_restrict(dm, c->pw[1], nullptr, c->crs_part, c->crs_r, c->st);
    if (ctx_allreduce(c, c->crs_r, (int)nvt, false)) return 1;
    NLK_CUDA(cudaMemcpyAsync(r.data(), c->crs_r, sizeof(double) * nvt, cudaMemcpyDeviceToHost, c->st));
    NLK_CUDA(cudaStreamSynchronize(c->st));
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < nvt; ++w) {
      for (int32_t k = rowptr[w]; k < rowptr[w + 1]; ++k) if (color[col[k]] == cc) { val[k] = r[w]; break; }
    }
  }
  std::vector<double> dinv(nvt, 1.0);
  for (int64_t w = 0; w < nvt; ++w) for (int32_t k = rowptr[w]; k < rowptr[w + 1]; ++k) if (col[k] == w) dinv[w] = val[k] != 0.0 ? 1.0 / val[k] : 1.0;
  if (dev_upload(c, &c->crs_rowptr, rowptr) || dev_upload(c, &c->crs_col, col) || dev_upload(c, &c->crs_val, val) || dev_upload(c, &c->crs_dinv, dinv)) return 1;
  if (dev_alloc(c, &c->crs_p, nvt) || dev_alloc(c, &c->crs_q, nvt) || dev_alloc(c, &c->crs_z, nvt) || dev_alloc(c, &c->crs_rr, nvt)) return 1;
  if (dev_alloc(c, &c->crs_partial, 2 * CRS_MAXB_RZ + CRS_MAXB_PQ)) return 1;
  c->crs_nnz = nnz; c->coarse_sparse = true; c->have_coarse = true;
  if (c->crs_iters <= 0) c->crs_iters = 12;
  if (getenv("NLK_VERBOSE")) printf("[nlk] sparse coarse operator: %lld vertices, %lld nnz, %d colours, %d PCG iterations per apply\n", (long long)nvt, (long long)nnz, ncolors, c->crs_iters);
  return 0;
}

}  // namespace nlk
