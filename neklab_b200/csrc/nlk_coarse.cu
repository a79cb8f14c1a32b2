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

// scal layout: [0] rz, [1] pq, [2] alpha, [3] beta, [4] rz_new, [5] done
// q = A p (warp per row) and pq = p.q ; the last block finalises alpha = rz / pq
__global__ void __launch_bounds__(256)
k_crs_spmv_dot(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val, const double* __restrict__ p,
               double* __restrict__ q, int n, double* scal, double* partial, unsigned int* counter) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double acc = 0.0;
  for (int row = blockIdx.x * wpb + wib; row < n; row += gridDim.x * wpb) {
    double s = 0.0;
    for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) s += val[k] * p[col[k]];
    s = warp_sum_c(s);
    if (lane == 0) { q[row] = s; acc += s * p[row]; }
  }
  __shared__ double sp[8]; __shared__ int last;
  if (lane == 0) sp[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int i = 0; i < wpb; ++i) t += sp[i];
    partial[blockIdx.x] = t; __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double t = 0; for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&partial[b]);
    t = warp_sum_c(t);
    if (lane == 0) { scal[1] = t; scal[2] = (t != 0.0) ? scal[0] / t : 0.0; *counter = 0u; }
  }
}

// x += alpha p ; r -= alpha q ; z = dinv r ; rz_new = r.z ; last block: beta = rz_new / rz, rz = rz_new
__global__ void __launch_bounds__(256)
k_crs_update(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, const double* __restrict__ p, const double* __restrict__ q,
             const double* __restrict__ dinv, int n, double* scal, double* partial, unsigned int* counter, int first) {
  const double alpha = first ? 0.0 : scal[2];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double ri = r[i];
    if (!first) { x[i] += alpha * p[i]; ri -= alpha * q[i]; r[i] = ri; }
    double zi = ri * dinv[i]; z[i] = zi; acc += ri * zi;
  }
  acc = warp_sum_c(acc);
  __shared__ double sp[8]; __shared__ int last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  if (lane == 0) sp[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int i = 0; i < wpb; ++i) t += sp[i];
    partial[blockIdx.x] = t; __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double t = 0; for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&partial[b]);
    t = warp_sum_c(t);
    if (lane == 0) { scal[3] = (first || scal[0] == 0.0) ? 0.0 : t / scal[0]; scal[0] = t; *counter = 0u; }
  }
}
__global__ void k_crs_p(double* __restrict__ p, const double* __restrict__ z, int n, const double* scal) {
  const double beta = scal[3];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}

int coarse_solve_sparse(nlk_ctx* c, const double* rc, double* yc) {
  const int n = (int)c->dm.nvert;
  cudaStream_t st = c->st;
  const int gu = std::min((n + 255) / 256, 296), gs = std::min((n + 7) / 8, 1184);
  NLK_CUDA(cudaMemsetAsync(yc, 0, sizeof(double) * n, st));
  NLK_CUDA(cudaMemsetAsync(c->crs_p, 0, sizeof(double) * n, st));
  NLK_CUDA(cudaMemcpyAsync(c->crs_rr, rc, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  double* partial = c->red.partial; unsigned int* counter = c->red.counter;
  k_crs_update<<<gu, 256, 0, st>>>(yc, c->crs_rr, c->crs_z, c->crs_p, c->crs_q, c->crs_dinv, n, c->crs_scal, partial, counter, 1); ++g_launches;
  for (int it = 0; it < c->crs_iters; ++it) {
    k_crs_p<<<gu, 256, 0, st>>>(c->crs_p, c->crs_z, n, c->crs_scal); ++g_launches;
    k_crs_spmv_dot<<<gs, 256, 0, st>>>(c->crs_rowptr, c->crs_col, c->crs_val, c->crs_p, c->crs_q, n, c->crs_scal, partial, counter); ++g_launches;
    k_crs_update<<<gu, 256, 0, st>>>(yc, c->crs_rr, c->crs_z, c->crs_p, c->crs_q, c->crs_dinv, n, c->crs_scal, partial, counter, 0); ++g_launches;
  }
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
  if (dev_alloc(c, &c->crs_p, nvt) || dev_alloc(c, &c->crs_q, nvt) || dev_alloc(c, &c->crs_z, nvt) || dev_alloc(c, &c->crs_rr, nvt) || dev_alloc(c, &c->crs_scal, 8)) return 1;
  c->crs_nnz = nnz; c->coarse_sparse = true; c->have_coarse = true;
  if (c->crs_iters <= 0) c->crs_iters = 12;
  if (getenv("NLK_VERBOSE")) printf("[nlk] sparse coarse operator: %lld vertices, %lld nnz, %d colours, %d PCG iterations per apply\n", (long long)nvt, (long long)nnz, ncolors, c->crs_iters);
  return 0;
}

}  // namespace nlk
