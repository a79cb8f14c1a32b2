// EXPERIMENTAL, opt-in (environment NLK_CG3=1; single rank, streamed regime only): the Helmholtz Jacobi-PCG of all velocity
// components in lockstep, with ONE Helmholtz kernel per iteration that loads the six geometric factors and bm1 once for the
// three fields (12 -> ~7.3 streamed arrays per component and iteration, a third of the launches
// and host checks).  Written after the round-1 GPU budget was spent: it compiles, but it has NOT been run or measured on a
// B200 yet, which is why the default path (nlk_solver.cu::helmholtz_solve) does not use it.  Validate with
//     NLK_CG3=1 NLK_NO_CGP=1 python -m pytest tests -m gpu        and        NLK_CG3=1 python bench.py
// Same arithmetic as the per-component solver (Nek hmholtz + cggo): z = hd r, p = z + beta p, w = H p, alpha = r.z / p.w,
// x += alpha p, r -= alpha w, convergence on rbn2 = sqrt(sum r^2 mult binv / vol); fields that have converged are frozen.
#include "nlk_ctx.hpp"
#include <algorithm>
#include <cstdlib>

namespace nlk {

__constant__ double c3_D[16 * 16];        // GLL derivative matrix (constant memory is per translation unit)
static int c3_D_n = -1;

struct Cg3 {
  double* p[3]; double* r[3]; double* w[3]; double* x[3];
  const double* hd[3]; const double* wa[3]; const double* wb[3];
  int nf;
};

template <int N> __host__ __device__ constexpr int cg3_epb() { return N * N >= 100 ? 1 : (N * N >= 64 ? 2 : 4); }

// One element per threadIdx.y slice, thread (i, j) owns the k-column of every field in registers.
template <int N, int DIM>
__global__ void __launch_bounds__(N * N * cg3_epb<N>(), (N == 8 ? 3 : 1))
k_axhelm3(Cg3 a, const double* __restrict__ G, const double* __restrict__ bm1, const double* __restrict__ Dg, double h1, double h2,
          SolverScal* __restrict__ sc, int64_t E, double* __restrict__ pap_partial, unsigned int* pap_counter) {
  constexpr int NZ = DIM == 3 ? N : 1, NN = N * N, NP = NN * NZ, NG = DIM == 3 ? 6 : 3, EPB = cg3_epb<N>(), NT = NN * EPB;
  __shared__ double sD[NN], sDt[NN];
  __shared__ double s_u[EPB][NN], s_gr[EPB][NN], s_gs[EPB][NN];
  __shared__ int s_last;
  bool act[3];
#pragma unroll
  for (int f = 0; f < 3; ++f) act[f] = f < a.nf && !sc[f].done;                  // block-uniform
  if (!act[0] && !act[1] && !act[2]) return;
  const int tid = threadIdx.x, le = threadIdx.y, lt = le * NN + tid;
  const int64_t e = (int64_t)blockIdx.x * EPB + le;
  const int i = tid % N, j = tid / N;
  for (int idx = lt; idx < NN; idx += NT) { const double v = Dg[idx]; sD[idx] = v; sDt[(idx % N) * N + idx / N] = v; }
  const bool active = e < E;
  const size_t eb = (size_t)(active ? e : 0) * NP;
  double ru[3][NZ], rw[3][NZ];
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    const double beta = act[f] ? sc[f].beta : 0.0;
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      ru[f][k] = 0.0; rw[f][k] = 0.0;
      if (act[f]) {
        const size_t g = eb + k * NN + tid;
        const double pv = a.r[f][g] * a.hd[f][g] + beta * a.p[f][g];
        if (active) { ru[f][k] = pv; a.p[f][g] = pv; }
      }
    }
  }
  const double* Ge = G + (size_t)(active ? e : 0) * NG * NP + tid;
  double gn[NG];
#pragma unroll
  for (int c = 0; c < NG; ++c) gn[c] = Ge[c * NP];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NZ; ++k) {
    double gc[NG];
#pragma unroll
    for (int c = 0; c < NG; ++c) gc[c] = gn[c];
    if (k + 1 < NZ) {
#pragma unroll
      for (int c = 0; c < NG; ++c) gn[c] = Ge[c * NP + (k + 1) * NN];
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (!act[f]) continue;                                                    // uniform: the barriers below stay convergent
      s_u[le][tid] = ru[f][k];
      __syncthreads();
      double ur = 0, us = 0, ut = 0;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        ur += sDt[l * N + i] * s_u[le][j * N + l];
        us += sD[j * N + l] * s_u[le][l * N + i];
      }
      if (DIM == 3) {
#pragma unroll
        for (int l = 0; l < N; ++l) ut += c3_D[k * N + l] * ru[f][l];
      }
      double gr, gs, gt = 0;
      if (DIM == 3) {
        gr = gc[0] * ur + gc[3] * us + gc[4] * ut;
        gs = gc[3] * ur + gc[1] * us + gc[5] * ut;
        gt = gc[4] * ur + gc[5] * us + gc[2] * ut;
      } else {
        gr = gc[0] * ur + gc[2] * us;
        gs = gc[2] * ur + gc[1] * us;
      }
      s_gr[le][tid] = gr; s_gs[le][tid] = gs;
      __syncthreads();
      double acc = 0;
#pragma unroll
      for (int l = 0; l < N; ++l) acc += sD[l * N + i] * s_gr[le][j * N + l] + sD[l * N + j] * s_gs[le][l * N + i];
      rw[f][k] += acc;
      if (DIM == 3) {
#pragma unroll
        for (int l = 0; l < N; ++l) rw[f][l] += c3_D[k * N + l] * gt;
      }
      __syncthreads();
    }
  }
  double pap[3] = {0.0, 0.0, 0.0};
  if (active) {
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      const size_t g = eb + k * NN + tid;
      const double b = bm1[g];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        if (!act[f]) continue;
        const double wv = h1 * rw[f][k] + h2 * b * ru[f][k];
        a.w[f][g] = wv; pap[f] += ru[f][k] * wv;
      }
    }
  }
  // p.Hp of every field: block partials, the last block sums them in a fixed order and sets alpha
  double* sred = &s_gr[0][0];
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    if (!act[f]) continue;
    sred[lt] = pap[f];
    __syncthreads();
    if (lt < 32) { double t = 0; for (int q = lt; q < NT; q += 32) t += sred[q]; sred[lt] = t; }
    __syncthreads();
    if (lt == 0) { double t = 0; for (int q = 0; q < (NT < 32 ? NT : 32); ++q) t += sred[q]; pap_partial[(size_t)f * gridDim.x + blockIdx.x] = t; }
    __syncthreads();
  }
  if (lt == 0) { __threadfence(); s_last = (atomicAdd(pap_counter, 1u) == gridDim.x - 1); }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    if (!act[f]) continue;
    double t = 0;
    for (unsigned b = lt; b < gridDim.x; b += NT) t += __ldcg(&pap_partial[(size_t)f * gridDim.x + b]);
    sred[lt] = t;
    __syncthreads();
    if (lt == 0) {
      double tt = 0; for (int q = 0; q < NT; ++q) tt += sred[q];
      sc[f].red[2] = tt; sc[f].pap = tt; sc[f].alpha = sc[f].rtz1 / tt;
    }
    __syncthreads();
  }
  if (lt == 0) *pap_counter = 0u;
}

__global__ void k_cg3_init(SolverScal* sc, int nf, double tol, int maxit) {
  const int f = threadIdx.x;
  if (f >= nf) return;
  SolverScal& s = sc[f];
  s.rtz1 = 1.0; s.rtz2 = 1.0; s.rbn2 = 0; s.rbn0 = 0; s.pap = 0; s.alpha = 0; s.beta = 0; s.tol = tol; s.iter = 0; s.done = 0; s.maxit = maxit;
}

// blockIdx.y = field.  x += alpha p ; r -= alpha w (skipped when first) ; r.z = sum r^2 wa ; rbn2^2 = sum r^2 wb / vol ;
// the last block of each field applies cggo's convergence test and sets beta.
__global__ void __launch_bounds__(256)
k_cg3_update(Cg3 a, size_t n, double vol, SolverScal* sc, double* __restrict__ partial /*[3][2][gridDim.x]*/, unsigned int* counter /*[3]*/, int first) {
  const int f = blockIdx.y;
  SolverScal& s = sc[f];
  if (s.done) return;
  const double alpha = first ? 0.0 : s.alpha;
  double* __restrict__ x = a.x[f]; double* __restrict__ r = a.r[f];
  const double* __restrict__ p = a.p[f]; const double* __restrict__ w = a.w[f];
  const double* __restrict__ wa = a.wa[f]; const double* __restrict__ wb = a.wb[f];
  double v0 = 0.0, v1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double ri = r[i];
    if (!first) { x[i] += alpha * p[i]; ri -= alpha * w[i]; r[i] = ri; }
    const double r2 = ri * ri;
    v0 += r2 * wa[i]; v1 += r2 * wb[i];
  }
  __shared__ double sp[2][8];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_down_sync(0xffffffffu, v0, o); v1 += __shfl_down_sync(0xffffffffu, v1, o); }
  if (lane == 0) { sp[0][wid] = v0; sp[1][wid] = v1; }
  __syncthreads();
  double* part = partial + (size_t)f * 2 * gridDim.x;
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { t0 += sp[0][q]; t1 += sp[1][q]; }
    part[blockIdx.x] = t0; part[gridDim.x + blockIdx.x] = t1;
    __threadfence();
    s_last = (atomicAdd(&counter[f], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    double t0 = 0, t1 = 0;
    for (unsigned b = lane; b < gridDim.x; b += 32) { t0 += __ldcg(&part[b]); t1 += __ldcg(&part[gridDim.x + b]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_down_sync(0xffffffffu, t0, o); t1 += __shfl_down_sync(0xffffffffu, t1, o); }
    if (lane == 0) {
      s.red[0] = t0; s.red[1] = t1;
      s.rtz2 = s.rtz1; s.rtz1 = t0;                                         // cg_finalize_zr (nlk_kernels.cu), Nek cggo
      const double rbn2 = sqrt(t1 / vol);
      s.rbn2 = rbn2;
      if (s.iter == 0) s.rbn0 = rbn2;
      if (rbn2 <= s.tol || s.iter >= s.maxit || !(t0 == t0)) s.done = 1;
      else { s.beta = s.iter == 0 ? 0.0 : t0 / s.rtz2; s.iter += 1; }
      counter[f] = 0u;
    }
  }
}

template <int N, int DIM>
static void axhelm3_dispatch(const DevMesh& dm, const Cg3& a, double h1, double h2, SolverScal* sc, double* pap_partial, unsigned int* pap_counter, cudaStream_t st) {
  constexpr int EPB = cg3_epb<N>();
  dim3 block(N * N, EPB), grid((unsigned)((dm.E + EPB - 1) / EPB));
  k_axhelm3<N, DIM><<<grid, block, 0, st>>>(a, dm.G, dm.bm1, dm.D, h1, h2, sc, dm.E, pap_partial, pap_counter);
  ++g_launches;
}

struct Cg3State {          // lazily allocated, owned by the context (nlk_ctx::cg3, released by cg3_release at destruction)
  double* x[3] = {nullptr, nullptr, nullptr}; double* p[3] = {nullptr, nullptr, nullptr}; double* w[3] = {nullptr, nullptr, nullptr};
  SolverScal* d_sc = nullptr; SolverScal* h_sc = nullptr;
  double* pap_partial = nullptr; double* upd_partial = nullptr; unsigned int* counters = nullptr;
  int upd_grid = 0;
};
void cg3_release(nlk_ctx* c) {
  auto* S = static_cast<Cg3State*>(c->cg3);
  if (!S) return;
  if (S->h_sc) cudaFreeHost(S->h_sc);
  delete S;                                            // (device buffers are in c->allocs)
  c->cg3 = nullptr;
}

bool cg3_enabled() { static const bool on = getenv("NLK_CG3") != nullptr; return on; }

// sol[k] += (h1 A + h2 B)^-1 mask_k dssum(rhs[k]) for k < nf, all fields in lockstep; rhs arrays are overwritten (they become r)
int helmholtz_solve3(nlk_ctx* c, int nf, double* const* rhs, double h1, double h2, const double* const* masks, double tol, double* const* sol) {
  const DevMesh& dm = c->dm;
  if (nf < 1 || nf > 3 || c->nccl.nranks > 1) { set_error("helmholtz_solve3: 1..3 fields, single rank"); return 1; }
  if (!c->cg3) {
    auto* N = new Cg3State{};
    c->cg3 = N;
    for (int k = 0; k < 3; ++k) if (dev_alloc(c, &N->x[k], dm.N1) || dev_alloc(c, &N->p[k], dm.N1) || dev_alloc(c, &N->w[k], dm.N1)) return 1;
    N->upd_grid = (int)std::min<size_t>((dm.N1 + 255) / 256, 1184);
    if (dev_alloc(c, &N->d_sc, 3) || dev_alloc(c, &N->pap_partial, 3 * ((size_t)dm.E + 1)) || dev_alloc(c, &N->upd_partial, (size_t)6 * N->upd_grid) || dev_alloc(c, &N->counters, 4)) return 1;
    NLK_CUDA(cudaMallocHost((void**)&N->h_sc, 3 * sizeof(SolverScal)));
  }
  Cg3State& S = *static_cast<Cg3State*>(c->cg3);
  if (c3_D_n != dm.n) { NLK_CUDA(cudaMemcpyToSymbolAsync(c3_D, dm.D, sizeof(double) * dm.n * dm.n, 0, cudaMemcpyDeviceToDevice, c->st)); c3_D_n = dm.n; }
  Cg3 a{}; a.nf = nf;
  for (int k = 0; k < nf; ++k) {
    int slot = -1;
    if (cg_weights(c, masks[k], h1, h2, &slot)) return 1;
    a.hd[k] = c->cg_hd[slot]; a.wa[k] = c->cg_wa[slot]; a.wb[k] = c->cg_wb[slot];
    a.r[k] = rhs[k]; a.p[k] = S.p[k]; a.w[k] = S.w[k]; a.x[k] = S.x[k];
  }
  for (int k = nf; k < 3; ++k) { a.hd[k] = a.hd[0]; a.wa[k] = a.wa[0]; a.wb[k] = a.wb[0]; a.r[k] = a.r[0]; a.p[k] = a.p[0]; a.w[k] = a.w[0]; a.x[k] = a.x[0]; }
  Ptr3 rp{{rhs[0], nf > 1 ? rhs[1] : nullptr, nf > 2 ? rhs[2] : nullptr}};
  if (ctx_gs(c, rp, nf)) return 1;
  for (int k = 0; k < nf; ++k) {
    launch_lin(rhs[k], dm.N1, 1.0, rhs[k], 0, nullptr, 0, nullptr, 0, nullptr, masks[k], c->st);       // r = mask * dssum(rhs)
    NLK_CUDA(cudaMemsetAsync(S.x[k], 0, dm.N1 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(S.p[k], 0, dm.N1 * sizeof(double), c->st));
  }
  k_cg3_init<<<1, 32, 0, c->st>>>(S.d_sc, nf, tol, c->prm.cg_maxit); ++g_launches;
  const dim3 ugrid((unsigned)S.upd_grid, (unsigned)nf);
  k_cg3_update<<<ugrid, 256, 0, c->st>>>(a, dm.N1, dm.volvm1, S.d_sc, S.upd_partial, S.counters, 1); ++g_launches;
  Ptr3 wp{{S.w[0], nf > 1 ? S.w[1] : nullptr, nf > 2 ? S.w[2] : nullptr}};
  const int batch = 4;
  bool done = false;
  while (!done) {
    for (int b = 0; b < batch; ++b) {
      switch (dm.n * 10 + dm.ndim) {
#define C3(N_, D_) case N_ * 10 + D_: axhelm3_dispatch<N_, D_>(dm, a, h1, h2, S.d_sc, S.pap_partial, S.counters + 3, c->st); break;
        C3(4, 2) C3(4, 3) C3(5, 2) C3(5, 3) C3(6, 2) C3(6, 3) C3(7, 2) C3(7, 3) C3(8, 2) C3(8, 3) C3(9, 2) C3(9, 3) C3(10, 2) C3(10, 3) C3(11, 2) C3(12, 2)
#undef C3
        default: set_error("helmholtz_solve3: no instantiation for this lx1"); return 1;
      }
      if (ctx_gs(c, wp, nf)) return 1;
      k_cg3_update<<<ugrid, 256, 0, c->st>>>(a, dm.N1, dm.volvm1, S.d_sc, S.upd_partial, S.counters, 0); ++g_launches;
    }
    NLK_CUDA(cudaMemcpyAsync(S.h_sc, S.d_sc, 3 * sizeof(SolverScal), cudaMemcpyDeviceToHost, c->st));
    NLK_CUDA(cudaStreamSynchronize(c->st));
    done = true;
    for (int k = 0; k < nf; ++k) if (!S.h_sc[k].done && S.h_sc[k].iter < c->prm.cg_maxit) done = false;
  }
  for (int k = 0; k < nf; ++k) {
    launch_lin(sol[k], dm.N1, 1.0, sol[k], 1.0, S.x[k], 0, nullptr, 0, nullptr, nullptr, c->st);
    c->cg_iters += S.h_sc[k].iter;
  }
  return 0;
}

}  // namespace nlk
