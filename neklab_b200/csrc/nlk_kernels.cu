// Hand-written sm_100a FP64 CUDA kernels of the exptA hot path.
// All kernels are HBM/latency-bound integer+FP64 streaming or small tensor contractions: no tensor cores.
// Element data is staged in shared memory; 1-D operator matrices are staged in shared memory per CTA.
#include "nlk_device.cuh"
#include <cstdlib>

namespace nlk {

thread_local long g_launches = 0;
#define LAUNCH_COUNT() (++g_launches)

static inline int cdiv(size_t a, size_t b) { return (int)((a + b - 1) / b); }
// 4-way unrolled grid-stride loop: four independent (predicated) iterations per trip keep >= 4x the loads in flight
#define NLK_STREAM4_BEGIN(i, n)                                                                                        \
  for (size_t base_ = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; base_ < (n); base_ += (size_t)gridDim.x * blockDim.x * 4) { \
    _Pragma("unroll") for (int u_ = 0; u_ < 4; ++u_) {                                                                \
      const size_t i = base_ + (size_t)u_ * blockDim.x;                                                               \
      if (i < (n)) {
#define NLK_STREAM4_END }}}
static inline int stream_grid(size_t n, int cap) { int g = (int)((n + 1023) / 1024); return g < 1 ? 1 : (g > cap ? cap : g); }
static const int RED_BLOCKS = 1184;  // 148 SMs x 8
static const int RED_THREADS = 256;

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic two-stage reduction: every block writes its partial sums; the last block to finish (ticket
// counter) sums the partials in a fixed order.  Returns true on thread 0 of the last block with v[] = totals.
template <int NV, bool MAXOP = false>
__device__ bool grid_reduce(double (&v)[NV], Reducer red) {
  __shared__ double s_part[NV][32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = MAXOP ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0) s_part[k][wid] = t;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nw ? s_part[k][lane] : (MAXOP ? -1e300 : 0.0);
      t = MAXOP ? warp_max(t) : warp_sum(t);
      if (lane == 0) red.partial[(size_t)k * red.maxblocks + blockIdx.x] = t;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int ticket = atomicAdd(red.counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double acc = MAXOP ? -1e300 : 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      double t = __ldcg(&red.partial[(size_t)k * red.maxblocks + b]);
      acc = MAXOP ? fmax(acc, t) : acc + t;
    }
    acc = MAXOP ? warp_max(acc) : warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_part[k][wid] = acc;
    __syncthreads();
    if (wid == 0) {
      double t = lane < nw ? s_part[k][lane] : (MAXOP ? -1e300 : 0.0);
      t = MAXOP ? warp_max(t) : warp_sum(t);
      v[k] = t;
    }
  }
  if (threadIdx.x == 0) *red.counter = 0u;
  return threadIdx.x == 0;
}

// ------------------------------------------------------------------------------------------------ K1 axhelm
// GLL derivative matrix in constant memory for the warp-uniform accesses of the t-direction (k is a compile-time loop
// index): constant-cache broadcasts instead of shared-memory wavefronts (ncu r01: the LSU/shared pipe was the limiter).
__constant__ double c_D[16 * 16];
static int c_D_n = -1;
static void ensure_const_D(const DevMesh& dm, cudaStream_t st) {
  if (c_D_n == dm.n) return;
  cudaMemcpyToSymbolAsync(c_D, dm.D, sizeof(double) * dm.n * dm.n, 0, cudaMemcpyDeviceToDevice, st);
  c_D_n = dm.n;
}
__device__ __forceinline__ void cg_finalize_pap(SolverScal* sc);
// One element per (threadIdx.y) slice; thread (i,j) owns the k-column of the element (register-tiled k-loop, 3-D).
// FUSE_CG: the CG search-direction update p = hd*r + beta*p (hd: masked inverse Jacobi diagonal) is applied while loading, and p.Ap is reduced
// on the way out: p is continuous and masked, so the sum over local copies of p * w(unassembled) equals the assembled,
// multiplicity-weighted inner product Nek's cggo takes after dssum -- one full pass over w, p, mask and mult saved per
// iteration.  Block partials -> last block sums them in a fixed order (deterministic).
template <int N, int DIM, bool FUSE_CG>
__global__ void __launch_bounds__(N * N * (N * N >= 100 ? 1 : (N * N >= 64 ? 2 : 4)), (N == 8 ? 8 : 1))
k_axhelm(const double* __restrict__ u, double* __restrict__ pio, const double* __restrict__ r, double* __restrict__ w,
         const double* __restrict__ G, const double* __restrict__ bm1, const double* __restrict__ Dg,
         const double* __restrict__ hd, double h1, double h2,
         SolverScal* __restrict__ sc, int64_t E, double* __restrict__ pap_partial, unsigned int* pap_counter, int defer) {
  constexpr int NZ = DIM == 3 ? N : 1;
  constexpr int NN = N * N;
  constexpr int NP = NN * NZ;
  constexpr int NG = DIM == 3 ? 6 : 3;
  constexpr int EPB = NN >= 100 ? 1 : (NN >= 64 ? 2 : 4);
  __shared__ double sD[NN], sDt[NN];
  __shared__ double s_u[EPB][NN], s_gr[EPB][NN], s_gs[EPB][NN];
  if (FUSE_CG && sc->done) return;
  const int tid = threadIdx.x, le = threadIdx.y;
  const int64_t e = (int64_t)blockIdx.x * EPB + le;
  const int i = tid % N, j = tid / N;
  for (int idx = tid + le * NN; idx < NN; idx += NN * EPB) { double v = Dg[idx]; sD[idx] = v; sDt[(idx % N) * N + idx / N] = v; }
  const bool active = e < E;
  double ru[NZ], rw[NZ];
  const size_t eb = (size_t)(active ? e : 0) * NP;
  if (FUSE_CG) {
    const double beta = sc->beta;
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      size_t g = eb + k * NN + tid;
      double z = r[g] * hd[g];                    // hd = mask / (h1 diagA + h2 diagB), precomputed per (h1, h2) (k_cg_weights)
      double pv = z + beta * pio[g];
      ru[k] = active ? pv : 0.0;
      if (active) pio[g] = pv;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NZ; ++k) ru[k] = active ? u[eb + k * NN + tid] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < NZ; ++k) rw[k] = 0.0;
  // geometric factors of slice k are fetched one iteration ahead (software prefetch: more HBM bytes in flight per warp)
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
    s_u[le][tid] = ru[k];
    __syncthreads();
    double ur = 0, us = 0, ut = 0;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      ur += sDt[l * N + i] * s_u[le][j * N + l];
      us += sD[j * N + l] * s_u[le][l * N + i];
    }
    if (DIM == 3) {
#pragma unroll
      for (int l = 0; l < N; ++l) ut += c_D[k * N + l] * ru[l];
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
    rw[k] += acc;
    if (DIM == 3) {
#pragma unroll
      for (int l = 0; l < N; ++l) rw[l] += c_D[k * N + l] * gt;
    }
    __syncthreads();
  }
  double pap = 0.0;
  if (active) {
#pragma unroll
    for (int k = 0; k < NZ; ++k) { size_t g = eb + k * NN + tid; double wv = h1 * rw[k] + h2 * bm1[g] * ru[k]; w[g] = wv; if (FUSE_CG) pap += ru[k] * wv; }
  }
  if (FUSE_CG) {
    constexpr int NT = NN * EPB;
    double* sred = &s_gr[0][0];                     // free after the last barrier of the k-loop
    __shared__ int s_last;
    const int lt = le * NN + tid;
    sred[lt] = pap;
    __syncthreads();
    if (lt < 32) { double t = 0; for (int q = lt; q < NT; q += 32) t += sred[q]; sred[lt] = t; }
    __syncthreads();
    if (lt == 0) {
      double t = 0; for (int q = 0; q < (NT < 32 ? NT : 32); ++q) t += sred[q];
      pap_partial[blockIdx.x] = t; __threadfence();
      s_last = (atomicAdd(pap_counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      double t = 0; for (unsigned b = lt; b < gridDim.x; b += NT) t += __ldcg(&pap_partial[b]);
      sred[lt] = t;
      __syncthreads();
      if (lt == 0) {
        double tt = 0; for (int q = 0; q < NT; ++q) tt += sred[q];
        sc->red[2] = tt; if (!defer) cg_finalize_pap(sc);
        *pap_counter = 0u;
      }
    }
  }
}

// lx1 = 8, 3-D: the same kernel with the six geometric factors of the element staged through shared memory by cp.async
// (LDGSTS), one commit group per k-slice, all issued before the first contraction.  ncu r02 on k_axhelm<8,3,1>: 59 % of the
// samples stalled on the long scoreboard, DRAM at 51 % -- the register prefetch of ONE slice ahead keeps ~49 KB per SM in
// flight, at the edge of what the HBM latency-bandwidth product needs, and the 64-register cap spilled.  Here ~300 KB per SM
// are in flight from the first instruction on and the k-loop never waits on global memory again.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_le(int pending) {      // wait until at most `pending` groups are in flight
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}
// REGD: the four rows/columns of D a thread needs (D[i][.], D[j][.], D[.][i], D[.][j]) live in registers for the whole k-loop,
// so that a contraction FMA costs ONE shared-memory operand instead of two.  ncu r02 (after the cp.async staging): the LSU
// data pipe of this kernel ran at 64 % of its peak (1 913 shared-memory wavefronts per element), short scoreboard + MIO
// throttle = 56 % of the stall samples -- operand delivery, not DRAM, was what kept it at 67 % of the HBM roof.
template <bool FUSE_CG, bool REGD>
__global__ void __launch_bounds__(128, REGD ? 3 : 4)
k_axhelm8a(const double* __restrict__ u, double* __restrict__ pio, const double* __restrict__ r, double* __restrict__ w,
           const double* __restrict__ G, const double* __restrict__ bm1, const double* __restrict__ Dg,
           const double* __restrict__ hd, double h1, double h2,
           SolverScal* __restrict__ sc, int64_t E, double* __restrict__ pap_partial, unsigned int* pap_counter, int defer) {
  constexpr int N = 8, NZ = 8, NN = 64, NP = 512, NG = 6, EPB = 2;
  extern __shared__ __align__(16) double sG[];              // [EPB][NG][NP]
  __shared__ double sD[NN], sDt[NN];
  __shared__ double s_u[EPB][NN], s_gr[EPB][NN], s_gs[EPB][NN];
  if (FUSE_CG && sc->done) return;
  const int tid = threadIdx.x, le = threadIdx.y;
  const int64_t e = (int64_t)blockIdx.x * EPB + le;
  const int i = tid % N, j = tid / N;
  const bool active = e < E;
  const size_t eb = (size_t)(active ? e : 0) * NP;
  // ---- stage the geometric factors: slice k = 6 arrays x 64 doubles = 192 sixteen-byte chunks, 3 per thread
  {
    const double* Ge = G + (size_t)(active ? e : 0) * NG * NP;
    double* sGe = sG + (size_t)le * NG * NP;
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        const int chunk = tid + NN * c3, a = chunk >> 5, off = a * NP + k * NN + (chunk & 31) * 2;
        cp_async16(sGe + off, Ge + off);
      }
      cp_async_commit();
    }
  }
  for (int idx = tid + le * NN; idx < NN; idx += NN * EPB) { double v = Dg[idx]; sD[idx] = v; sDt[(idx % N) * N + idx / N] = v; }
  double ru[NZ], rw[NZ], bmv[NZ];
  if (FUSE_CG) {
    const double beta = sc->beta;
    double rv[NZ], hv[NZ], pv[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) { const size_t g = eb + k * NN + tid; rv[k] = r[g]; hv[k] = hd[g]; pv[k] = pio[g]; }
#pragma unroll
    for (int k = 0; k < NZ; ++k) bmv[k] = bm1[eb + k * NN + tid];
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      const double p_ = rv[k] * hv[k] + beta * pv[k];       // hd = mask / (h1 diagA + h2 diagB), precomputed per (h1, h2) (k_cg_weights)
      ru[k] = active ? p_ : 0.0;
      if (active) pio[eb + k * NN + tid] = p_;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NZ; ++k) { ru[k] = active ? u[eb + k * NN + tid] : 0.0; bmv[k] = bm1[eb + k * NN + tid]; }
  }
#pragma unroll
  for (int k = 0; k < NZ; ++k) rw[k] = 0.0;
  const double* sGe = sG + (size_t)le * NG * NP + tid;
  double dI[N], dJ[N], tI[N], tJ[N];                        // D[i][l], D[j][l], D[l][i], D[l][j]
  if (REGD) {
    __syncthreads();                                        // sD / sDt written above
#pragma unroll
    for (int l = 0; l < N; ++l) { dI[l] = sD[i * N + l]; dJ[l] = sD[j * N + l]; tI[l] = sDt[i * N + l]; tJ[l] = sDt[j * N + l]; }
  }
#pragma unroll
  for (int k = 0; k < NZ; ++k) {
    s_u[le][tid] = ru[k];
    cp_async_wait_le(NZ - 1 - k);
    __syncthreads();
    double gc[NG];
#pragma unroll
    for (int c = 0; c < NG; ++c) gc[c] = sGe[c * NP + k * NN];
    double ur = 0, us = 0, ut = 0;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      ur += (REGD ? dI[l] : sDt[l * N + i]) * s_u[le][j * N + l];
      us += (REGD ? dJ[l] : sD[j * N + l]) * s_u[le][l * N + i];
    }
#pragma unroll
    for (int l = 0; l < N; ++l) ut += c_D[k * N + l] * ru[l];
    const double gr = gc[0] * ur + gc[3] * us + gc[4] * ut;
    const double gs = gc[3] * ur + gc[1] * us + gc[5] * ut;
    const double gt = gc[4] * ur + gc[5] * us + gc[2] * ut;
    s_gr[le][tid] = gr; s_gs[le][tid] = gs;
    __syncthreads();
    double acc = 0;
#pragma unroll
    for (int l = 0; l < N; ++l) acc += (REGD ? tI[l] : sD[l * N + i]) * s_gr[le][j * N + l] + (REGD ? tJ[l] : sD[l * N + j]) * s_gs[le][l * N + i];
    rw[k] += acc;
#pragma unroll
    for (int l = 0; l < N; ++l) rw[l] += c_D[k * N + l] * gt;
    __syncthreads();
  }
  double pap = 0.0;
  if (active) {
#pragma unroll
    for (int k = 0; k < NZ; ++k) { const size_t g = eb + k * NN + tid; const double wv = h1 * rw[k] + h2 * bmv[k] * ru[k]; w[g] = wv; if (FUSE_CG) pap += ru[k] * wv; }
  }
  if (FUSE_CG) {
    constexpr int NT = NN * EPB;
    double* sred = &s_gr[0][0];                     // free after the last barrier of the k-loop
    __shared__ int s_last;
    const int lt = le * NN + tid;
    sred[lt] = pap;
    __syncthreads();
    if (lt < 32) { double t = 0; for (int q = lt; q < NT; q += 32) t += sred[q]; sred[lt] = t; }
    __syncthreads();
    if (lt == 0) {
      double t = 0; for (int q = 0; q < 32; ++q) t += sred[q];
      pap_partial[blockIdx.x] = t; __threadfence();
      s_last = (atomicAdd(pap_counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      double t = 0; for (unsigned b = lt; b < gridDim.x; b += NT) t += __ldcg(&pap_partial[b]);
      sred[lt] = t;
      __syncthreads();
      if (lt == 0) {
        double tt = 0; for (int q = 0; q < NT; ++q) tt += sred[q];
        sc->red[2] = tt; if (!defer) cg_finalize_pap(sc);
        *pap_counter = 0u;
      }
    }
  }
}

template <int N, int DIM>
static void axhelm_dispatch(const DevMesh& dm, const double* u, double* pio, const double* r, double* w, const double* hd, double h1, double h2,
                            SolverScal* sc, bool fuse, double* pap_partial, unsigned int* pap_counter, int defer, cudaStream_t st) {
  constexpr int EPB = N * N >= 100 ? 1 : (N * N >= 64 ? 2 : 4);
  ensure_const_D(dm, st);
  dim3 block(N * N, EPB), grid(cdiv(dm.E, EPB));
  if constexpr (N == 8 && DIM == 3) {
    static const bool no_async = getenv("NLK_NO_AXASYNC") != nullptr;
    if (!no_async) {
      constexpr size_t smem = (size_t)EPB * 6 * 512 * sizeof(double);
      static bool attr = false;
      static const bool lds_d = getenv("NLK_AX8_LDSD") != nullptr;      // A/B switch: D operands from shared memory (the r02 first version)
      if (!attr) {
        cudaFuncSetAttribute(k_axhelm8a<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cudaFuncSetAttribute(k_axhelm8a<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_axhelm8a<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cudaFuncSetAttribute(k_axhelm8a<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
      }
      if (lds_d) {
        if (fuse) k_axhelm8a<true, false><<<grid, block, smem, st>>>(u, pio, r, w, dm.G, dm.bm1, dm.D, hd, h1, h2, sc, dm.E, pap_partial, pap_counter, defer);
        else k_axhelm8a<false, false><<<grid, block, smem, st>>>(u, pio, r, w, dm.G, dm.bm1, dm.D, nullptr, h1, h2, nullptr, dm.E, nullptr, nullptr, 0);
      } else {
        if (fuse) k_axhelm8a<true, true><<<grid, block, smem, st>>>(u, pio, r, w, dm.G, dm.bm1, dm.D, hd, h1, h2, sc, dm.E, pap_partial, pap_counter, defer);
        else k_axhelm8a<false, true><<<grid, block, smem, st>>>(u, pio, r, w, dm.G, dm.bm1, dm.D, nullptr, h1, h2, nullptr, dm.E, nullptr, nullptr, 0);
      }
      LAUNCH_COUNT(); return;
    }
  }
  if (fuse) k_axhelm<N, DIM, true><<<grid, block, 0, st>>>(u, pio, r, w, dm.G, dm.bm1, dm.D, hd, h1, h2, sc, dm.E, pap_partial, pap_counter, defer);
  else k_axhelm<N, DIM, false><<<grid, block, 0, st>>>(u, pio, r, w, dm.G, dm.bm1, dm.D, nullptr, h1, h2, nullptr, dm.E, nullptr, nullptr, 0);
  LAUNCH_COUNT();
}

#define NLK_FOR_N(FN, ...)                                                    \
  switch (dm.n * 10 + dm.ndim) {                                              \
    case 42: FN<4, 2>(__VA_ARGS__); break;   case 43: FN<4, 3>(__VA_ARGS__); break;   \
    case 52: FN<5, 2>(__VA_ARGS__); break;   case 53: FN<5, 3>(__VA_ARGS__); break;   \
    case 62: FN<6, 2>(__VA_ARGS__); break;   case 63: FN<6, 3>(__VA_ARGS__); break;   \
    case 72: FN<7, 2>(__VA_ARGS__); break;   case 73: FN<7, 3>(__VA_ARGS__); break;   \
    case 82: FN<8, 2>(__VA_ARGS__); break;   case 83: FN<8, 3>(__VA_ARGS__); break;   \
    case 92: FN<9, 2>(__VA_ARGS__); break;   case 93: FN<9, 3>(__VA_ARGS__); break;   \
    case 102: FN<10, 2>(__VA_ARGS__); break; case 103: FN<10, 3>(__VA_ARGS__); break; \
    case 112: FN<11, 2>(__VA_ARGS__); break; case 113: FN<11, 3>(__VA_ARGS__); break; \
    case 122: FN<12, 2>(__VA_ARGS__); break; case 123: FN<12, 3>(__VA_ARGS__); break; \
    default: break;                                                           \
  }

void launch_axhelm(const DevMesh& dm, const double* u, double* w, double h1, double h2, cudaStream_t st) {
  NLK_FOR_N(axhelm_dispatch, dm, u, nullptr, nullptr, w, nullptr, h1, h2, nullptr, false, nullptr, nullptr, 0, st)
}
void launch_axhelm_cg(const DevMesh& dm, double* p, const double* r, double* w, const double* hd, double h1, double h2, SolverScal* sc,
                      double* pap_partial, unsigned int* pap_counter, int defer, cudaStream_t st) {
  NLK_FOR_N(axhelm_dispatch, dm, p, p, r, w, hd, h1, h2, sc, true, pap_partial, pap_counter, defer, st)
}

// ------------------------------------------------------------------------------------------------ K2 dssum
// Direct-stiffness sum over the shared nodes, one thread per group.  The groups are split by multiplicity at set-up:
//   pairs (face-interior nodes, three quarters of all groups)  -> one int2 load gives both local indices,
//   quadruples (edge-interior nodes of a conforming hex mesh)  -> one int4 load,
//   the rest (vertices, irregular edges)                        -> CSR.
// All index loads, then all value loads of a group are issued together (the r01 kernel chained offset -> index -> value loads
// per copy).  Copies are always summed in ascending local order: deterministic and identical on every copy.
__global__ void __launch_bounds__(128)
k_gs(Ptr3 f, const int2* __restrict__ g2, int n2, const int4* __restrict__ g4, int n4, const int32_t* __restrict__ off, const int32_t* __restrict__ idx, int nr) {
  // one grid row (blockIdx.y) per field: the rows are scheduled one after the other, so each field's surface lines stay in L2
  // between their read and their write (three fields per thread measured 186 us against 3 x 51 us, ncu: 927 vs 660 MB of DRAM)
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  double* const u = blockIdx.y == 0 ? f.p[0] : (blockIdx.y == 1 ? f.p[1] : f.p[2]);
  if (g < n2) {
    const int2 ii = g2[g];
    const double a0 = u[ii.x], a1 = u[ii.y];
    const double sa = a0 + a1; u[ii.x] = sa; u[ii.y] = sa;
  } else if (g < n2 + n4) {
    const int4 ii = g4[g - n2];
    const double a0 = u[ii.x], a1 = u[ii.y], a2 = u[ii.z], a3 = u[ii.w];
    const double sa = ((a0 + a1) + a2) + a3; u[ii.x] = sa; u[ii.y] = sa; u[ii.z] = sa; u[ii.w] = sa;
  } else if (g < n2 + n4 + nr) {
    const int gg = g - n2 - n4;
    const int b = off[gg], e = off[gg + 1];
    double s = 0;
    for (int t = b; t < e; ++t) s += u[idx[t]];
    for (int t = b; t < e; ++t) u[idx[t]] = s;
  }
}
void launch_gs(const DevMesh& dm, Ptr3 f, int nf, cudaStream_t st) {
  if (dm.ngs == 0) return;
  const int nt = dm.ngs2 + dm.ngs4 + dm.ngsr;
  const int2* g2 = reinterpret_cast<const int2*>(dm.gs2); const int4* g4 = reinterpret_cast<const int4*>(dm.gs4);
  k_gs<<<dim3(cdiv(nt, 128), nf), 128, 0, st>>>(f, g2, dm.ngs2, g4, dm.ngs4, dm.gsr_off, dm.gsr_idx, dm.ngsr);
  LAUNCH_COUNT();
}

__global__ void k_pack(const double* __restrict__ u, const int32_t* __restrict__ rep, int cnt, double* __restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) buf[t] = u[rep[t]];
}
void launch_pack(const double* u, const int32_t* rep, int cnt, double* buf, cudaStream_t st) {
  if (!cnt) return;
  k_pack<<<cdiv(cnt, 128), 128, 0, st>>>(u, rep, cnt, buf); LAUNCH_COUNT();
}
__global__ void k_unpack_add(double* __restrict__ u, const int32_t* __restrict__ off, const int32_t* __restrict__ idx, int cnt,
                             const double* __restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  double v = buf[t];
  for (int s = off[t]; s < off[t + 1]; ++s) u[idx[s]] += v;
}
void launch_unpack_add(double* u, const int32_t* off, const int32_t* idx, int cnt, const double* buf, cudaStream_t st) {
  if (!cnt) return;
  k_unpack_add<<<cdiv(cnt, 128), 128, 0, st>>>(u, off, idx, cnt, buf); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ pointwise
__global__ void k_lin(double* __restrict__ out, size_t n, double a0, const double* __restrict__ x0, double a1, const double* __restrict__ x1,
                      double a2, const double* __restrict__ x2, double a3, const double* __restrict__ x3, const double* __restrict__ mul) {
  NLK_STREAM4_BEGIN(i, n)
    double v = 0;
    if (x0) v += a0 * x0[i];
    if (x1) v += a1 * x1[i];
    if (x2) v += a2 * x2[i];
    if (x3) v += a3 * x3[i];
    if (mul) v *= mul[i];
    out[i] = v;
    NLK_STREAM4_END
}
void launch_lin(double* out, size_t n, double a0, const double* x0, double a1, const double* x1, double a2, const double* x2, double a3,
                const double* x3, const double* mul, cudaStream_t st) {
  if (!n) return;
  int grid = stream_grid(n, 148 * 16);
  k_lin<<<grid, 256, 0, st>>>(out, n, a0, x0, a1, x1, a2, x2, a3, x3, mul); LAUNCH_COUNT();
}
__global__ void k_axpy_mm(double* __restrict__ out, size_t n, const double* __restrict__ x0, double a, const double* __restrict__ x1,
                          const double* __restrict__ m1, const double* __restrict__ m2) {
  NLK_STREAM4_BEGIN(i, n)
    double v = a * x1[i] * m1[i];
    if (m2) v *= m2[i];
    out[i] = (x0 ? x0[i] : 0.0) + v;
    NLK_STREAM4_END
}
void launch_axpy_mm(double* out, size_t n, const double* x0, double a, const double* x1, const double* m1, const double* m2, cudaStream_t st) {
  k_axpy_mm<<<stream_grid(n, 148 * 16), 256, 0, st>>>(out, n, x0, a, x1, m1, m2); LAUNCH_COUNT();
}
__global__ void k_fill(double* out, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
void launch_fill(double* out, size_t n, double v, cudaStream_t st) {
  if (!n) return;
  k_fill<<<std::min(cdiv(n, 256), 148 * 16), 256, 0, st>>>(out, n, v); LAUNCH_COUNT();
}
__global__ void k_sub_mean(double* p, size_t n, const double* sum, double inv_count) {
  const double mean = *sum * inv_count;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] -= mean;
}
void launch_sub_mean(double* p, size_t n, const double* sum, double inv_count, cudaStream_t st) {
  k_sub_mean<<<std::min(cdiv(n, 256), 148 * 16), 256, 0, st>>>(p, n, sum, inv_count); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ generic smem contraction
// out = M (mo x mi, row-major) applied along direction dir of `in` with dims (n0 x fastest, n1, n2); block-strided.
template <bool ACC>
__device__ __forceinline__ void contract(double* __restrict__ out, const double* __restrict__ in, const double* __restrict__ M, int mo,
                                         int mi, int dir, int n0, int n1, int n2) {
  const int o0 = dir == 0 ? mo : n0, o1 = dir == 1 ? mo : n1, o2 = dir == 2 ? mo : n2;
  const int tot = o0 * o1 * o2;
  const int stride = dir == 0 ? 1 : (dir == 1 ? n0 : n0 * n1);
  for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
    const int i = idx % o0, j = (idx / o0) % o1, k = idx / (o0 * o1);
    const int r = dir == 0 ? i : (dir == 1 ? j : k);
    const int base = (dir == 0 ? 0 : i) + n0 * ((dir == 1 ? 0 : j) + n1 * (dir == 2 ? 0 : k));
    const double* Mr = M + r * mi;
    double s = 0;
    for (int l = 0; l < mi; ++l) s += Mr[l] * in[base + l * stride];
    if (ACC) out[idx] += s; else out[idx] = s;
  }
  __syncthreads();
}
__device__ __forceinline__ void load_mat(double* s, const double* g, int cnt) {
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) s[i] = g[i];
}

// ------------------------------------------------------------------------------------------------ K5 opdiv / opgradt
// p = scale * sum_c sum_k rxw2[k][c] * (d_k u_c)|GL     (Nek multd / opdiv); one element per CTA.
__global__ void k_opdiv(CPtr3 u, double* __restrict__ p, const double* __restrict__ rxw2, const double* __restrict__ I12g,
                        const double* __restrict__ D12g, int n, int q, int d, double scale, const double* __restrict__ in_mul, CPtr3 in_mask,
                        const double* __restrict__ out_mul) {
  extern __shared__ double sm[];
  const int nz = d == 3 ? n : 1, qz = d == 3 ? q : 1;
  const int np1 = n * n * nz, np2 = q * q * qz;
  double* sI = sm; double* sDm = sI + q * n;
  double* U = sDm + q * n; double* A = U + np1; double* B = A + np1; double* T0 = B + np1; double* T1 = T0 + np1; double* T2 = T1 + np1;
  double* acc = T2 + np1;
  const size_t e = blockIdx.x;
  load_mat(sI, I12g, q * n); load_mat(sDm, D12g, q * n);
  for (int i = threadIdx.x; i < np2; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  for (int c = 0; c < d; ++c) {
    const double* uc = u.p[c] + e * np1;
    if (in_mul) { const double* mk = in_mask.p[c] + e * np1; const double* bi = in_mul + e * np1; for (int i = threadIdx.x; i < np1; i += blockDim.x) U[i] = uc[i] * bi[i] * mk[i]; }
    else for (int i = threadIdx.x; i < np1; i += blockDim.x) U[i] = uc[i];
    __syncthreads();
    contract<false>(A, U, sDm, q, n, 0, n, n, nz);      // D_x u   (q,n,nz)
    contract<false>(B, U, sI, q, n, 0, n, n, nz);       // I_x u
    const double* rw = rxw2 + e * (size_t)(d * d) * np2;
    if (d == 2) {
      contract<false>(T0, A, sI, q, n, 1, q, n, 1);     // d/dr
      contract<false>(T1, B, sDm, q, n, 1, q, n, 1);    // d/ds
      for (int i = threadIdx.x; i < np2; i += blockDim.x) acc[i] += rw[(0 * d + c) * np2 + i] * T0[i] + rw[(1 * d + c) * np2 + i] * T1[i];
      __syncthreads();
    } else {
      contract<false>(T0, A, sI, q, n, 1, q, n, n);     // I_y D_x u  (q,q,n)
      contract<false>(T1, B, sDm, q, n, 1, q, n, n);    // D_y I_x u
      contract<false>(T2, B, sI, q, n, 1, q, n, n);     // I_y I_x u
      contract<false>(A, T0, sI, q, n, 2, q, q, n);     // d/dr
      contract<false>(B, T1, sI, q, n, 2, q, q, n);     // d/ds
      contract<false>(U, T2, sDm, q, n, 2, q, q, n);    // d/dt
      for (int i = threadIdx.x; i < np2; i += blockDim.x)
        acc[i] += rw[(0 * d + c) * np2 + i] * A[i] + rw[(1 * d + c) * np2 + i] * B[i] + rw[(2 * d + c) * np2 + i] * U[i];
      __syncthreads();
    }
  }
  if (out_mul) for (int i = threadIdx.x; i < np2; i += blockDim.x) p[e * np2 + i] = scale * acc[i] * out_mul[e * np2 + i];
  else for (int i = threadIdx.x; i < np2; i += blockDim.x) p[e * np2 + i] = scale * acc[i];
}
static int elem_threads(int np) { int t = ((np + 31) / 32) * 32; return t > 256 ? 256 : (t < 64 ? 64 : t); }

void launch_opdiv_fused(const DevMesh& dm, CPtr3 u, double* p, double scale, const double* in_mul, const double* out_mul, cudaStream_t st) {
  if (tp_opdiv(dm, u, p, scale, in_mul, out_mul, st)) return;
  size_t smem = (size_t)(2 * dm.q * dm.n + 7 * dm.np1) * sizeof(double);
  static size_t set = 0;
  if (smem > 48 * 1024 && smem > set) { cudaFuncSetAttribute(k_opdiv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = smem; }
  CPtr3 mk{{dm.mask[0], dm.mask[1], dm.mask[2]}};
  k_opdiv<<<(unsigned)dm.E, elem_threads(dm.np1), smem, st>>>(u, p, dm.rxw2, dm.I12, dm.D12, dm.n, dm.q, dm.ndim, scale, in_mul, mk, out_mul);
  LAUNCH_COUNT();
}
void launch_opdiv(const DevMesh& dm, CPtr3 u, double* p, double scale, cudaStream_t st) { launch_opdiv_fused(dm, u, p, scale, nullptr, nullptr, st); }

// w_c = sum_k (transposed interp/deriv)[ p * rxw2[k][c] ]    (Nek cdtp / opgradt)
__global__ void k_opgradt(const double* __restrict__ p, Ptr3 w, const double* __restrict__ rxw2, const double* __restrict__ I12tg,
                          const double* __restrict__ D12tg, int n, int q, int d) {
  extern __shared__ double sm[];
  const int nz = d == 3 ? n : 1, qz = d == 3 ? q : 1;
  const int np1 = n * n * nz, np2 = q * q * qz;
  double* sIt = sm; double* sDt = sIt + q * n;
  double* P = sDt + q * n; double* S0 = P + np1; double* S1 = S0 + np1; double* S2 = S1 + np1; double* A0 = S2 + np1; double* A1 = A0 + np1;
  double* A2 = A1 + np1;
  const size_t e = blockIdx.x;
  load_mat(sIt, I12tg, q * n); load_mat(sDt, D12tg, q * n);
  for (int i = threadIdx.x; i < np2; i += blockDim.x) P[i] = p[e * np2 + i];
  __syncthreads();
  const double* rw = rxw2 + e * (size_t)(d * d) * np2;
  for (int c = 0; c < d; ++c) {
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
      S0[i] = P[i] * rw[(0 * d + c) * np2 + i];
      S1[i] = P[i] * rw[(1 * d + c) * np2 + i];
      if (d == 3) S2[i] = P[i] * rw[(2 * d + c) * np2 + i];
    }
    __syncthreads();
    double* wc = w.p[c] + e * np1;
    if (d == 2) {
      contract<false>(A0, S0, sDt, n, q, 0, q, q, 1);   // D^T_x  (n,q)
      contract<false>(A1, S1, sIt, n, q, 0, q, q, 1);   // I^T_x
      contract<false>(S0, A0, sIt, n, q, 1, n, q, 1);   // I^T_y
      contract<true>(S0, A1, sDt, n, q, 1, n, q, 1);    // + D^T_y
      for (int i = threadIdx.x; i < np1; i += blockDim.x) wc[i] = S0[i];
      __syncthreads();
    } else {
      contract<false>(A0, S0, sDt, n, q, 0, q, q, q);   // (n,q,q)
      contract<false>(A1, S1, sIt, n, q, 0, q, q, q);
      contract<false>(A2, S2, sIt, n, q, 0, q, q, q);
      contract<false>(S0, A0, sIt, n, q, 1, n, q, q);   // (n,n,q)
      contract<true>(S0, A1, sDt, n, q, 1, n, q, q);
      contract<false>(S1, A2, sIt, n, q, 1, n, q, q);
      contract<false>(A0, S0, sIt, n, q, 2, n, n, q);   // (n,n,n)
      contract<true>(A0, S1, sDt, n, q, 2, n, n, q);
      for (int i = threadIdx.x; i < np1; i += blockDim.x) wc[i] = A0[i];
      __syncthreads();
    }
  }
}
void launch_opgradt(const DevMesh& dm, const double* p, Ptr3 w, cudaStream_t st) {
  if (tp_opgradt(dm, p, w, st)) return;
  size_t smem = (size_t)(2 * dm.q * dm.n + 7 * dm.np1) * sizeof(double);
  static size_t set = 0;
  if (smem > 48 * 1024 && smem > set) { cudaFuncSetAttribute(k_opgradt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = smem; }
  k_opgradt<<<(unsigned)dm.E, elem_threads(dm.np1), smem, st>>>(p, w, dm.rxw2, dm.I12t, dm.D12t, dm.n, dm.q, dm.ndim);
  LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K3/K4 convection
// out_f (+)= alpha * I^T [ sum_k (sum_c rxd[k][c] * I C_c) * D_k (I u_f) ]   -- Nek convect_new, ifcf=.false.
__global__ void k_convect(CPtr4 u, int nf, CPtr3 C, Ptr4 out, const double* __restrict__ rxd, const double* __restrict__ I1dg,
                          const double* __restrict__ I1dtg, const double* __restrict__ Ddg, int n, int m, int d, double alpha,
                          int accumulate) {
  extern __shared__ double sm[];
  const int nz = d == 3 ? n : 1, mz = d == 3 ? m : 1;
  const int np1 = n * n * nz, npd = m * m * mz;
  double* sI = sm; double* sIt = sI + m * n; double* sDd = sIt + m * n;
  double* TR = sDd + m * m;             // d * npd
  double* UF = TR + 3 * npd;            // npd
  double* W1 = UF + npd;                // npd (scratch)
  double* W2 = W1 + npd;                // npd (scratch)
  double* ACC = W2 + npd;               // npd
  const size_t e = blockIdx.x;
  load_mat(sI, I1dg, m * n); load_mat(sIt, I1dtg, m * n); load_mat(sDd, Ddg, m * m);
  __syncthreads();
  // interpolate the convecting field, then tr_k = sum_c rxd[k][c] * Cf_c (in place, pointwise)
  for (int c = 0; c < d; ++c) {
    const double* cc = C.p[c] + e * np1;
    for (int i = threadIdx.x; i < np1; i += blockDim.x) W1[i] = cc[i];
    __syncthreads();
    if (d == 2) {
      contract<false>(W2, W1, sI, m, n, 0, n, n, 1);
      contract<false>(TR + c * npd, W2, sI, m, n, 1, m, n, 1);
    } else {
      contract<false>(W2, W1, sI, m, n, 0, n, n, n);
      contract<false>(W1, W2, sI, m, n, 1, m, n, n);
      contract<false>(TR + c * npd, W1, sI, m, n, 2, m, m, n);
    }
  }
  const double* rx = rxd + e * (size_t)(d * d) * npd;
  for (int i = threadIdx.x; i < npd; i += blockDim.x) {
    double cf[3] = {TR[i], TR[npd + i], d == 3 ? TR[2 * npd + i] : 0.0};
    for (int k = 0; k < d; ++k) {
      double s = 0;
      for (int c = 0; c < d; ++c) s += rx[(k * d + c) * npd + i] * cf[c];
      TR[k * npd + i] = s;
    }
  }
  __syncthreads();
  for (int f = 0; f < nf; ++f) {
    const double* uf = u.p[f] + e * np1;
    for (int i = threadIdx.x; i < np1; i += blockDim.x) W1[i] = uf[i];
    __syncthreads();
    if (d == 2) {
      contract<false>(W2, W1, sI, m, n, 0, n, n, 1);
      contract<false>(UF, W2, sI, m, n, 1, m, n, 1);
      contract<false>(W1, UF, sDd, m, m, 0, m, m, 1);
      contract<false>(W2, UF, sDd, m, m, 1, m, m, 1);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) ACC[i] = TR[i] * W1[i] + TR[npd + i] * W2[i];
      __syncthreads();
      contract<false>(W1, ACC, sIt, n, m, 0, m, m, 1);
      contract<false>(W2, W1, sIt, n, m, 1, n, m, 1);
    } else {
      contract<false>(W2, W1, sI, m, n, 0, n, n, n);
      contract<false>(W1, W2, sI, m, n, 1, m, n, n);
      contract<false>(UF, W1, sI, m, n, 2, m, m, n);
      contract<false>(W1, UF, sDd, m, m, 0, m, m, m);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) ACC[i] = TR[i] * W1[i];
      __syncthreads();
      contract<false>(W1, UF, sDd, m, m, 1, m, m, m);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) ACC[i] += TR[npd + i] * W1[i];
      __syncthreads();
      contract<false>(W1, UF, sDd, m, m, 2, m, m, m);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) ACC[i] += TR[2 * npd + i] * W1[i];
      __syncthreads();
      contract<false>(W1, ACC, sIt, n, m, 0, m, m, m);
      contract<false>(UF, W1, sIt, n, m, 1, n, m, m);
      contract<false>(W2, UF, sIt, n, m, 2, n, n, m);
    }
    double* of = out.p[f] + e * np1;
    for (int i = threadIdx.x; i < np1; i += blockDim.x) of[i] = (accumulate ? of[i] : 0.0) + alpha * W2[i];
    __syncthreads();
  }
}
void launch_convect(const DevMesh& dm, CPtr4 u, int nf, CPtr3 C, Ptr4 out, double alpha, int accumulate, cudaStream_t st) {
  if (tp_convect(dm, u, nf, C, out, alpha, accumulate, st)) return;
  size_t smem = (size_t)(2 * dm.m * dm.n + dm.m * dm.m + 7 * dm.npd) * sizeof(double);
  static size_t set = 0;
  if (smem > 48 * 1024 && smem > set) { cudaFuncSetAttribute(k_convect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = smem; }
  k_convect<<<(unsigned)dm.E, elem_threads(dm.npd), smem, st>>>(u, nf, C, out, dm.rxd, dm.I1d, dm.I1dt, dm.Dd, dm.n, dm.m, dm.ndim, alpha, accumulate);
  LAUNCH_COUNT();
}

// out_i (+)= alpha * I^T [ sum_j (I c_j) * sum_k rxd[k][i] D_k (I U_j) ]   -- Nek convect_adj
__global__ void k_convect_adj(CPtr3 U, CPtr3 cf, Ptr3 out, const double* __restrict__ rxd, const double* __restrict__ I1dg,
                              const double* __restrict__ I1dtg, const double* __restrict__ Ddg, int n, int m, int d, double alpha,
                              int accumulate, int nj) {
  extern __shared__ double sm[];
  const int nz = d == 3 ? n : 1, mz = d == 3 ? m : 1;
  const int np1 = n * n * nz, npd = m * m * mz;
  double* sI = sm; double* sIt = sI + m * n; double* sDd = sIt + m * n;
  double* AC = sDd + m * m;             // d * npd accumulators
  double* UF = AC + 3 * npd; double* W1 = UF + npd; double* W2 = W1 + npd; double* CF = W2 + npd;
  const size_t e = blockIdx.x;
  load_mat(sI, I1dg, m * n); load_mat(sIt, I1dtg, m * n); load_mat(sDd, Ddg, m * m);
  for (int i = threadIdx.x; i < d * npd; i += blockDim.x) AC[i] = 0.0;
  __syncthreads();
  const double* rx = rxd + e * (size_t)(d * d) * npd;
  for (int j = 0; j < nj; ++j) {
    // CF = I c_j ; UF = I U_j
    for (int pass = 0; pass < 2; ++pass) {
      const double* src = (pass == 0 ? cf.p[j] : U.p[j]) + e * np1;
      double* dst = pass == 0 ? CF : UF;
      for (int i = threadIdx.x; i < np1; i += blockDim.x) W1[i] = src[i];
      __syncthreads();
      if (d == 2) { contract<false>(W2, W1, sI, m, n, 0, n, n, 1); contract<false>(dst, W2, sI, m, n, 1, m, n, 1); }
      else { contract<false>(W2, W1, sI, m, n, 0, n, n, n); contract<false>(W1, W2, sI, m, n, 1, m, n, n); contract<false>(dst, W1, sI, m, n, 2, m, m, n); }
    }
    for (int k = 0; k < d; ++k) {
      contract<false>(W1, UF, sDd, m, m, k, m, m, mz);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) {
        double g = CF[i] * W1[i];
        for (int c = 0; c < d; ++c) AC[c * npd + i] += rx[(k * d + c) * npd + i] * g;
      }
      __syncthreads();
    }
  }
  for (int c = 0; c < d; ++c) {
    if (d == 2) { contract<false>(W1, AC + c * npd, sIt, n, m, 0, m, m, 1); contract<false>(W2, W1, sIt, n, m, 1, n, m, 1); }
    else { contract<false>(W1, AC + c * npd, sIt, n, m, 0, m, m, m); contract<false>(UF, W1, sIt, n, m, 1, n, m, m); contract<false>(W2, UF, sIt, n, m, 2, n, n, m); }
    double* of = out.p[c] + e * np1;
    for (int i = threadIdx.x; i < np1; i += blockDim.x) of[i] = (accumulate ? of[i] : 0.0) + alpha * W2[i];
    __syncthreads();
  }
}
void launch_convect_adj(const DevMesh& dm, CPtr3 U, CPtr3 c, Ptr3 out, double alpha, int accumulate, cudaStream_t st, int nj) {
  if (nj < 0) nj = dm.ndim;
  if (tp_convect_adj(dm, U, c, out, alpha, accumulate, st, nj)) return;
  size_t smem = (size_t)(2 * dm.m * dm.n + dm.m * dm.m + 7 * dm.npd) * sizeof(double);
  static size_t set = 0;
  if (smem > 48 * 1024 && smem > set) { cudaFuncSetAttribute(k_convect_adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = smem; }
  k_convect_adj<<<(unsigned)dm.E, elem_threads(dm.npd), smem, st>>>(U, c, out, dm.rxd, dm.I1d, dm.I1dt, dm.Dd, dm.n, dm.m, dm.ndim, alpha, accumulate, nj);
  LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K7 rhs tail
// makextp + makebdfp + lagfieldp fused, all components in one launch (blockIdx.y = component).
__global__ void k_rhs_tail(RhsTail t, size_t n, const double* __restrict__ bm1, double ab0, double ab1, double ab2, double bd1, double bd2,
                           double bd3) {
  const int f = blockIdx.y;
  double* bf = t.bf[f]; double* e1 = t.e1[f]; double* e2 = t.e2[f]; const double* u = t.u[f]; double* l1 = t.lag1[f]; double* l2 = t.lag2[f];
  const double coef = t.coef[f];
  NLK_STREAM4_BEGIN(i, n)
    double b = bf[i], x1 = e1[i], x2 = e2[i], uu = u[i], a1 = l1[i], a2 = l2[i];
    e2[i] = x1; e1[i] = b;
    bf[i] = ab0 * b + ab1 * x1 + ab2 * x2 + coef * bm1[i] * (bd1 * uu + bd2 * a1 + bd3 * a2);
    l2[i] = a1; l1[i] = uu;
    NLK_STREAM4_END
}
void launch_rhs_tail(const DevMesh& dm, const RhsTail& t, int nf, double ab0, double ab1, double ab2, double bd1, double bd2, double bd3,
                     cudaStream_t st) {
  dim3 grid(std::min(cdiv(dm.N1, 256), 148 * 8), nf);
  k_rhs_tail<<<grid, 256, 0, st>>>(t, dm.N1, dm.bm1, ab0, ab1, ab2, bd1, bd2, bd3); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K8 CG vector phases
// Nek cggo scalar logic: convergence test on rbn2 = sqrt(sum r^2 mult binv / vol), beta = rtz1/rtz2.
__device__ __forceinline__ void cg_finalize_zr(SolverScal* sc, double vol) {
  double v0 = sc->red[0], v1 = sc->red[1];
  sc->rtz2 = sc->rtz1; sc->rtz1 = v0;
  double rbn2 = sqrt(v1 / vol);
  sc->rbn2 = rbn2;
  if (sc->iter == 0) sc->rbn0 = rbn2;
  if (rbn2 <= sc->tol || sc->iter >= sc->maxit || !(v0 == v0)) { sc->done = 1; }
  else { sc->beta = sc->iter == 0 ? 0.0 : v0 / sc->rtz2; sc->iter += 1; }
}
__device__ __forceinline__ void cg_finalize_pap(SolverScal* sc) { sc->pap = sc->red[2]; sc->alpha = sc->rtz1 / sc->red[2]; }
__global__ void k_cg_finalize(SolverScal* sc, int which, double vol) {
  if (sc->done) return;
  if (which == 0) cg_finalize_zr(sc, vol); else cg_finalize_pap(sc);
}
void launch_cg_finalize(SolverScal* sc, int which, double vol, cudaStream_t st) { k_cg_finalize<<<1, 1, 0, st>>>(sc, which, vol); LAUNCH_COUNT(); }
__global__ void k_recip(double* __restrict__ out, const double* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = 1.0 / x[i];
}
void launch_recip(double* out, const double* x, size_t n, cudaStream_t st) {
  k_recip<<<std::min(cdiv(n, 256), 148 * 16), 256, 0, st>>>(out, x, n); LAUNCH_COUNT();
}
__global__ void k_cg_init(SolverScal* sc, double tol, int maxit) {
  sc->rtz1 = 1.0; sc->rtz2 = 1.0; sc->rbn2 = 0; sc->rbn0 = 0; sc->pap = 0; sc->alpha = 0; sc->beta = 0; sc->tol = tol;
  sc->iter = 0; sc->done = 0; sc->maxit = maxit;
}
void launch_cg_init(const DevMesh&, SolverScal* sc, double tol, int maxit, cudaStream_t st) { k_cg_init<<<1, 1, 0, st>>>(sc, tol, maxit); LAUNCH_COUNT(); }

// Per-(h1,h2) weights of the streamed PCG, so that its two per-iteration kernels read three arrays instead of seven:
//   hd = mask / (h1 diagA + h2 diagB)   (masked inverse Jacobi diagonal: z = hd r, so p and x stay 0 on Dirichlet nodes and
//                                        r needs no masking there -- whatever accumulates in r on those nodes is never used)
//   wa = hd * mult                      (r.z   = sum r^2 wa)
//   wb = mask * mult * binv             (rbn2^2 = sum r^2 wb / vol, Nek cggo's convergence norm)
__global__ void k_cg_weights(size_t n, const double* __restrict__ mask, const double* __restrict__ diagA, const double* __restrict__ diagB,
                             const double* __restrict__ mult, const double* __restrict__ binv, double h1, double h2,
                             double* __restrict__ hd, double* __restrict__ wa, double* __restrict__ wb) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double m = mask[i], d = m / (h1 * diagA[i] + h2 * diagB[i]), mu = mult[i];
    hd[i] = d; wa[i] = d * mu; wb[i] = m * mu * binv[i];
  }
}
void launch_cg_weights(const DevMesh& dm, const double* mask, double h1, double h2, double* hd, double* wa, double* wb, cudaStream_t st) {
  k_cg_weights<<<std::min(cdiv(dm.N1, 256), 148 * 16), 256, 0, st>>>(dm.N1, mask, dm.diagA, dm.diagB, dm.vmult, dm.binvm1, h1, h2, hd, wa, wb); LAUNCH_COUNT();
}
// x += alpha p ; r -= alpha w   [skipped when first]; then rtz1 = sum r^2 wa, rbn2 = sum r^2 wb.
// Finaliser applies Nek cggo's convergence test and sets beta.
__global__ void __launch_bounds__(256)
k_cg_update_reduce(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ w,
                   const double* __restrict__ wa, const double* __restrict__ wb, double vol, size_t n,
                   SolverScal* sc, Reducer red, int first, int defer) {
  if (sc->done) return;
  const double alpha = first ? 0.0 : sc->alpha;
  double v[2] = {0.0, 0.0};
  NLK_STREAM4_BEGIN(i, n)
    double ri = r[i];
    if (!first) {
      x[i] += alpha * p[i];
      ri -= alpha * w[i];
      r[i] = ri;
    }
    const double r2 = ri * ri;
    v[0] += r2 * wa[i];
    v[1] += r2 * wb[i];
    NLK_STREAM4_END
  if (grid_reduce<2>(v, red)) {
    sc->red[0] = v[0]; sc->red[1] = v[1];
    if (!defer) cg_finalize_zr(sc, vol);
  }
}
void launch_cg_update_reduce(const DevMesh& dm, double* x, double* r, const double* p, const double* w, const double* wa, const double* wb,
                             SolverScal* sc, Reducer red, int first, int defer, cudaStream_t st) {
  int grid = stream_grid(dm.N1, RED_BLOCKS);
  k_cg_update_reduce<<<grid, RED_THREADS, 0, st>>>(x, r, p, w, wa, wb, dm.volvm1, dm.N1, sc, red, first, defer);
  LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ dots
template <int NP, bool W>
__global__ void __launch_bounds__(256)
k_dot(size_t n, const double* __restrict__ a0, const double* __restrict__ a1, const double* __restrict__ a2, const double* __restrict__ a3,
      const double* __restrict__ b0, const double* __restrict__ b1, const double* __restrict__ b2, const double* __restrict__ b3,
      const double* __restrict__ c, double* out, Reducer red) {
  double v[1] = {0.0};
  NLK_STREAM4_BEGIN(i, n)
    double s = a0[i] * b0[i];
    if (NP > 1) s += a1[i] * b1[i];
    if (NP > 2) s += a2[i] * b2[i];
    if (NP > 3) s += a3[i] * b3[i];
    v[0] += W ? s * c[i] : s;
    NLK_STREAM4_END
  if (grid_reduce<1>(v, red)) out[0] = v[0];
}
void launch_dot(size_t n, CPtr4 a, CPtr4 b, int npairs, const double* c, double* out, Reducer red, cudaStream_t st) {
  int grid = stream_grid(n, RED_BLOCKS);
#define NLK_DOT(NP_) { if (c) k_dot<NP_, true><<<grid, RED_THREADS, 0, st>>>(n, a.p[0], a.p[1], a.p[2], a.p[3], b.p[0], b.p[1], b.p[2], b.p[3], c, out, red); \
                       else k_dot<NP_, false><<<grid, RED_THREADS, 0, st>>>(n, a.p[0], a.p[1], a.p[2], a.p[3], b.p[0], b.p[1], b.p[2], b.p[3], c, out, red); }
  switch (npairs) { case 1: NLK_DOT(1) break; case 2: NLK_DOT(2) break; case 3: NLK_DOT(3) break; default: NLK_DOT(4) break; }
#undef NLK_DOT
  LAUNCH_COUNT();
}

// h[j] = sum_i V[j*ld + i] * w[i], j < k <= 32: w is streamed once for all k rows (tall-skinny V^T w).
template <int KB>
__global__ void __launch_bounds__(256)
k_multidot(const double* __restrict__ V, size_t ld, int k0, int k, const double* __restrict__ w, size_t n, double* h, Reducer red) {
  double v[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) v[j] = 0.0;
  NLK_STREAM4_BEGIN(i, n)
    double wi = w[i];
#pragma unroll
    for (int j = 0; j < KB; ++j) if (k0 + j < k) v[j] += V[(size_t)(k0 + j) * ld + i] * wi;
    NLK_STREAM4_END
  if (grid_reduce<KB>(v, red)) {
    for (int j = 0; j < KB; ++j) if (k0 + j < k) h[k0 + j] = v[j];
  }
}
void launch_multidot(const double* V, size_t ld, int k, const double* w, size_t n, double* h, Reducer red, cudaStream_t st) {
  int grid = stream_grid(n, RED_BLOCKS);
  for (int k0 = 0; k0 < k; k0 += 8) { k_multidot<8><<<grid, RED_THREADS, 0, st>>>(V, ld, k0, k, w, n, h, red); LAUNCH_COUNT(); }
}
__global__ void k_multiaxpy(double* __restrict__ w, const double* __restrict__ V, size_t ld, int k, const double* __restrict__ h, double sign, size_t n) {
  extern __shared__ double sh[];
  for (int j = threadIdx.x; j < k; j += blockDim.x) sh[j] = sign * h[j];
  __syncthreads();
  NLK_STREAM4_BEGIN(i, n)
    double s = w[i];
    for (int j = 0; j < k; ++j) s += sh[j] * V[(size_t)j * ld + i];
    w[i] = s;
    NLK_STREAM4_END
}
// w += sign * V h, and out[0] = sum_i w_i^2 of the result (one pass)
__global__ void __launch_bounds__(256)
k_multiaxpy_norm(double* __restrict__ w, const double* __restrict__ V, size_t ld, int k, const double* __restrict__ h, double sign, size_t n,
                 double* out, Reducer red) {
  extern __shared__ double sh[];
  for (int j = threadIdx.x; j < k; j += blockDim.x) sh[j] = sign * h[j];
  __syncthreads();
  double v[1] = {0.0};
  NLK_STREAM4_BEGIN(i, n)
    double s = w[i];
    for (int j = 0; j < k; ++j) s += sh[j] * V[(size_t)j * ld + i];
    w[i] = s; v[0] += s * s;
    NLK_STREAM4_END
  if (grid_reduce<1>(v, red)) out[0] = v[0];
}
void launch_multiaxpy_norm(double* w, const double* V, size_t ld, int k, const double* h, double sign, size_t n, double* out, Reducer red, cudaStream_t st) {
  int grid = stream_grid(n, RED_BLOCKS);
  k_multiaxpy_norm<<<grid, RED_THREADS, (k > 0 ? k : 1) * sizeof(double), st>>>(w, V, ld, k, h, sign, n, out, red); LAUNCH_COUNT();
}
// out = in * rsqrt(*s)  (zero if *s <= 0): normalisation with a device-resident scalar
__global__ void k_scale_rsqrt(double* __restrict__ out, const double* __restrict__ in, size_t n, const double* __restrict__ s) {
  const double sv = *s; const double f = sv > 0 ? rsqrt(sv) : 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i] * f;
}
void launch_scale_rsqrt(double* out, const double* in, size_t n, const double* s, cudaStream_t st) {
  k_scale_rsqrt<<<std::min(cdiv(n, 256), 148 * 8), 256, 0, st>>>(out, in, n, s); LAUNCH_COUNT();
}
void launch_multiaxpy(double* w, const double* V, size_t ld, int k, const double* h, double sign, size_t n, cudaStream_t st) {
  if (k <= 0) return;
  k_multiaxpy<<<stream_grid(n, 148 * 8), 256, k * sizeof(double), st>>>(w, V, ld, k, h, sign, n); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K10 Schwarz
__device__ __forceinline__ int clamp_inner(int i, int n) { return i == 0 ? 1 : (i == n - 1 ? n - 2 : i); }

// w (n^d) <- r (q^d): interior copy; face layers (tangentially interior) = first interior layer; edges/corners = 0
__global__ void k_schwarz_embed(const double* __restrict__ r, const double* __restrict__ mul, double* __restrict__ w, int n, int d, size_t N1) {
  const int q = n - 2, np1 = d == 3 ? n * n * n : n * n, np2 = d == 3 ? q * q * q : q * q;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < N1; g += (size_t)gridDim.x * blockDim.x) {
    size_t e = g / np1; int p = (int)(g - e * np1);
    int i = p % n, j = (p / n) % n, k = d == 3 ? p / (n * n) : 1;
    int nb = (i == 0 || i == n - 1) + (j == 0 || j == n - 1) + (d == 3 ? (k == 0 || k == n - 1) : 0);
    double v = 0.0;
    if (nb <= 1) {
      int ii = clamp_inner(i, n) - 1, jj = clamp_inner(j, n) - 1, kk = d == 3 ? clamp_inner(k, n) - 1 : 0;
      size_t s = e * np2 + ((size_t)kk * q + jj) * q + ii;
      v = mul ? r[s] * mul[s] : r[s];
    }
    w[g] = v;
  }
}
void launch_schwarz_embed(const DevMesh& dm, const double* r, const double* mul, double* w, cudaStream_t st) {
  k_schwarz_embed<<<std::min(cdiv(dm.N1, 256), 148 * 8), 256, 0, st>>>(r, mul, w, dm.n, dm.ndim, dm.N1); LAUNCH_COUNT();
}

// per element: face fix (w_face -= w_inner), z = (S (x) S (x) S) dinv (S^T (x) S^T (x) S^T) w ; t = z on faces else 0.
// If S == nullptr the tensor solve is skipped (used to build the overlap-count weights).
__global__ void k_schwarz_fdm(const double* __restrict__ w, double* __restrict__ z, double* __restrict__ t, const double* __restrict__ S,
                              const double* __restrict__ St, const double* __restrict__ dinv, int n, int d) {
  extern __shared__ double sm[];
  const int nz = d == 3 ? n : 1, np1 = n * n * nz, nn = n * n;
  double* sS = sm; double* sSt = sS + d * nn; double* A = sSt + d * nn; double* B = A + np1;
  const size_t e = blockIdx.x;
  if (S) { load_mat(sS, S + e * (size_t)d * nn, d * nn); load_mat(sSt, St + e * (size_t)d * nn, d * nn); }
  for (int p = threadIdx.x; p < np1; p += blockDim.x) B[p] = w[e * np1 + p];
  __syncthreads();
  for (int p = threadIdx.x; p < np1; p += blockDim.x) {
    int i = p % n, j = (p / n) % n, k = d == 3 ? p / nn : 1;
    int nb = (i == 0 || i == n - 1) + (j == 0 || j == n - 1) + (d == 3 ? (k == 0 || k == n - 1) : 0);
    double v = B[p];
    if (nb == 1) {
      int ii = clamp_inner(i, n), jj = clamp_inner(j, n), kk = d == 3 ? clamp_inner(k, n) : 0;
      v -= B[(kk * n + jj) * n + ii];
    }
    A[p] = v;
  }
  __syncthreads();
  double* res = A;
  if (S) {
    contract<false>(B, A, sSt, n, n, 0, n, n, nz);
    contract<false>(A, B, sSt + nn, n, n, 1, n, n, nz);
    if (d == 3) { contract<false>(B, A, sSt + 2 * nn, n, n, 2, n, n, nz); }
    double* cur = d == 3 ? B : A; double* oth = d == 3 ? A : B;
    for (int p = threadIdx.x; p < np1; p += blockDim.x) cur[p] *= dinv[e * np1 + p];
    __syncthreads();
    contract<false>(oth, cur, sS, n, n, 0, n, n, nz);
    contract<false>(cur, oth, sS + nn, n, n, 1, n, n, nz);
    res = cur;
    if (d == 3) { contract<false>(oth, cur, sS + 2 * nn, n, n, 2, n, n, nz); res = oth; }
  }
  for (int p = threadIdx.x; p < np1; p += blockDim.x) {
    int i = p % n, j = (p / n) % n, k = d == 3 ? p / nn : 1;
    int nb = (i == 0 || i == n - 1) + (j == 0 || j == n - 1) + (d == 3 ? (k == 0 || k == n - 1) : 0);
    double v = res[p];
    z[e * np1 + p] = v;
    t[e * np1 + p] = nb == 1 ? v : 0.0;
  }
}
void launch_schwarz_fdm(const DevMesh& dm, const double* w, double* z, double* t, cudaStream_t st) {
  if (tp_schwarz_fdm(dm, w, z, t, st)) return;
  size_t smem = (size_t)(2 * dm.ndim * dm.n * dm.n + 2 * dm.np1) * sizeof(double);
  k_schwarz_fdm<<<(unsigned)dm.E, elem_threads(dm.np1), smem, st>>>(w, z, t, dm.fdmS, dm.fdmSt, dm.fdmDinv, dm.n, dm.ndim); LAUNCH_COUNT();
}
void launch_schwarz_count(const DevMesh& dm, const double* w, double* z, double* t, cudaStream_t st) {
  size_t smem = (size_t)(2 * dm.ndim * dm.n * dm.n + 2 * dm.np1) * sizeof(double);
  k_schwarz_fdm<<<(unsigned)dm.E, elem_threads(dm.np1), smem, st>>>(w, z, t, nullptr, nullptr, nullptr, dm.n, dm.ndim); LAUNCH_COUNT();
}

// out (q^d) = wt * ( z_interior + [first interior layers] (tsum_face - z_face) )
__global__ void k_schwarz_gather(const double* __restrict__ z, const double* __restrict__ ts, double* __restrict__ out,
                                 const double* __restrict__ wt, int n, int d, size_t N2) {
  const int q = n - 2, np1 = d == 3 ? n * n * n : n * n, np2 = d == 3 ? q * q * q : q * q, nn = n * n;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < N2; g += (size_t)gridDim.x * blockDim.x) {
    size_t e = g / np2; int p = (int)(g - e * np2);
    int i = p % q + 1, j = (p / q) % q + 1, k = d == 3 ? p / (q * q) + 1 : 0;
    const double* ze = z + e * np1; const double* te = ts + e * np1;
    double v = ze[(k * n + j) * n + i];
    if (i == 1) { int f = (k * n + j) * n + 0; v += te[f] - ze[f]; }
    if (i == n - 2) { int f = (k * n + j) * n + n - 1; v += te[f] - ze[f]; }
    if (j == 1) { int f = (k * n + 0) * n + i; v += te[f] - ze[f]; }
    if (j == n - 2) { int f = (k * n + n - 1) * n + i; v += te[f] - ze[f]; }
    if (d == 3) {
      if (k == 1) { int f = (0 * n + j) * n + i; v += te[f] - ze[f]; }
      if (k == n - 2) { int f = ((n - 1) * n + j) * n + i; v += te[f] - ze[f]; }
    }
    (void)nn;
    out[g] = wt ? v * wt[g] : v;
  }
}
void launch_schwarz_gather(const DevMesh& dm, const double* z, const double* t, double* out, cudaStream_t st) {
  k_schwarz_gather<<<std::min(cdiv(dm.N2, 256), 148 * 8), 256, 0, st>>>(z, t, out, dm.swt, dm.n, dm.ndim, dm.N2); LAUNCH_COUNT();
}
void launch_schwarz_gather_nowt(const DevMesh& dm, const double* z, const double* t, double* out, cudaStream_t st) {
  k_schwarz_gather<<<std::min(cdiv(dm.N2, 256), 148 * 8), 256, 0, st>>>(z, t, out, nullptr, dm.n, dm.ndim, dm.N2); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K11 coarse grid
__device__ __forceinline__ double corner_shape(int c, int p, int q, int d, const double* __restrict__ z2) {
  int i = p % q, j = (p / q) % q, k = d == 3 ? p / (q * q) : 0;
  double hx = (c & 1) ? 0.5 * (1 + z2[i]) : 0.5 * (1 - z2[i]);
  double hy = ((c >> 1) & 1) ? 0.5 * (1 + z2[j]) : 0.5 * (1 - z2[j]);
  double hz = d == 3 ? (((c >> 2) & 1) ? 0.5 * (1 + z2[k]) : 0.5 * (1 - z2[k])) : 1.0;
  return hx * hy * hz;
}
__global__ void k_coarse_part(const double* __restrict__ r, const double* __restrict__ mul, double* __restrict__ part, const double* __restrict__ z2, int q, int d, size_t nec) {
  const int nv = 1 << d, np2 = d == 3 ? q * q * q : q * q;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nec) return;
  size_t e = t / nv; int c = (int)(t % nv);
  double s = 0;
  if (mul) for (int p = 0; p < np2; ++p) s += corner_shape(c, p, q, d, z2) * r[e * np2 + p] * mul[e * np2 + p];
  else for (int p = 0; p < np2; ++p) s += corner_shape(c, p, q, d, z2) * r[e * np2 + p];
  part[t] = s;
}
__global__ void k_vert_gather(const double* __restrict__ part, const int32_t* __restrict__ off, const int32_t* __restrict__ ec, double* __restrict__ rc, int64_t nvert) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvert) return;
  double s = 0;
  for (int t = off[v]; t < off[v + 1]; ++t) s += part[ec[t]];
  rc[v] = s;
}
// one warp per plane group: bm1-weighted averages of 2 u cv and 2 u sv (Nek planar_avg), fixed summation order (deterministic)
__global__ void k_planar_avg(const double* __restrict__ u, const double* __restrict__ bm1, const double* __restrict__ cv, const double* __restrict__ sv,
                             const int32_t* __restrict__ off, const int32_t* __restrict__ idx, int64_t ngroups, double* __restrict__ coef) {
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  double sc = 0, ss = 0, sm_ = 0;
  for (int t = off[g] + lane; t < off[g + 1]; t += 32) { const int p = idx[t]; const double b = bm1[p], ub = 2.0 * u[p] * b; sc += ub * cv[p]; ss += ub * sv[p]; sm_ += b; }
  sc = warp_sum(sc); ss = warp_sum(ss); sm_ = warp_sum(sm_);
  if (lane == 0) { coef[2 * g] = sc / sm_; coef[2 * g + 1] = ss / sm_; }
}
__global__ void k_planar_apply(double* __restrict__ u, const double* __restrict__ cv, const double* __restrict__ sv, const int32_t* __restrict__ gid,
                               const double* __restrict__ coef, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const int g = gid[i]; u[i] = cv[i] * coef[2 * g] + sv[i] * coef[2 * g + 1]; }
}
void launch_planar_proj(double* u, const double* bm1, const double* cv, const double* sv, const int32_t* off, const int32_t* idx, const int32_t* gid,
                        int64_t ngroups, size_t N1, double* coef, cudaStream_t st) {
  k_planar_avg<<<cdiv((size_t)ngroups, 4), 128, 0, st>>>(u, bm1, cv, sv, off, idx, ngroups, coef); LAUNCH_COUNT();
  k_planar_apply<<<std::min(cdiv(N1, 256), 148 * 8), 256, 0, st>>>(u, cv, sv, gid, coef, N1); LAUNCH_COUNT();
}
__global__ void k_publish(const double* __restrict__ src, int count, double* __restrict__ h_dst, volatile unsigned int* h_seq, unsigned int seq) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) h_dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) { *h_seq = seq; __threadfence_system(); }
}
void launch_publish(const double* d_src, int count, double* h_dst, unsigned int* h_seq, unsigned int seq, cudaStream_t st) {
  k_publish<<<1, 64, 0, st>>>(d_src, count, h_dst, h_seq, seq); LAUNCH_COUNT();
}
void launch_vert_gather(const DevMesh& dm, const double* part, double* rc, cudaStream_t st) {
  k_vert_gather<<<cdiv(dm.nvert, 128), 128, 0, st>>>(part, dm.vert_off, dm.vert_ec, rc, dm.nvert); LAUNCH_COUNT();
}
void launch_coarse_restrict(const DevMesh& dm, const double* r, const double* mul, double* part, double* rc, cudaStream_t st) {
  size_t nec = (size_t)dm.E << dm.ndim;
  k_coarse_part<<<cdiv(nec, 128), 128, 0, st>>>(r, mul, part, dm.w2 + dm.q, dm.q, dm.ndim, nec); LAUNCH_COUNT();
  k_vert_gather<<<cdiv(dm.nvert, 128), 128, 0, st>>>(part, dm.vert_off, dm.vert_ec, rc, dm.nvert); LAUNCH_COUNT();
}
__global__ void k_gemv(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y, int n) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* a = A + (size_t)row * n;
  double s = 0;
  for (int j = lane; j < n; j += 32) s += a[j] * x[j];
  s = warp_sum(s);
  if (lane == 0) y[row] = s;
}
void launch_gemv(const double* A, const double* x, double* y, int n, cudaStream_t st) {
  k_gemv<<<cdiv(n, 8), 256, 0, st>>>(A, x, y, n); LAUNCH_COUNT();
}
__global__ void k_coarse_prolong(const double* __restrict__ c, double* __restrict__ z, const int64_t* __restrict__ vertex,
                                 const double* __restrict__ z2, int q, int d, size_t N2, int accumulate) {
  const int nv = 1 << d, np2 = d == 3 ? q * q * q : q * q;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < N2; g += (size_t)gridDim.x * blockDim.x) {
    size_t e = g / np2; int p = (int)(g - e * np2);
    double s = 0;
    for (int k = 0; k < nv; ++k) s += corner_shape(k, p, q, d, z2) * c[vertex[e * nv + k] - 1];
    z[g] = (accumulate ? z[g] : 0.0) + s;
  }
}
void launch_coarse_prolong_add(const DevMesh& dm, const double* c, double* z, int accumulate, cudaStream_t st) {
  k_coarse_prolong<<<std::min(cdiv(dm.N2, 256), 148 * 8), 256, 0, st>>>(c, z, dm.vertex, dm.w2 + dm.q, dm.q, dm.ndim, dm.N2, accumulate); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K14 CFL
__global__ void __launch_bounds__(256)
k_cfl(CPtr3 u, const double* __restrict__ rxj, const double* __restrict__ dri, int n, int d, size_t N1, double* out, Reducer red) {
  const int np1 = d == 3 ? n * n * n : n * n;
  double v[1] = {0.0};
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < N1; g += (size_t)gridDim.x * blockDim.x) {
    size_t e = g / np1; int p = (int)(g - e * np1);
    int idx[3] = {p % n, (p / n) % n, d == 3 ? p / (n * n) : 0};
    double s = 0;
    for (int k = 0; k < d; ++k) {
      double uk = 0;
      for (int c = 0; c < d; ++c) uk += u.p[c][g] * rxj[(e * d * d + (size_t)(k * d + c)) * np1 + p];
      s += fabs(uk) * dri[idx[k]];
    }
    v[0] = fmax(v[0], s);
  }
  if (grid_reduce<1, true>(v, red)) out[0] = v[0];
}
void launch_cfl(const DevMesh& dm, CPtr3 u, double* out, Reducer red, cudaStream_t st) {
  int grid = std::min(cdiv(dm.N1, RED_THREADS), RED_BLOCKS);
  k_cfl<<<grid, RED_THREADS, 0, st>>>(u, dm.rxj, dm.dri, dm.n, dm.ndim, dm.N1, out, red); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ K15 filter
__global__ void k_filter(const double* __restrict__ Fg, Ptr4 u, int n, int d) {
  extern __shared__ double sm[];
  const int nz = d == 3 ? n : 1, np1 = n * n * nz;
  double* sF = sm; double* A = sF + n * n; double* B = A + np1;
  const size_t e = blockIdx.x;
  double* ue = u.p[blockIdx.y] + e * np1;
  load_mat(sF, Fg, n * n);
  for (int p = threadIdx.x; p < np1; p += blockDim.x) A[p] = ue[p];
  __syncthreads();
  contract<false>(B, A, sF, n, n, 0, n, n, nz);
  contract<false>(A, B, sF, n, n, 1, n, n, nz);
  double* res = A;
  if (d == 3) { contract<false>(B, A, sF, n, n, 2, n, n, nz); res = B; }
  for (int p = threadIdx.x; p < np1; p += blockDim.x) ue[p] = res[p];
}
void launch_filter(const DevMesh& dm, const double* F, Ptr4 u, int nf, cudaStream_t st) {
  size_t smem = (size_t)(dm.n * dm.n + 2 * dm.np1) * sizeof(double);
  dim3 grid((unsigned)dm.E, nf);
  k_filter<<<grid, elem_threads(dm.np1), smem, st>>>(F, u, dm.n, dm.ndim); LAUNCH_COUNT();
}

// ------------------------------------------------------------------------------------------------ seeded field
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void k_rand_field(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                             const int64_t* __restrict__ lglel, int np1, size_t N1, uint64_t seed, int comp) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < N1; g += (size_t)gridDim.x * blockDim.x) {
    size_t e = g / np1; int p = (int)(g - e * np1);
    uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)lglel[e] * 4096ull + (uint64_t)p + ((uint64_t)comp << 40)));
    double noise = ((double)(h >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
    double f = sin(0.7 * x[g] + 0.3 * comp) * cos(0.9 * y[g] - 0.2 * comp);
    if (z) f *= cos(0.5 * z[g] + 0.1 * comp);
    out[g] = f + 0.1 * noise;
  }
}
void launch_rand_field_impl(double* out, const double* x, const double* y, const double* z, const int64_t* lglel, int np1, size_t N1,
                            uint64_t seed, int comp, cudaStream_t st) {
  k_rand_field<<<std::min(cdiv(N1, 256), 148 * 8), 256, 0, st>>>(out, x, y, z, lglel, np1, N1, seed, comp); LAUNCH_COUNT();
}

}  // namespace nlk
