// Compile-time-sized tensor-product kernels (K3/K4/K5/K10/K15) for sm_100a.
// Each 1-D contraction is "pencil"-tiled in 3-D: one thread loads the MI inputs of a pencil into registers and produces
// all MO outputs (MO*MI FMAs per MI shared-memory loads, operator matrix read as a warp broadcast), so the kernels are
// bounded by HBM traffic of the element data rather than by shared-memory bandwidth.  In 2-D (few pencils) one thread
// per output point with fully unrolled loops is used.  Runtime-sized fallbacks live in nlk_kernels.cu.
#include "nlk_device.cuh"
#include "nlk_dmma.cuh"
#include <cstdlib>

namespace nlk {

// 1-D operators of the current (lx1, lxd) in constant memory: with fully unrolled (o, l) the matrix entry becomes a
// constant-bank operand of the DFMA instead of a shared-memory broadcast load (ncu r01: LSU/shared pipe was the limiter).
enum { MAT_SMEM = -1, MAT_I12 = 0, MAT_D12, MAT_I12T, MAT_D12T, MAT_I1D, MAT_I1DT, MAT_DD };
__constant__ double c_I12[16 * 16], c_D12[16 * 16], c_I12T[16 * 16], c_D12T[16 * 16], c_I1D[24 * 16], c_I1DT[24 * 16], c_DD[24 * 24];
static int c_ops_n = -1, c_ops_m = -1;
static void ensure_const_ops(const DevMesh& dm, cudaStream_t st) {
  if (c_ops_n == dm.n && c_ops_m == dm.m) return;
  const size_t qn = sizeof(double) * dm.q * dm.n, mn = sizeof(double) * dm.m * dm.n, mm = sizeof(double) * dm.m * dm.m;
  cudaMemcpyToSymbolAsync(c_I12, dm.I12, qn, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyToSymbolAsync(c_D12, dm.D12, qn, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyToSymbolAsync(c_I12T, dm.I12t, qn, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyToSymbolAsync(c_D12T, dm.D12t, qn, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyToSymbolAsync(c_I1D, dm.I1d, mn, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyToSymbolAsync(c_I1DT, dm.I1dt, mn, 0, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyToSymbolAsync(c_DD, dm.Dd, mm, 0, cudaMemcpyDeviceToDevice, st);
  c_ops_n = dm.n; c_ops_m = dm.m;
}
template <int MAT>
__device__ __forceinline__ double mat_at(const double* __restrict__ M, int idx) {
  if constexpr (MAT == MAT_I12) return c_I12[idx];
  else if constexpr (MAT == MAT_D12) return c_D12[idx];
  else if constexpr (MAT == MAT_I12T) return c_I12T[idx];
  else if constexpr (MAT == MAT_D12T) return c_D12T[idx];
  else if constexpr (MAT == MAT_I1D) return c_I1D[idx];
  else if constexpr (MAT == MAT_I1DT) return c_I1DT[idx];
  else if constexpr (MAT == MAT_DD) return c_DD[idx];
  else return M[idx];
}

// PI / PO: x-pitch (row length in doubles) of the input / output tile; 0 = dense.  An odd pitch makes the DIR = 0 stage
// (consecutive threads one row apart) free of shared-memory bank conflicts.
// (Sharing a pencil between two threads -- twice the warps, half the outputs each -- was measured slower: 3.4 -> 4.0 ms on
// the 3-D convection kernel, the register file then limits the SM to fewer resident blocks or forces spills.)
template <int MO, int MI, int DIR, bool ACC, int N0, int N1, int N2, int MAT = MAT_SMEM, int PI_ = 0, int PO_ = 0>
__device__ __forceinline__ void contract_t(double* __restrict__ out, const double* __restrict__ in, const double* __restrict__ M) {
  constexpr int O0 = DIR == 0 ? MO : N0, O1 = DIR == 1 ? MO : N1, O2 = DIR == 2 ? MO : N2;
  constexpr int P0 = DIR == 0 ? 1 : N0, P1 = DIR == 1 ? 1 : N1, P2 = DIR == 2 ? 1 : N2;
  constexpr int NPEN = P0 * P1 * P2;
  constexpr int PI = PI_ ? PI_ : N0, PO = PO_ ? PO_ : O0;
  constexpr int istr = DIR == 0 ? 1 : (DIR == 1 ? PI : PI * N1);
  constexpr int ostr = DIR == 0 ? 1 : (DIR == 1 ? PO : PO * O1);
  if constexpr (NPEN >= 32) {
    auto pencil = [&](int t) {
      const int a = t % P0, b = (t / P0) % P1, c = t / (P0 * P1);
      const int ibase = a + PI * (b + N1 * c), obase = a + PO * (b + O1 * c);
      double r[MI];
#pragma unroll
      for (int l = 0; l < MI; ++l) r[l] = in[ibase + l * istr];
#pragma unroll
      for (int o = 0; o < MO; ++o) {
        double s = 0;
#pragma unroll
        for (int l = 0; l < MI; ++l) s += mat_at<MAT>(M, o * MI + l) * r[l];
        if (ACC) out[obase + o * ostr] += s; else out[obase + o * ostr] = s;
      }
    };
    // Every 3-D launch in this file uses blockDim.x >= min(256, largest pencil count) (tp_threads), so up to 256 pencils
    // the "loop" is a single guarded trip.  Written as a loop, the compiler hoists the MO*MI loop-invariant operator
    // entries into uniform registers, runs out of them and spills (ncu: 77% of the issued instructions were
    // R2UR / MOV.SPILL / LDCU shuffles, 23% DFMA).
    if constexpr (NPEN <= 256) { if (threadIdx.x < NPEN) pencil(threadIdx.x); }
    else { for (int t = threadIdx.x; t < NPEN; t += blockDim.x) pencil(t); }
  } else {
    constexpr int TOT = O0 * O1 * O2;
    for (int idx = threadIdx.x; idx < TOT; idx += blockDim.x) {
      const int i = idx % O0, j = (idx / O0) % O1, k = idx / (O0 * O1);
      const int r = DIR == 0 ? i : (DIR == 1 ? j : k);
      const int base = (DIR == 0 ? 0 : i) + PI * ((DIR == 1 ? 0 : j) + N1 * (DIR == 2 ? 0 : k));
      const int oidx = i + PO * (j + O1 * k);
      double s = 0;
#pragma unroll
      for (int l = 0; l < MI; ++l) s += M[r * MI + l] * in[base + l * istr];   // (2-D path: per-thread row index -> shared memory)
      if (ACC) out[oidx] += s; else out[oidx] = s;
    }
  }
  __syncthreads();
}
// pointer selects instead of dynamic indexing into by-value pointer structs (which the compiler spills to local memory and then
// reads with generic LD.E loads)
__device__ __forceinline__ const double* sel3(const CPtr3& a, int c) { return c == 0 ? a.p[0] : (c == 1 ? a.p[1] : a.p[2]); }
__device__ __forceinline__ double* sel3(const Ptr3& a, int c) { return c == 0 ? a.p[0] : (c == 1 ? a.p[1] : a.p[2]); }
__device__ __forceinline__ const double* sel4(const CPtr4& a, int c) { return c == 0 ? a.p[0] : (c == 1 ? a.p[1] : (c == 2 ? a.p[2] : a.p[3])); }
__device__ __forceinline__ double* sel4(const Ptr4& a, int c) { return c == 0 ? a.p[0] : (c == 1 ? a.p[1] : (c == 2 ? a.p[2] : a.p[3])); }
__device__ __forceinline__ void load_mat_t(double* s, const double* g, int cnt) {
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) s[i] = g[i];
}
// Software L2 prefetch of a future element's operands: blocks are short-lived, so instead of an in-block pipeline every
// block pulls the lines that the block `dist` elements ahead will need into the 126 MB L2 (DRAM latency -> L2 latency).
__device__ __forceinline__ void prefetch_l2(const double* base, int ndoubles, int t, int nthreads) {
  const char* p = reinterpret_cast<const char*>(base);
  for (int off = t * 128; off < ndoubles * 8; off += nthreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}
constexpr int PF_DIST = 296;     // elements ahead (~2 CTAs per SM)
// 3-D block size: one thread per pencil of the largest stage, at most 256
__host__ __device__ constexpr int tp_block_threads(int nmax) { return (nmax * nmax + 31) / 32 * 32 > 256 ? 256 : (nmax * nmax + 31) / 32 * 32; }
static inline int tp_threads(int nmax, int d, int np) {
  if (d == 3) return tp_block_threads(nmax);
  int t = ((np + 31) / 32) * 32; return t > 256 ? 256 : (t < 64 ? 64 : t);
}

// ------------------------------------------------------------------------------------------------ K5 opdiv
template <int N, int DIM>
__global__ void k_opdiv_t(CPtr3 u, double* __restrict__ p, const double* __restrict__ rxw2, const double* __restrict__ I12g,
                          const double* __restrict__ D12g, double scale, const double* __restrict__ in_mul, CPtr3 in_mask,
                          const double* __restrict__ out_mul) {
  constexpr int n = N, q = N - 2, d = DIM, nz = DIM == 3 ? N : 1, qz = DIM == 3 ? q : 1;
  constexpr int np1 = n * n * nz, np2 = q * q * qz;
  extern __shared__ double sm[];
  double* sI = sm; double* sDm = sI + q * n;
  double* U = sDm + q * n; double* A = U + np1; double* B = A + np1; double* T0 = B + np1; double* T1 = T0 + np1; double* T2 = T1 + np1;
  double* acc = T2 + np1;
  const size_t e = blockIdx.x;
  load_mat_t(sI, I12g, q * n); load_mat_t(sDm, D12g, q * n);
  for (int i = threadIdx.x; i < np2; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const double* rw = rxw2 + e * (size_t)(d * d) * np2;
#pragma unroll 1
  for (int c = 0; c < d; ++c) {
    const double* uc = u.p[c] + e * np1;
    if (in_mul) { const double* mk = in_mask.p[c] + e * np1; const double* bi = in_mul + e * np1; for (int i = threadIdx.x; i < np1; i += blockDim.x) U[i] = uc[i] * bi[i] * mk[i]; }
    else for (int i = threadIdx.x; i < np1; i += blockDim.x) U[i] = uc[i];
    __syncthreads();
    contract_t<q, n, 0, false, n, n, nz, MAT_D12>(A, U, sDm);
    contract_t<q, n, 0, false, n, n, nz, MAT_I12>(B, U, sI);
    if constexpr (DIM == 2) {
      contract_t<q, n, 1, false, q, n, 1, MAT_I12>(T0, A, sI);
      contract_t<q, n, 1, false, q, n, 1, MAT_D12>(T1, B, sDm);
      for (int i = threadIdx.x; i < np2; i += blockDim.x) acc[i] += rw[(0 * d + c) * np2 + i] * T0[i] + rw[(1 * d + c) * np2 + i] * T1[i];
      __syncthreads();
    } else {
      contract_t<q, n, 1, false, q, n, n, MAT_I12>(T0, A, sI);
      contract_t<q, n, 1, false, q, n, n, MAT_D12>(T1, B, sDm);
      contract_t<q, n, 1, false, q, n, n, MAT_I12>(T2, B, sI);
      contract_t<q, n, 2, false, q, q, n, MAT_I12>(A, T0, sI);
      contract_t<q, n, 2, false, q, q, n, MAT_I12>(B, T1, sI);
      contract_t<q, n, 2, false, q, q, n, MAT_D12>(U, T2, sDm);
      for (int i = threadIdx.x; i < np2; i += blockDim.x)
        acc[i] += rw[(0 * d + c) * np2 + i] * A[i] + rw[(1 * d + c) * np2 + i] * B[i] + rw[(2 * d + c) * np2 + i] * U[i];
      __syncthreads();
    }
  }
  if (out_mul) for (int i = threadIdx.x; i < np2; i += blockDim.x) p[e * np2 + i] = scale * acc[i] * out_mul[e * np2 + i];
  else for (int i = threadIdx.x; i < np2; i += blockDim.x) p[e * np2 + i] = scale * acc[i];
}

// ------------------------------------------------------------------------------------------------ K5 opgradt
template <int N, int DIM>
__global__ void k_opgradt_t(const double* __restrict__ p, Ptr3 w, const double* __restrict__ rxw2, const double* __restrict__ I12tg,
                            const double* __restrict__ D12tg) {
  constexpr int n = N, q = N - 2, d = DIM, nz = DIM == 3 ? N : 1, qz = DIM == 3 ? q : 1;
  constexpr int np1 = n * n * nz, np2 = q * q * qz;
  extern __shared__ double sm[];
  double* sIt = sm; double* sDt = sIt + q * n;
  double* P = sDt + q * n; double* S0 = P + np1; double* S1 = S0 + np1; double* S2 = S1 + np1; double* A0 = S2 + np1; double* A1 = A0 + np1;
  double* A2 = A1 + np1;
  const size_t e = blockIdx.x;
  load_mat_t(sIt, I12tg, q * n); load_mat_t(sDt, D12tg, q * n);
  for (int i = threadIdx.x; i < np2; i += blockDim.x) P[i] = p[e * np2 + i];
  __syncthreads();
  const double* rw = rxw2 + e * (size_t)(d * d) * np2;
#pragma unroll 1
  for (int c = 0; c < d; ++c) {
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
      S0[i] = P[i] * rw[(0 * d + c) * np2 + i];
      S1[i] = P[i] * rw[(1 * d + c) * np2 + i];
      if (d == 3) S2[i] = P[i] * rw[(2 * d + c) * np2 + i];
    }
    __syncthreads();
    double* wc = w.p[c] + e * np1;
    if constexpr (DIM == 2) {
      contract_t<n, q, 0, false, q, q, 1, MAT_D12T>(A0, S0, sDt);
      contract_t<n, q, 0, false, q, q, 1, MAT_I12T>(A1, S1, sIt);
      contract_t<n, q, 1, false, n, q, 1, MAT_I12T>(S0, A0, sIt);
      contract_t<n, q, 1, true, n, q, 1, MAT_D12T>(S0, A1, sDt);
      for (int i = threadIdx.x; i < np1; i += blockDim.x) wc[i] = S0[i];
      __syncthreads();
    } else {
      contract_t<n, q, 0, false, q, q, q, MAT_D12T>(A0, S0, sDt);
      contract_t<n, q, 0, false, q, q, q, MAT_I12T>(A1, S1, sIt);
      contract_t<n, q, 0, false, q, q, q, MAT_I12T>(A2, S2, sIt);
      contract_t<n, q, 1, false, n, q, q, MAT_I12T>(S0, A0, sIt);
      contract_t<n, q, 1, true, n, q, q, MAT_D12T>(S0, A1, sDt);
      contract_t<n, q, 1, false, n, q, q, MAT_I12T>(S1, A2, sIt);
      contract_t<n, q, 2, false, n, n, q, MAT_I12T>(A0, S0, sIt);
      contract_t<n, q, 2, true, n, n, q, MAT_D12T>(A0, S1, sDt);
      for (int i = threadIdx.x; i < np1; i += blockDim.x) wc[i] = A0[i];
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------ K5, 3-D stage-fused versions
// The independent contractions of one sum-factorisation stage run concurrently on different warp groups (3 x P threads),
// the last stage is fused with the metric contraction / the global store: 3 barriers per component instead of 8-9 and
// 5-6 element tiles in shared memory instead of 7 (higher occupancy).  Pencil = all points along the contracted direction.
__host__ __device__ constexpr int roundup32(int x) { return (x + 31) / 32 * 32; }

// one pencil: r[] (already loaded) -> MO outputs written with stride ostr; optional second operand pair accumulates
template <int MO, int MI, int MAT>
__device__ __forceinline__ void pen_apply(const double (&r)[MI], double (&o)[MO]) {
#pragma unroll
  for (int a = 0; a < MO; ++a) {
    double s = 0;
#pragma unroll
    for (int l = 0; l < MI; ++l) s += mat_at<MAT>(nullptr, a * MI + l) * r[l];
    o[a] += s;
  }
}

__device__ __forceinline__ void group_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Component-parallel layout: warp group g (P threads, named barrier g+1) carries velocity component g through all three
// sum-factorisation stages on its own shared-memory tiles; the three partial divergences are summed at the end.
template <int N, int MINB>
__global__ void __launch_bounds__(3 * roundup32(N * N), MINB)
k_opdiv3_t(CPtr3 u, double* __restrict__ p, const double* __restrict__ rxw2, double scale, const double* __restrict__ in_mul, CPtr3 in_mask,
           const double* __restrict__ out_mul) {
  constexpr int n = N, q = N - 2, np1 = n * n * n, np2 = q * q * q, P = roundup32(n * n);
  constexpr int qp = q | 1, npad = n | 1;            // odd pitches: conflict-free pencil accesses
  constexpr int TILE = npad * n * n + 2 * qp * n * n + 2 * q * q * n;
  extern __shared__ double sm[];
  const size_t e = blockIdx.x;
  const int tid = threadIdx.x, c = tid / P, t = tid - c * P;
  double* U = sm + c * TILE; double* A = U + npad * n * n; double* B = A + qp * n * n; double* T1 = B + qp * n * n; double* T2 = T1 + q * q * n;
  double* T0 = U;
  double* R = sm + 3 * TILE;                         // [3][np2] partial results
  const double* rw = rxw2 + e * (size_t)9 * np2;
  if (e + PF_DIST < gridDim.x) {
    const size_t en = e + PF_DIST;
    prefetch_l2((c == 0 ? u.p[0] : (c == 1 ? u.p[1] : u.p[2])) + en * np1, np1, t, P);
    if (in_mul) { prefetch_l2((c == 0 ? in_mask.p[0] : (c == 1 ? in_mask.p[1] : in_mask.p[2])) + en * np1, np1, t, P); if (c == 0) prefetch_l2(in_mul + en * np1, np1, t, P); }
    prefetch_l2(rxw2 + en * (size_t)9 * np2 + (size_t)c * 3 * np2, 3 * np2, t, P);
  }
  // metrics for the z stage, fetched first (consumed last)
  double m0[q], m1[q], m2[q];
  if (t < q * q) {
#pragma unroll
    for (int a = 0; a < q; ++a) { const int g = a * q * q + t; m0[a] = rw[(0 * 3 + c) * np2 + g]; m1[a] = rw[(1 * 3 + c) * np2 + g]; m2[a] = rw[(2 * 3 + c) * np2 + g]; }
  }
  {
    // all loads of the element are issued before the first use (ncu r02: the rolled loop `load, load, multiply, store` exposed one
    // DRAM latency per trip -- 53 % of the samples stalled on the long scoreboard at that multiply)
    constexpr int NI = (np1 + P - 1) / P;
    const double* __restrict__ uc = (c == 0 ? u.p[0] : (c == 1 ? u.p[1] : u.p[2])) + e * np1;
    double uv[NI];
#pragma unroll
    for (int q_ = 0; q_ < NI; ++q_) { const int i = t + q_ * P; uv[q_] = i < np1 ? __ldg(uc + i) : 0.0; }
    if (in_mul) {
      const double* __restrict__ mk = (c == 0 ? in_mask.p[0] : (c == 1 ? in_mask.p[1] : in_mask.p[2])) + e * np1; const double* __restrict__ bi = in_mul + e * np1;
      double bv[NI], mv[NI];
#pragma unroll
      for (int q_ = 0; q_ < NI; ++q_) { const int i = t + q_ * P; bv[q_] = i < np1 ? __ldg(bi + i) : 0.0; mv[q_] = i < np1 ? __ldg(mk + i) : 0.0; }
#pragma unroll
      for (int q_ = 0; q_ < NI; ++q_) uv[q_] *= bv[q_] * mv[q_];
    }
#pragma unroll
    for (int q_ = 0; q_ < NI; ++q_) { const int i = t + q_ * P; if (i < np1) U[(i / n) * npad + (i % n)] = uv[q_]; }
  }
  group_sync(c + 1, P);
  // ---- x stage: A = D12_x u, B = I12_x u
  if (t < n * n) {
    double r[n], o[q];
#pragma unroll
    for (int l = 0; l < n; ++l) r[l] = U[t * npad + l];
#pragma unroll
    for (int a = 0; a < q; ++a) o[a] = 0.0;
    pen_apply<q, n, MAT_D12>(r, o);
#pragma unroll
    for (int a = 0; a < q; ++a) { A[t * qp + a] = o[a]; o[a] = 0.0; }
    pen_apply<q, n, MAT_I12>(r, o);
#pragma unroll
    for (int a = 0; a < q; ++a) B[t * qp + a] = o[a];
  }
  group_sync(c + 1, P);
  // ---- y stage: T0 = I12_y A, T1 = D12_y B, T2 = I12_y B
  if (t < q * n) {
    const int iq = t % q, k = t / q;
    double r[n], o[q];
#pragma unroll
    for (int l = 0; l < n; ++l) r[l] = A[(l + n * k) * qp + iq];
#pragma unroll
    for (int a = 0; a < q; ++a) o[a] = 0.0;
    pen_apply<q, n, MAT_I12>(r, o);
#pragma unroll
    for (int a = 0; a < q; ++a) { T0[iq + q * (a + q * k)] = o[a]; o[a] = 0.0; }
#pragma unroll
    for (int l = 0; l < n; ++l) r[l] = B[(l + n * k) * qp + iq];
    pen_apply<q, n, MAT_D12>(r, o);
#pragma unroll
    for (int a = 0; a < q; ++a) { T1[iq + q * (a + q * k)] = o[a]; o[a] = 0.0; }
    pen_apply<q, n, MAT_I12>(r, o);
#pragma unroll
    for (int a = 0; a < q; ++a) T2[iq + q * (a + q * k)] = o[a];
  }
  group_sync(c + 1, P);
  // ---- z stage fused with the metric contraction
  if (t < q * q) {
    double r[n], d0[q], acc[q];
#pragma unroll
    for (int l = 0; l < n; ++l) r[l] = T0[t + q * q * l];
#pragma unroll
    for (int a = 0; a < q; ++a) d0[a] = 0.0;
    pen_apply<q, n, MAT_I12>(r, d0);
#pragma unroll
    for (int a = 0; a < q; ++a) { acc[a] = m0[a] * d0[a]; d0[a] = 0.0; }
#pragma unroll
    for (int l = 0; l < n; ++l) r[l] = T1[t + q * q * l];
    pen_apply<q, n, MAT_I12>(r, d0);
#pragma unroll
    for (int a = 0; a < q; ++a) { acc[a] += m1[a] * d0[a]; d0[a] = 0.0; }
#pragma unroll
    for (int l = 0; l < n; ++l) r[l] = T2[t + q * q * l];
    pen_apply<q, n, MAT_D12>(r, d0);
#pragma unroll
    for (int a = 0; a < q; ++a) R[c * np2 + a * q * q + t] = acc[a] + m2[a] * d0[a];
  }
  __syncthreads();
  for (int i = tid; i < np2; i += blockDim.x) {
    const double v = scale * (R[i] + R[np2 + i] + R[2 * np2 + i]);
    p[e * np2 + i] = out_mul ? v * out_mul[e * np2 + i] : v;
  }
}

// Component-parallel: warp group g produces w_g = D_g^T p.
template <int N, int MINB>
__global__ void __launch_bounds__(3 * roundup32(N * N), MINB)
k_opgradt3_t(const double* __restrict__ p, Ptr3 w, const double* __restrict__ rxw2) {
  constexpr int n = N, q = N - 2, np1 = n * n * n, np2 = q * q * q, P = roundup32(n * n);
  constexpr int npad = n | 1, qp = q | 1;
  constexpr int TILE = 3 * qp * q * q + 3 * npad * q * q + 2 * n * n * q;
  extern __shared__ double sm[];
  const size_t e = blockIdx.x;
  const int tid = threadIdx.x, c = tid / P, t = tid - c * P;
  double* S = sm + c * TILE; double* A0 = S + 3 * qp * q * q; double* A1 = A0 + npad * q * q; double* A2 = A1 + npad * q * q;
  double* B0 = A2 + npad * q * q; double* B1 = B0 + n * n * q;
  const double* rw = rxw2 + e * (size_t)9 * np2;
  const double* pe = p + e * np2;
  if (e + PF_DIST < gridDim.x) {
    const size_t en = e + PF_DIST;
    if (c == 0) prefetch_l2(p + en * np2, np2, t, P);
    prefetch_l2(rxw2 + en * (size_t)9 * np2 + (size_t)c * 3 * np2, 3 * np2, t, P);
  }
  // S_k = p * rxw2[k][c], k = 0..2, coalesced into padded tiles; every load issued before the first use (see k_opdiv3_t)
  {
    constexpr int NI = (np2 + P - 1) / P;
    double pv[NI], rv[3][NI];
#pragma unroll
    for (int q_ = 0; q_ < NI; ++q_) {
      const int r_ = t + q_ * P;
      pv[q_] = r_ < np2 ? __ldg(pe + r_) : 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) rv[k][q_] = r_ < np2 ? __ldg(rw + (size_t)(k * 3 + c) * np2 + r_) : 0.0;
    }
#pragma unroll
    for (int q_ = 0; q_ < NI; ++q_) {
      const int r_ = t + q_ * P;
      if (r_ < np2) {
#pragma unroll
        for (int k = 0; k < 3; ++k) S[k * qp * q * q + (r_ / q) * qp + (r_ % q)] = pv[q_] * rv[k][q_];
      }
    }
  }
  group_sync(c + 1, P);
  // ---- x stage: A0 = D12^T_x S0, A1 = I12^T_x S1, A2 = I12^T_x S2       pencils (jq,kq)
  if (t < q * q) {
    double r[q], o[n];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int l = 0; l < q; ++l) r[l] = S[k * qp * q * q + t * qp + l];
#pragma unroll
      for (int a = 0; a < n; ++a) o[a] = 0.0;
      if (k == 0) pen_apply<n, q, MAT_D12T>(r, o); else pen_apply<n, q, MAT_I12T>(r, o);
      double* dst = (k == 0 ? A0 : (k == 1 ? A1 : A2)) + t * npad;
#pragma unroll
      for (int a = 0; a < n; ++a) dst[a] = o[a];
    }
  }
  group_sync(c + 1, P);
  // ---- y stage: B0 = I12^T_y A0 + D12^T_y A1 ; B1 = I12^T_y A2            pencils (i,kq)
  if (t < n * q) {
    const int i = t % n, k = t / n;
    double r[q], o[n];
#pragma unroll
    for (int a = 0; a < n; ++a) o[a] = 0.0;
#pragma unroll
    for (int l = 0; l < q; ++l) r[l] = A0[(l + q * k) * npad + i];
    pen_apply<n, q, MAT_I12T>(r, o);
#pragma unroll
    for (int l = 0; l < q; ++l) r[l] = A1[(l + q * k) * npad + i];
    pen_apply<n, q, MAT_D12T>(r, o);
#pragma unroll
    for (int a = 0; a < n; ++a) { B0[i + n * (a + n * k)] = o[a]; o[a] = 0.0; }
#pragma unroll
    for (int l = 0; l < q; ++l) r[l] = A2[(l + q * k) * npad + i];
    pen_apply<n, q, MAT_I12T>(r, o);
#pragma unroll
    for (int a = 0; a < n; ++a) B1[i + n * (a + n * k)] = o[a];
  }
  group_sync(c + 1, P);
  // ---- z stage: w = I12^T_z B0 + D12^T_z B1, stored straight to HBM        pencils (i,j)
  if (t < n * n) {
    double r[q], o[n];
#pragma unroll
    for (int a = 0; a < n; ++a) o[a] = 0.0;
#pragma unroll
    for (int l = 0; l < q; ++l) r[l] = B0[t + n * n * l];
    pen_apply<n, q, MAT_I12T>(r, o);
#pragma unroll
    for (int l = 0; l < q; ++l) r[l] = B1[t + n * n * l];
    pen_apply<n, q, MAT_D12T>(r, o);
    double* wc = (c == 0 ? w.p[0] : (c == 1 ? w.p[1] : w.p[2])) + e * np1;
#pragma unroll
    for (int a = 0; a < n; ++a) wc[t + n * n * a] = o[a];
  }
}

// ------------------------------------------------------------------------------------------------ K3 convect
template <int N, int MD, int DIM>
__device__ __forceinline__ void convect_t_body(CPtr4 u, int nf, CPtr3 C, Ptr4 out, const double* __restrict__ rxd, const double* __restrict__ I1dg,
                            const double* __restrict__ I1dtg, const double* __restrict__ Ddg, double alpha, int accumulate) {
  constexpr int n = N, m = MD, d = DIM, nz = DIM == 3 ? N : 1, mz = DIM == 3 ? MD : 1;
  constexpr int np1 = n * n * nz, npd = m * m * mz;
  extern __shared__ double sm[];
  double* sI = sm; double* sIt = sI + m * n; double* sDd = sIt + m * n;
  // 3-D tiles carry odd x-pitches (PN, PM): the x-direction stages then run without shared-memory bank conflicts
  constexpr int PN = DIM == 3 ? (n | 1) : n, PM = DIM == 3 ? (m | 1) : m, npdP = PM * m * mz;
  double* TR = sDd + m * m; double* UF = TR + 3 * npdP; double* W1 = UF + npdP; double* W2 = W1 + npdP; double* ACC = W2 + npdP;
  const size_t e = blockIdx.x;
  // (no L2 prefetch here: at 161 KB per element the look-ahead of one resident wave does not survive in L2 -- ncu showed
  // a 10% hit rate and twice the DRAM reads.)  Instead the next field's nodal values are fetched into registers while the
  // current field is being contracted: PRE values per thread, NT = the block size tp_threads() launches.
  constexpr int NT = tp_block_threads(m);
  constexpr int PRE = (np1 + NT - 1) / NT;
  double pre[PRE];
  auto fetch = [&](const double* src) {
#pragma unroll
    for (int q = 0; q < PRE; ++q) { const int i = threadIdx.x + q * NT; pre[q] = i < np1 ? src[i] : 0.0; }
  };
  auto stage = [&]() {
#pragma unroll
    for (int q = 0; q < PRE; ++q) { const int i = threadIdx.x + q * NT; if (i < np1) W1[(i % n) + PN * (i / n)] = pre[q]; }
  };
  if constexpr (DIM == 3) fetch(sel3(C, 0) + e * np1);
  // pull THIS element's dealiasing metrics (d*d*m^3 doubles: 124 KB at m = 12, the bulk of the kernel's traffic) into L2 now;
  // they are consumed after the three interpolations of C, a few microseconds from here (the r01 look-ahead prefetch of FUTURE
  // elements was removed because it did not survive in L2; this one is used by the same block almost immediately)
  if constexpr (DIM == 3) prefetch_l2(rxd + e * (size_t)(d * d) * npd, d * d * npd, threadIdx.x, blockDim.x);
  load_mat_t(sI, I1dg, m * n); load_mat_t(sIt, I1dtg, m * n); load_mat_t(sDd, Ddg, m * m);
  __syncthreads();
#pragma unroll 1
  for (int c = 0; c < d; ++c) {
    if constexpr (DIM == 2) {
      const double* cc = sel3(C, c) + e * np1;
      for (int i = threadIdx.x; i < np1; i += blockDim.x) W1[i] = cc[i];
      __syncthreads();
      contract_t<m, n, 0, false, n, n, 1, MAT_I1D>(W2, W1, sI);
      contract_t<m, n, 1, false, m, n, 1, MAT_I1D>(TR + c * npd, W2, sI);
    } else {
      stage();
      __syncthreads();
      fetch((c + 1 < d ? sel3(C, c + 1) : sel4(u, 0)) + e * np1);
      contract_t<m, n, 0, false, n, n, n, MAT_I1D, PN, PM>(W2, W1, sI);
      contract_t<m, n, 1, false, m, n, n, MAT_I1D, PM, PM>(W1, W2, sI);
      contract_t<m, n, 2, false, m, m, n, MAT_I1D, PM, PM>(TR + c * npdP, W1, sI);
    }
  }
  const double* rx = rxd + e * (size_t)(d * d) * npd;
#pragma unroll 2
  for (int i = threadIdx.x; i < npd; i += blockDim.x) {
    const int ip = (i % m) + PM * (i / m);
    double cf[3] = {TR[ip], TR[npdP + ip], d == 3 ? TR[2 * npdP + ip] : 0.0};
#pragma unroll
    for (int k = 0; k < d; ++k) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < d; ++c) s += rx[(k * d + c) * npd + i] * cf[c];
      TR[k * npdP + ip] = s;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int f = 0; f < nf; ++f) {
    if constexpr (DIM == 2) {
      const double* uf = sel4(u, f) + e * np1;
      for (int i = threadIdx.x; i < np1; i += blockDim.x) W1[i] = uf[i];
    } else {
      stage();
      if (f + 1 < nf) fetch(sel4(u, f + 1) + e * np1);
    }
    __syncthreads();
    if constexpr (DIM == 2) {
      contract_t<m, n, 0, false, n, n, 1, MAT_I1D>(W2, W1, sI);
      contract_t<m, n, 1, false, m, n, 1, MAT_I1D>(UF, W2, sI);
      contract_t<m, m, 0, false, m, m, 1, MAT_DD>(W1, UF, sDd);
      contract_t<m, m, 1, false, m, m, 1, MAT_DD>(W2, UF, sDd);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) ACC[i] = TR[i] * W1[i] + TR[npd + i] * W2[i];
      __syncthreads();
      contract_t<n, m, 0, false, m, m, 1, MAT_I1DT>(W1, ACC, sIt);
      contract_t<n, m, 1, false, n, m, 1, MAT_I1DT>(W2, W1, sIt);
    } else {
      // the padded slots of the tiles are never read by a contraction, so the element-wise passes run over the whole
      // padded range without index arithmetic
      contract_t<m, n, 0, false, n, n, n, MAT_I1D, PN, PM>(W2, W1, sI);
      contract_t<m, n, 1, false, m, n, n, MAT_I1D, PM, PM>(W1, W2, sI);
      contract_t<m, n, 2, false, m, m, n, MAT_I1D, PM, PM>(UF, W1, sI);
      contract_t<m, m, 0, false, m, m, m, MAT_DD, PM, PM>(W1, UF, sDd);
      for (int i = threadIdx.x; i < npdP; i += blockDim.x) ACC[i] = TR[i] * W1[i];
      __syncthreads();
      contract_t<m, m, 1, false, m, m, m, MAT_DD, PM, PM>(W1, UF, sDd);
      for (int i = threadIdx.x; i < npdP; i += blockDim.x) ACC[i] += TR[npdP + i] * W1[i];
      __syncthreads();
      contract_t<m, m, 2, false, m, m, m, MAT_DD, PM, PM>(W1, UF, sDd);
      for (int i = threadIdx.x; i < npdP; i += blockDim.x) ACC[i] += TR[2 * npdP + i] * W1[i];
      __syncthreads();
      contract_t<n, m, 0, false, m, m, m, MAT_I1DT, PM, PN>(W1, ACC, sIt);
      contract_t<n, m, 1, false, n, m, m, MAT_I1DT, PN, PN>(UF, W1, sIt);
      contract_t<n, m, 2, false, n, n, m, MAT_I1DT, PN, PN>(W2, UF, sIt);
    }
    double* of = sel4(out, f) + e * np1;
    for (int i = threadIdx.x; i < np1; i += blockDim.x) of[i] = (accumulate ? of[i] : 0.0) + alpha * W2[(i % n) + PN * (i / n)];
    __syncthreads();
  }
}

template <int N, int MD, int DIM>
__global__ void k_convect_t(CPtr4 u, int nf, CPtr3 C, Ptr4 out, const double* __restrict__ rxd, const double* __restrict__ I1dg,
                            const double* __restrict__ I1dtg, const double* __restrict__ Ddg, double alpha, int accumulate) {
  convect_t_body<N, MD, DIM>(u, nf, C, out, rxd, I1dg, I1dtg, Ddg, alpha, accumulate);
}
// the same body under explicit launch bounds (A/B: NLK_CONVECT_LB=1|2 -> minBlocks 1|2)
template <int N, int MD, int DIM, int MINB>
__global__ void __launch_bounds__(tp_block_threads(MD), MINB)
k_convect_tb(CPtr4 u, int nf, CPtr3 C, Ptr4 out, const double* __restrict__ rxd, const double* __restrict__ I1dg,
             const double* __restrict__ I1dtg, const double* __restrict__ Ddg, double alpha, int accumulate) {
  convect_t_body<N, MD, DIM>(u, nf, C, out, rxd, I1dg, I1dtg, Ddg, alpha, accumulate);
}

// ------------------------------------------------------------------------------------------------ K4 convect_adj
template <int N, int MD, int DIM>
__global__ void k_convect_adj_t(CPtr3 U, CPtr3 cf, Ptr3 out, const double* __restrict__ rxd, const double* __restrict__ I1dg,
                                const double* __restrict__ I1dtg, const double* __restrict__ Ddg, double alpha, int accumulate, int nj) {
  constexpr int n = N, m = MD, d = DIM, nz = DIM == 3 ? N : 1, mz = DIM == 3 ? MD : 1;
  constexpr int np1 = n * n * nz, npd = m * m * mz;
  constexpr int PN = DIM == 3 ? (n | 1) : n, PM = DIM == 3 ? (m | 1) : m, npdP = PM * m * mz;   // odd x-pitches, see k_convect_t
  extern __shared__ double sm[];
  double* sI = sm; double* sIt = sI + m * n; double* sDd = sIt + m * n;
  double* AC = sDd + m * m; double* UF = AC + 3 * npdP; double* W1 = UF + npdP; double* W2 = W1 + npdP; double* CF = W2 + npdP;
  const size_t e = blockIdx.x;
  load_mat_t(sI, I1dg, m * n); load_mat_t(sIt, I1dtg, m * n); load_mat_t(sDd, Ddg, m * m);
  for (int i = threadIdx.x; i < d * npdP; i += blockDim.x) AC[i] = 0.0;
  __syncthreads();
  const double* rx = rxd + e * (size_t)(d * d) * npd;
#pragma unroll 1
  for (int j = 0; j < nj; ++j) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const double* src = (pass == 0 ? sel3(cf, j) : sel3(U, j)) + e * np1;
      double* dst = pass == 0 ? CF : UF;
      for (int i = threadIdx.x; i < np1; i += blockDim.x) W1[(i % n) + PN * (i / n)] = src[i];
      __syncthreads();
      if constexpr (DIM == 2) { contract_t<m, n, 0, false, n, n, 1, MAT_I1D>(W2, W1, sI); contract_t<m, n, 1, false, m, n, 1, MAT_I1D>(dst, W2, sI); }
      else {
        contract_t<m, n, 0, false, n, n, n, MAT_I1D, PN, PM>(W2, W1, sI);
        contract_t<m, n, 1, false, m, n, n, MAT_I1D, PM, PM>(W1, W2, sI);
        contract_t<m, n, 2, false, m, m, n, MAT_I1D, PM, PM>(dst, W1, sI);
      }
    }
#pragma unroll
    for (int k = 0; k < d; ++k) {
      if (k == 0) contract_t<m, m, 0, false, m, m, mz, MAT_DD, PM, PM>(W1, UF, sDd);
      else if (k == 1) contract_t<m, m, 1, false, m, m, mz, MAT_DD, PM, PM>(W1, UF, sDd);
      else contract_t<m, m, 2, false, m, m, mz, MAT_DD, PM, PM>(W1, UF, sDd);
      for (int i = threadIdx.x; i < npd; i += blockDim.x) {
        const int ip = (i % m) + PM * (i / m);
        double g = CF[ip] * W1[ip];
#pragma unroll
        for (int c = 0; c < d; ++c) AC[c * npdP + ip] += rx[(k * d + c) * npd + i] * g;
      }
      __syncthreads();
    }
  }
#pragma unroll 1
  for (int c = 0; c < d; ++c) {
    if constexpr (DIM == 2) { contract_t<n, m, 0, false, m, m, 1, MAT_I1DT>(W1, AC + c * npd, sIt); contract_t<n, m, 1, false, n, m, 1, MAT_I1DT>(W2, W1, sIt); }
    else {
      contract_t<n, m, 0, false, m, m, m, MAT_I1DT, PM, PN>(W1, AC + c * npdP, sIt);
      contract_t<n, m, 1, false, n, m, m, MAT_I1DT, PN, PN>(UF, W1, sIt);
      contract_t<n, m, 2, false, n, n, m, MAT_I1DT, PN, PN>(W2, UF, sIt);
    }
    double* of = sel3(out, c) + e * np1;
    for (int i = threadIdx.x; i < np1; i += blockDim.x) of[i] = (accumulate ? of[i] : 0.0) + alpha * W2[(i % n) + PN * (i / n)];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ K10 Schwarz FDM
__device__ __forceinline__ int clamp_inner_t(int i, int n) { return i == 0 ? 1 : (i == n - 1 ? n - 2 : i); }
template <int N, int DIM>
__global__ void k_schwarz_fdm_t(const double* __restrict__ w, double* __restrict__ z, double* __restrict__ t, const double* __restrict__ S,
                                const double* __restrict__ St, const double* __restrict__ dinv) {
  constexpr int n = N, d = DIM, nz = DIM == 3 ? N : 1, np1 = n * n * nz, nn = n * n;
  constexpr int PN = DIM == 3 ? (n | 1) : n, npP = PN * n * nz;       // odd x-pitch in 3-D (bank-conflict-free x stage)
  extern __shared__ double sm[];
  double* sS = sm; double* sSt = sS + d * nn; double* A = sSt + d * nn; double* B = A + npP;
  const size_t e = blockIdx.x;
  if (DIM == 3 && e + PF_DIST < gridDim.x) {
    const size_t en = e + PF_DIST;
    prefetch_l2(w + en * np1, np1, threadIdx.x, blockDim.x); prefetch_l2(dinv + en * np1, np1, threadIdx.x, blockDim.x);
    prefetch_l2(S + en * (size_t)d * nn, d * nn, threadIdx.x, blockDim.x); prefetch_l2(St + en * (size_t)d * nn, d * nn, threadIdx.x, blockDim.x);
  }
  load_mat_t(sS, S + e * (size_t)d * nn, d * nn); load_mat_t(sSt, St + e * (size_t)d * nn, d * nn);
  for (int p = threadIdx.x; p < np1; p += blockDim.x) B[(p % n) + PN * (p / n)] = w[e * np1 + p];
  __syncthreads();
  for (int p = threadIdx.x; p < np1; p += blockDim.x) {
    int i = p % n, j = (p / n) % n, k = d == 3 ? p / nn : 1;
    int nb = (i == 0 || i == n - 1) + (j == 0 || j == n - 1) + (d == 3 ? (k == 0 || k == n - 1) : 0);
    double v = B[i + PN * (p / n)];
    if (nb == 1) { int ii = clamp_inner_t(i, n), jj = clamp_inner_t(j, n), kk = d == 3 ? clamp_inner_t(k, n) : 0; v -= B[(kk * n + jj) * PN + ii]; }
    A[i + PN * (p / n)] = v;
  }
  __syncthreads();
  double* res;
  contract_t<n, n, 0, false, n, n, nz, MAT_SMEM, PN, PN>(B, A, sSt);
  contract_t<n, n, 1, false, n, n, nz, MAT_SMEM, PN, PN>(A, B, sSt + nn);
  if constexpr (DIM == 3) {
    contract_t<n, n, 2, false, n, n, nz, MAT_SMEM, PN, PN>(B, A, sSt + 2 * nn);
    for (int p = threadIdx.x; p < np1; p += blockDim.x) B[(p % n) + PN * (p / n)] *= dinv[e * np1 + p];
    __syncthreads();
    contract_t<n, n, 0, false, n, n, nz, MAT_SMEM, PN, PN>(A, B, sS);
    contract_t<n, n, 1, false, n, n, nz, MAT_SMEM, PN, PN>(B, A, sS + nn);
    contract_t<n, n, 2, false, n, n, nz, MAT_SMEM, PN, PN>(A, B, sS + 2 * nn);
    res = A;
  } else {
    for (int p = threadIdx.x; p < np1; p += blockDim.x) A[p] *= dinv[e * np1 + p];
    __syncthreads();
    contract_t<n, n, 0, false, n, n, nz>(B, A, sS);
    contract_t<n, n, 1, false, n, n, nz>(A, B, sS + nn);
    res = A;
  }
  for (int p = threadIdx.x; p < np1; p += blockDim.x) {
    int i = p % n, j = (p / n) % n, k = d == 3 ? p / nn : 1;
    int nb = (i == 0 || i == n - 1) + (j == 0 || j == n - 1) + (d == 3 ? (k == 0 || k == n - 1) : 0);
    double v = res[i + PN * (p / n)];
    z[e * np1 + p] = v;
    t[e * np1 + p] = nb == 1 ? v : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------ K10 fused Schwarz (pull tables)
// The overlap exchange of the additive Schwarz smoother only ever moves FACE-INTERIOR values between the two elements that
// share a face (multiplicity 2), so "dssum, then subtract the own copy" is a pairwise fetch.  With two precomputed index
// tables per (element, face, slot) -- t1: where the neighbour's first interior pressure layer lives in the residual array,
// t2: where the neighbour's extended-layer FDM value lives in the compact face array ZF (negative = ghost slot of another
// rank, -1 = no neighbour) -- the five-kernel chain embed -> dssum -> fdm -> dssum -> gather over three velocity-sized
// work arrays becomes two kernels over pressure-sized ones:
//   A: z = (S (x) S (x) S) dinv (S^T (x) S^T (x) S^T) [r_e extended by the neighbours' layers]  -> zint (q^d), ZF (2d x q^(d-1))
//   B: out = wt * (zint + neighbours' ZF on the first interior layer) [+ coarse prolongation]
template <int N, int DIM>
__device__ __forceinline__ void swf_face_of(int i, int j, int k, int& f, int& s, int& nb) {
  constexpr int q = N - 2;
  const bool bi = (i == 0 || i == N - 1), bj = (j == 0 || j == N - 1), bk = DIM == 3 && (k == 0 || k == N - 1);
  nb = (int)bi + (int)bj + (int)bk;
  if (bi) { f = (i == 0 ? 0 : 1); s = (j - 1) + q * (k - (DIM == 3 ? 1 : 0)); }
  else if (bj) { f = 2 + (j == 0 ? 0 : 1); s = (i - 1) + q * (k - (DIM == 3 ? 1 : 0)); }
  else { f = 4 + (k == 0 ? 0 : 1); s = (i - 1) + q * (j - 1); }
}
template <int N, int DIM>
__global__ void k_swf_a(const double* __restrict__ r, const double* __restrict__ mul, const int32_t* __restrict__ t1, const double* __restrict__ ghost,
                        const double* __restrict__ S, const double* __restrict__ St, const double* __restrict__ dinv,
                        double* __restrict__ zint, double* __restrict__ ZF) {
  constexpr int n = N, q = N - 2, d = DIM, nz = DIM == 3 ? N : 1, np1 = n * n * nz, nn = n * n, qz = DIM == 3 ? q : 1, np2 = q * q * qz;
  constexpr int NF = 2 * DIM, FS = DIM == 3 ? q * q : q;
  constexpr int PN = DIM == 3 ? (n | 1) : n, npP = PN * n * nz;
  extern __shared__ double sm[];
  double* sS = sm; double* sSt = sS + d * nn; double* A = sSt + d * nn; double* B = A + npP;
  const size_t e = blockIdx.x;
  load_mat_t(sS, S + e * (size_t)d * nn, d * nn); load_mat_t(sSt, St + e * (size_t)d * nn, d * nn);
  const int32_t* te = t1 + e * (size_t)(NF * FS);
  for (int p = threadIdx.x; p < np1; p += blockDim.x) {
    const int i = p % n, j = (p / n) % n, k = DIM == 3 ? p / nn : 0;
    int f, s_, nb; swf_face_of<N, DIM>(i, j, k, f, s_, nb);
    double v = 0.0;
    if (nb == 0) { const size_t g = e * np2 + (size_t)((DIM == 3 ? (k - 1) * q : 0) + (j - 1)) * q + (i - 1); v = mul ? r[g] * mul[g] : r[g]; }
    else if (nb == 1) { const int idx = te[f * FS + s_]; if (idx >= 0) v = mul ? r[idx] * mul[idx] : r[idx]; else if (idx <= -2) v = ghost[-2 - idx]; }
    A[i + PN * (p / n)] = v;
  }
  __syncthreads();
  double* res;
  contract_t<n, n, 0, false, n, n, nz, MAT_SMEM, PN, PN>(B, A, sSt);
  contract_t<n, n, 1, false, n, n, nz, MAT_SMEM, PN, PN>(A, B, sSt + nn);
  if constexpr (DIM == 3) {
    contract_t<n, n, 2, false, n, n, nz, MAT_SMEM, PN, PN>(B, A, sSt + 2 * nn);
    for (int p = threadIdx.x; p < np1; p += blockDim.x) B[(p % n) + PN * (p / n)] *= dinv[e * np1 + p];
    __syncthreads();
    contract_t<n, n, 0, false, n, n, nz, MAT_SMEM, PN, PN>(A, B, sS);
    contract_t<n, n, 1, false, n, n, nz, MAT_SMEM, PN, PN>(B, A, sS + nn);
    contract_t<n, n, 2, false, n, n, nz, MAT_SMEM, PN, PN>(A, B, sS + 2 * nn);
    res = A;
  } else {
    for (int p = threadIdx.x; p < np1; p += blockDim.x) A[p] *= dinv[e * np1 + p];
    __syncthreads();
    contract_t<n, n, 0, false, n, n, nz>(B, A, sS);
    contract_t<n, n, 1, false, n, n, nz>(A, B, sS + nn);
    res = A;
  }
  for (int p = threadIdx.x; p < np1; p += blockDim.x) {
    const int i = p % n, j = (p / n) % n, k = DIM == 3 ? p / nn : 0;
    int f, s_, nb; swf_face_of<N, DIM>(i, j, k, f, s_, nb);
    const double v = res[i + PN * (p / n)];
    if (nb == 0) zint[e * np2 + (size_t)((DIM == 3 ? (k - 1) * q : 0) + (j - 1)) * q + (i - 1)] = v;
    else if (nb == 1) ZF[e * (size_t)(NF * FS) + f * FS + s_] = v;
  }
}
template <int N, int DIM>
__global__ void k_swf_b(const double* __restrict__ zint, const double* __restrict__ ZF, const int32_t* __restrict__ t2, const double* __restrict__ ghost,
                        const double* __restrict__ wt, const double* __restrict__ yc, const int64_t* __restrict__ vertex,
                        const double* __restrict__ z2, double* __restrict__ out, size_t N2) {
  constexpr int q = N - 2, qz = DIM == 3 ? q : 1, np2 = q * q * qz, NF = 2 * DIM, FS = DIM == 3 ? q * q : q, NV = 1 << DIM;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < N2; g += (size_t)gridDim.x * blockDim.x) {
    const size_t e = g / np2; const int p = (int)(g - e * np2);
    const int i = p % q, j = (p / q) % q, k = DIM == 3 ? p / (q * q) : 0;
    const int32_t* te = t2 + e * (size_t)(NF * FS);
    double v = zint[g];
    auto add = [&](int f, int s_) { const int idx = te[f * FS + s_]; if (idx >= 0) v += ZF[idx]; else if (idx <= -2) v += ghost[-2 - idx]; };
    if (i == 0) add(0, j + q * k);
    if (i == q - 1) add(1, j + q * k);
    if (j == 0) add(2, i + q * k);
    if (j == q - 1) add(3, i + q * k);
    if (DIM == 3) { if (k == 0) add(4, i + q * j); if (k == q - 1) add(5, i + q * j); }
    v *= wt[g];
    if (yc) {
      const double hx1 = 0.5 * (1 + z2[i]), hy1 = 0.5 * (1 + z2[j]), hz1 = DIM == 3 ? 0.5 * (1 + z2[k]) : 1.0;
      const int64_t* ve = vertex + e * NV;
      double s = 0;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double h = ((c & 1) ? hx1 : 1 - hx1) * (((c >> 1) & 1) ? hy1 : 1 - hy1) * (DIM == 3 ? (((c >> 2) & 1) ? hz1 : 1 - hz1) : 1.0);
        s += h * yc[ve[c] - 1];
      }
      v += s;
    }
    out[g] = v;
  }
}
// ---- FP64 tensor-core (DMMA) version of kernel A for lx1 = 8 in 3-D: one WARP per element.
// The six contractions of the FDM solve are (8 x 8) x (8 x 64) products.  With scalar FMAs every one of the 24.6 k FMAs of an
// element needs its own shared-memory operand (the S matrices differ from element to element, so the constant-bank trick of
// the fixed-operator kernels does not apply): ncu r01/r02 show the LSU pipe, not DRAM or the FP64 pipe, as the limiter
// (168 registers, 18 % of the warps active).  mma.sync.m8n8k4.f64 takes the operator as a 2-register A fragment held for a
// whole contraction and the data as one 8-byte shared-memory load per lane and 256 FMAs: 16 DMMA + 16 LDS + 16 STS per
// contraction and warp instead of ~1 000 LDS.  FP64 has no tcgen05 path; DMMA is the FP64 tensor pipe of sm_100a.
constexpr int SWF8_WARPS = 4;
template <int MINB>
__global__ void __launch_bounds__(32 * SWF8_WARPS, MINB)
k_swf_a8(const double* __restrict__ r, const double* __restrict__ mul, const int32_t* __restrict__ t1, const double* __restrict__ ghost,
         const double* __restrict__ S, const double* __restrict__ St, const double* __restrict__ dinv,
         double* __restrict__ zint, double* __restrict__ ZF, int64_t E) {
  constexpr int n = 8, q = 6, np1 = 512, nn = 64, np2 = 216, NF = 6, FS = 36, PN = 9, npP = PN * 64;
  __shared__ double tiles[SWF8_WARPS][2][npP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * SWF8_WARPS + w;
  if (e >= E) return;
  double* A = tiles[w][0]; double* B = tiles[w][1];
  // operator fragments: A-fragment of M is M[lane>>2][kb*4 + (lane&3)]
  const double* Se = S + e * (size_t)(3 * nn); const double* Ste = St + e * (size_t)(3 * nn);
  const int fr = (lane >> 2) * n + (lane & 3);
  double st0[3], st1[3], s0[3], s1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { st0[k] = Ste[k * nn + fr]; st1[k] = Ste[k * nn + fr + 4]; s0[k] = Se[k * nn + fr]; s1[k] = Se[k * nn + fr + 4]; }
  const int32_t* te = t1 + e * (size_t)(NF * FS);
  {
    // two batched load phases (all 16 index loads, then all 16 value loads) before the first shared-memory store: a rolled loop
    // exposed two dependent DRAM latencies per trip (ncu r02: 60 % of the samples on the long scoreboard)
    long long src[16];                       // >= 0: index into r ; -1: zero ; <= -2: ghost slot
#pragma unroll
    for (int q_ = 0; q_ < 16; ++q_) {
      const int p = lane + 32 * q_;
      const int i = p & 7, j = (p >> 3) & 7, k = p >> 6;
      int f, s_, nb; swf_face_of<8, 3>(i, j, k, f, s_, nb);
      src[q_] = -1;
      if (nb == 0) src[q_] = (long long)(e * np2 + (size_t)((k - 1) * q + (j - 1)) * q + (i - 1));
      else if (nb == 1) src[q_] = __ldg(te + f * FS + s_);
    }
    double v[16], m_[16];
#pragma unroll
    for (int q_ = 0; q_ < 16; ++q_) {
      v[q_] = 0.0; m_[q_] = 1.0;
      if (src[q_] >= 0) { v[q_] = __ldg(r + src[q_]); if (mul) m_[q_] = __ldg(mul + src[q_]); }
      else if (src[q_] <= -2) v[q_] = __ldg(ghost + (-2 - src[q_]));
    }
#pragma unroll
    for (int q_ = 0; q_ < 16; ++q_) { const int p = lane + 32 * q_; A[(p & 7) + PN * (p >> 3)] = v[q_] * m_[q_]; }
  }
  __syncwarp();
  double dv[16];                             // inverse eigenvalue sums, fetched now, consumed after the three forward contractions
#pragma unroll
  for (int q_ = 0; q_ < 16; ++q_) dv[q_] = __ldg(dinv + e * np1 + lane + 32 * q_);
  dmma_contract8<0, PN>(B, A, st0[0], st1[0], lane);
  dmma_contract8<1, PN>(A, B, st0[1], st1[1], lane);
  dmma_contract8<2, PN>(B, A, st0[2], st1[2], lane);
#pragma unroll
  for (int q_ = 0; q_ < 16; ++q_) { const int p = lane + 32 * q_; B[(p & 7) + PN * (p >> 3)] *= dv[q_]; }
  __syncwarp();
  dmma_contract8<0, PN>(A, B, s0[0], s1[0], lane);
  dmma_contract8<1, PN>(B, A, s0[1], s1[1], lane);
  dmma_contract8<2, PN>(A, B, s0[2], s1[2], lane);
  for (int p = lane; p < np1; p += 32) {
    const int i = p & 7, j = (p >> 3) & 7, k = p >> 6;
    int f, s_, nb; swf_face_of<8, 3>(i, j, k, f, s_, nb);
    const double v = A[i + PN * (p >> 3)];
    if (nb == 0) zint[e * np2 + (size_t)((k - 1) * q + (j - 1)) * q + (i - 1)] = v;
    else if (nb == 1) ZF[e * (size_t)(NF * FS) + f * FS + s_] = v;
  }
}

// coarse restriction, one warp per element: part[e][c] = sum_p shape_c(p) r[e][p] (* mul), coalesced
template <int DIM>
__global__ void k_coarse_part_w(const double* __restrict__ r, const double* __restrict__ mul, double* __restrict__ part, const double* __restrict__ z2, int q, int64_t E) {
  constexpr int NV = 1 << DIM;
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  const int np2 = DIM == 3 ? q * q * q : q * q;
  double acc[NV];
#pragma unroll
  for (int c = 0; c < NV; ++c) acc[c] = 0.0;
  for (int p = lane; p < np2; p += 32) {
    const size_t g = (size_t)e * np2 + p;
    const double v = mul ? r[g] * mul[g] : r[g];
    const int i = p % q, j = (p / q) % q, k = DIM == 3 ? p / (q * q) : 0;
    const double hx1 = 0.5 * (1 + z2[i]), hy1 = 0.5 * (1 + z2[j]), hz1 = DIM == 3 ? 0.5 * (1 + z2[k]) : 1.0;
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] += v * ((c & 1) ? hx1 : 1 - hx1) * (((c >> 1) & 1) ? hy1 : 1 - hy1) * (DIM == 3 ? (((c >> 2) & 1) ? hz1 : 1 - hz1) : 1.0);
  }
#pragma unroll
  for (int c = 0; c < NV; ++c) { double t = acc[c]; for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o); if (lane == 0) part[e * NV + c] = t; }
}
__global__ void k_swf_pack(const double* __restrict__ src, const double* __restrict__ mul, const int32_t* __restrict__ idx, int n, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { const int g = idx[t]; out[t] = mul ? src[g] * mul[g] : src[g]; }
}

// ------------------------------------------------------------------------------------------------ dispatch
#define TP_CASE_N(N_) \
  case N_ * 10 + 2: FN(N_, 2); break; \
  case N_ * 10 + 3: FN(N_, 3); break;
#define TP_SWITCH_N(key) switch (key) { TP_CASE_N(4) TP_CASE_N(5) TP_CASE_N(6) TP_CASE_N(7) TP_CASE_N(8) TP_CASE_N(9) TP_CASE_N(10) TP_CASE_N(11) TP_CASE_N(12) default: return false; }

template <class K> static void set_smem(K kern, size_t smem) { if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }

bool tp_opdiv(const DevMesh& dm, CPtr3 u, double* p, double scale, const double* in_mul, const double* out_mul, cudaStream_t st) {
  size_t smem = (size_t)(2 * dm.q * dm.n + 7 * dm.np1) * sizeof(double);
  if (smem > 220 * 1024) return false;
  CPtr3 mk{{dm.mask[0], dm.mask[1], dm.mask[2]}};
  ensure_const_ops(dm, st);
  if (dm.ndim == 3) {
    const int n = dm.n, q = dm.q;
    size_t sm3 = (size_t)(3 * ((n | 1) * n * n + 2 * (q | 1) * n * n + 2 * q * q * n) + 3 * q * q * q) * sizeof(double);
    int thr3 = 3 * roundup32(n * n);
    // minBlocks = 1 in __launch_bounds__ (108 registers instead of 110, same 3 blocks per SM, different instruction schedule):
    // 313.5 -> 244.1 us at 24k elements, measured A/B in one run (NLK_OPDIV_MINB0=1 selects the old code)
    static const int minb0 = getenv("NLK_OPDIV_MINB0") != nullptr;
    static const int minb4 = getenv("NLK_OPDIV_MINB4") != nullptr;      // A/B: cap the registers so that four blocks fit an SM
#define FN3(N_) case N_: { static bool s_ = false; if (!s_) { set_smem(k_opdiv3_t<N_, 1>, sm3); set_smem(k_opdiv3_t<N_, 4>, sm3); set_smem(k_opdiv3_t<N_, 0>, sm3); s_ = true; } \
      if (minb4) k_opdiv3_t<N_, 4><<<(unsigned)dm.E, thr3, sm3, st>>>(u, p, dm.rxw2, scale, in_mul, mk, out_mul); \
      else if (minb0) k_opdiv3_t<N_, 0><<<(unsigned)dm.E, thr3, sm3, st>>>(u, p, dm.rxw2, scale, in_mul, mk, out_mul); \
      else k_opdiv3_t<N_, 1><<<(unsigned)dm.E, thr3, sm3, st>>>(u, p, dm.rxw2, scale, in_mul, mk, out_mul); ++g_launches; return true; }
    switch (n) { FN3(4) FN3(5) FN3(6) FN3(7) FN3(8) FN3(9) FN3(10) default: break; }
#undef FN3
  }
  int thr = tp_threads(dm.n, dm.ndim, dm.np1);
#define FN(N_, D_) { static bool s_ = false; if (!s_) { set_smem(k_opdiv_t<N_, D_>, smem); s_ = true; } k_opdiv_t<N_, D_><<<(unsigned)dm.E, thr, smem, st>>>(u, p, dm.rxw2, dm.I12, dm.D12, scale, in_mul, mk, out_mul); }
  TP_SWITCH_N(dm.n * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}
bool tp_opgradt(const DevMesh& dm, const double* p, Ptr3 w, cudaStream_t st) {
  size_t smem = (size_t)(2 * dm.q * dm.n + 7 * dm.np1) * sizeof(double);
  if (smem > 220 * 1024) return false;
  ensure_const_ops(dm, st);
  if (dm.ndim == 3) {
    const int n = dm.n, q = dm.q;
    size_t sm3 = (size_t)(3 * (3 * (q | 1) * q * q + 3 * (n | 1) * q * q + 2 * n * n * q)) * sizeof(double);
    int thr3 = 3 * roundup32(n * n);
    static const int gminb1 = getenv("NLK_OPGRADT_MINB0") == nullptr;     // minBlocks = 1 changes ptxas' schedule: 188.6 -> 182.4 us at 24k elements (A/B: NLK_OPGRADT_MINB0=1)
#define FN3(N_) case N_: { static bool s_ = false; if (!s_) { set_smem(k_opgradt3_t<N_, 0>, sm3); set_smem(k_opgradt3_t<N_, 1>, sm3); s_ = true; } \
      if (gminb1) k_opgradt3_t<N_, 1><<<(unsigned)dm.E, thr3, sm3, st>>>(p, w, dm.rxw2); else k_opgradt3_t<N_, 0><<<(unsigned)dm.E, thr3, sm3, st>>>(p, w, dm.rxw2); ++g_launches; return true; }
    switch (n) { FN3(4) FN3(5) FN3(6) FN3(7) FN3(8) FN3(9) FN3(10) default: break; }
#undef FN3
  }
  int thr = tp_threads(dm.n, dm.ndim, dm.np1);
#define FN(N_, D_) { static bool s_ = false; if (!s_) { set_smem(k_opgradt_t<N_, D_>, smem); s_ = true; } k_opgradt_t<N_, D_><<<(unsigned)dm.E, thr, smem, st>>>(p, w, dm.rxw2, dm.I12t, dm.D12t); }
  TP_SWITCH_N(dm.n * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}
bool tp_schwarz_fdm(const DevMesh& dm, const double* w, double* z, double* t, cudaStream_t st) {
  const size_t npP = dm.ndim == 3 ? (size_t)(dm.n | 1) * dm.n * dm.n : (size_t)dm.np1;
  size_t smem = (size_t)(2 * dm.ndim * dm.n * dm.n + 2 * npP) * sizeof(double);
  int thr = tp_threads(dm.n, dm.ndim, dm.np1);
#define FN(N_, D_) { static bool s_ = false; if (!s_) { set_smem(k_schwarz_fdm_t<N_, D_>, smem); s_ = true; } k_schwarz_fdm_t<N_, D_><<<(unsigned)dm.E, thr, smem, st>>>(w, z, t, dm.fdmS, dm.fdmSt, dm.fdmDinv); }
  TP_SWITCH_N(dm.n * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}

#define TP_CASE_NM(N_, M_) \
  case (N_ * 100 + M_) * 10 + 2: FN(N_, M_, 2); break; \
  case (N_ * 100 + M_) * 10 + 3: FN(N_, M_, 3); break;
#define TP_SWITCH_NM(key) switch (key) { TP_CASE_NM(4, 6) TP_CASE_NM(5, 8) TP_CASE_NM(6, 9) TP_CASE_NM(7, 11) TP_CASE_NM(8, 12) TP_CASE_NM(9, 14) TP_CASE_NM(10, 15) TP_CASE_NM(12, 18) default: return false; }

static inline size_t convect_smem(const DevMesh& dm) {
  const size_t npdP = dm.ndim == 3 ? (size_t)(dm.m | 1) * dm.m * dm.m : (size_t)dm.npd;       // padded fine tile
  return (size_t)(2 * dm.m * dm.n + dm.m * dm.m + 7 * npdP) * sizeof(double);
}
bool tp_convect(const DevMesh& dm, CPtr4 u, int nf, CPtr3 C, Ptr4 out, double alpha, int accumulate, cudaStream_t st) {
  size_t smem = convect_smem(dm);
  if (smem > 220 * 1024) return false;
  ensure_const_ops(dm, st);
  int thr = tp_threads(dm.m, dm.ndim, dm.npd);
  static const int lb = getenv("NLK_CONVECT_LB") ? atoi(getenv("NLK_CONVECT_LB")) : 1;      // 2214 -> 2147 us at 24k elements with (160, 1); 0 = no launch bounds
  if (dm.n == 8 && dm.m == 12 && dm.ndim == 3 && lb > 0 && thr == tp_block_threads(12)) {          // headline shape only: launch-bounds variants
    static bool s_ = false; if (!s_) { set_smem(k_convect_tb<8, 12, 3, 1>, smem); set_smem(k_convect_tb<8, 12, 3, 2>, smem); s_ = true; }
    if (lb == 1) k_convect_tb<8, 12, 3, 1><<<(unsigned)dm.E, thr, smem, st>>>(u, nf, C, out, dm.rxd, dm.I1d, dm.I1dt, dm.Dd, alpha, accumulate);
    else k_convect_tb<8, 12, 3, 2><<<(unsigned)dm.E, thr, smem, st>>>(u, nf, C, out, dm.rxd, dm.I1d, dm.I1dt, dm.Dd, alpha, accumulate);
    ++g_launches; return true;
  }
#define FN(N_, M_, D_) { static bool s_ = false; if (!s_) { set_smem(k_convect_t<N_, M_, D_>, smem); s_ = true; } k_convect_t<N_, M_, D_><<<(unsigned)dm.E, thr, smem, st>>>(u, nf, C, out, dm.rxd, dm.I1d, dm.I1dt, dm.Dd, alpha, accumulate); }
  TP_SWITCH_NM((dm.n * 100 + dm.m) * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}
bool tp_convect_adj(const DevMesh& dm, CPtr3 U, CPtr3 c, Ptr3 out, double alpha, int accumulate, cudaStream_t st, int nj) {
  size_t smem = convect_smem(dm);
  if (smem > 220 * 1024) return false;
  ensure_const_ops(dm, st);
  int thr = tp_threads(dm.m, dm.ndim, dm.npd);
#define FN(N_, M_, D_) { static bool s_ = false; if (!s_) { set_smem(k_convect_adj_t<N_, M_, D_>, smem); s_ = true; } k_convect_adj_t<N_, M_, D_><<<(unsigned)dm.E, thr, smem, st>>>(U, c, out, dm.rxd, dm.I1d, dm.I1dt, dm.Dd, alpha, accumulate, nj); }
  TP_SWITCH_NM((dm.n * 100 + dm.m) * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}

bool tp_swf_a(const DevMesh& dm, const double* r, const double* mul, const int32_t* t1, const double* ghost, double* zint, double* ZF, cudaStream_t st) {
  static const bool no_dmma = getenv("NLK_NO_DMMA") != nullptr;
  if (dm.n == 8 && dm.ndim == 3 && !no_dmma) {            // FP64 tensor-core path, one warp per element
    static const int minb = getenv("NLK_SWF8_MINB") ? atoi(getenv("NLK_SWF8_MINB")) : 1;      // 6 (80 registers, 24 warps per SM) runs the branch ALONE 9 % faster (265 -> 241 us at 24k elements) but leaves the side-stream coarse chain no room: whole preconditioner 1576 -> 1659 us at 99 800 elements
    const unsigned g8 = (unsigned)((dm.E + SWF8_WARPS - 1) / SWF8_WARPS);
    if (minb == 6) k_swf_a8<6><<<g8, 32 * SWF8_WARPS, 0, st>>>(r, mul, t1, ghost, dm.fdmS, dm.fdmSt, dm.fdmDinv, zint, ZF, dm.E);
    else if (minb == 5) k_swf_a8<5><<<g8, 32 * SWF8_WARPS, 0, st>>>(r, mul, t1, ghost, dm.fdmS, dm.fdmSt, dm.fdmDinv, zint, ZF, dm.E);
    else k_swf_a8<1><<<g8, 32 * SWF8_WARPS, 0, st>>>(r, mul, t1, ghost, dm.fdmS, dm.fdmSt, dm.fdmDinv, zint, ZF, dm.E);
    ++g_launches; return true;
  }
  const size_t npP = dm.ndim == 3 ? (size_t)(dm.n | 1) * dm.n * dm.n : (size_t)dm.np1;
  size_t smem = (size_t)(2 * dm.ndim * dm.n * dm.n + 2 * npP) * sizeof(double);
  int thr = tp_threads(dm.n, dm.ndim, dm.np1);
#define FN(N_, D_) { static bool s_ = false; if (!s_) { set_smem(k_swf_a<N_, D_>, smem); s_ = true; } k_swf_a<N_, D_><<<(unsigned)dm.E, thr, smem, st>>>(r, mul, t1, ghost, dm.fdmS, dm.fdmSt, dm.fdmDinv, zint, ZF); }
  TP_SWITCH_N(dm.n * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}
bool tp_swf_b(const DevMesh& dm, const double* zint, const double* ZF, const int32_t* t2, const double* ghost, const double* yc, double* out, cudaStream_t st) {
  const int grid = (int)std::min<size_t>((dm.N2 + 255) / 256, (size_t)148 * 16);
#define FN(N_, D_) k_swf_b<N_, D_><<<grid, 256, 0, st>>>(zint, ZF, t2, ghost, dm.swt, yc, dm.vertex, dm.w2 + dm.q, out, dm.N2);
  TP_SWITCH_N(dm.n * 10 + dm.ndim)
#undef FN
  ++g_launches; return true;
}
void launch_coarse_part_w(const DevMesh& dm, const double* r, const double* mul, double* part, cudaStream_t st) {
  const int wpb = 8; const unsigned grid = (unsigned)((dm.E + wpb - 1) / wpb);
  if (dm.ndim == 3) k_coarse_part_w<3><<<grid, 32 * wpb, 0, st>>>(r, mul, part, dm.w2 + dm.q, dm.q, dm.E);
  else k_coarse_part_w<2><<<grid, 32 * wpb, 0, st>>>(r, mul, part, dm.w2 + dm.q, dm.q, dm.E);
  ++g_launches;
}
void launch_swf_pack(const double* src, const double* mul, const int32_t* idx, int n, double* out, cudaStream_t st) {
  if (n <= 0) return;
  k_swf_pack<<<(n + 255) / 256, 256, 0, st>>>(src, mul, idx, n, out); ++g_launches;
}

}  // namespace nlk
