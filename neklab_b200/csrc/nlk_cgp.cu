// Persistent cooperative Jacobi-PCG for the Helmholtz solves (K8, Nek `hmholtz` + `cggo`), all velocity components at once.
// One launch per solve: the phases of an iteration (direction update + axhelm | dssum | x/r update + reductions) are
// separated by grid-wide barriers instead of kernel boundaries, the CG scalars are recomputed redundantly (and
// identically: fixed summation order) by every CTA from per-CTA partial sums, and the convergence test of cggo
// (rbn2 = sqrt(sum r^2 mult binv / vol) <= tol) is evaluated on the device.  Used on one GPU when the problem is small
// enough to be launch/latency-bound (the reference's 2-D configs); large or multi-rank problems use the streamed kernels.
#include "nlk_device.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace nlk {

template <int N, int DIM>
__global__ void k_cg_persistent(CgArgs a) {
  constexpr int NZ = DIM == 3 ? N : 1, NN = N * N, NP = NN * NZ, NG = DIM == 3 ? 6 : 3;
  constexpr int EPB = NN >= 100 ? 1 : (NN >= 64 ? 2 : 4);
  constexpr int NT = (NN * EPB + 31) / 32 * 32;
  cg::grid_group grid = cg::this_grid();
  __shared__ double sD[NN], sDt[NN];
  __shared__ double s_u[EPB][NN], s_gr[EPB][NN], s_gs[EPB][NN];
  __shared__ double s_red[3][3];
  __shared__ double s_blk[9][NT / 32];
  const int lin = threadIdx.x, le = lin / NN, tid = lin - le * NN;
  const bool tile_thread = le < EPB;
  const int i = tid % N, j = tid / N;
  const int lane = lin & 31, wid = lin >> 5;
  constexpr int NW = NT / 32;
  const size_t gtid = (size_t)blockIdx.x * NT + lin, gsize = (size_t)gridDim.x * NT;
  const size_t N1 = (size_t)a.E * NP;
  const int nf = a.nf, nb = gridDim.x;
  for (int idx = lin; idx < NN; idx += NT) { double v = a.D[idx]; sD[idx] = v; sDt[(idx % N) * N + idx / N] = v; }

  // block-level sum of up to 9 per-thread values -> partial[q * nb + blockIdx.x]
  auto block_partials = [&](const double* v, int nq, int qbase) {
    for (int q = 0; q < nq; ++q) {
      double t = v[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if (lane == 0) s_blk[q][wid] = t;
    }
    __syncthreads();
    if (lin < nq) { double t = 0; for (int w = 0; w < NW; ++w) t += s_blk[lin][w]; a.partial[(size_t)(qbase + lin) * nb + blockIdx.x] = t; }
  };
  // every CTA sums the per-CTA partials of quantity (f, kind) in the same fixed order
  auto gather_partials = [&](int kind_lo, int kind_hi) {
    int q = 0;
    for (int f = 0; f < nf; ++f) for (int kind = kind_lo; kind <= kind_hi; ++kind, ++q) {
      if (wid == q % NW) {
        double t = 0;
        for (int b = lane; b < nb; b += 32) t += __ldcg(&a.partial[(size_t)(f * 3 + kind) * nb + b]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) s_red[f][kind] = t;
      }
    }
    __syncthreads();
  };

  // ---- init: rhs (in r) -> dssum -> mask ; x = p = 0 ; first (r.z, r.r) sums
  for (size_t g = gtid; g < (size_t)a.ngs; g += gsize) {
    const int b = a.gs_off[g], e = a.gs_off[g + 1];
    for (int f = 0; f < nf; ++f) { double* r = a.f[f].r; double s = 0; for (int t = b; t < e; ++t) s += r[a.gs_idx[t]]; for (int t = b; t < e; ++t) r[a.gs_idx[t]] = s; }
  }
  grid.sync();
  {
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (size_t p = gtid; p < N1; p += gsize) {
      const double dg = a.h1 * a.diagA[p] + a.h2 * a.diagB[p], m = a.mult[p], mb = m * a.binv[p];
      for (int f = 0; f < nf; ++f) {
        double ri = a.f[f].r[p] * a.f[f].mask[p];
        a.f[f].r[p] = ri; a.f[f].x[p] = 0.0; a.f[f].p[p] = 0.0;
        v[2 * f] += ri * (ri / dg) * m; v[2 * f + 1] += ri * ri * mb;
      }
    }
    double vv[9];
    for (int f = 0; f < 3; ++f) { vv[3 * f] = 0; vv[3 * f + 1] = v[2 * f]; vv[3 * f + 2] = v[2 * f + 1]; }
    block_partials(vv, 3 * nf, 0);
  }
  grid.sync();
  int iter[3] = {0, 0, 0};
  bool done[3] = {false, false, false};
  double rtz1[3] = {1, 1, 1}, beta[3] = {0, 0, 0};
  for (int f = nf; f < 3; ++f) done[f] = true;
  while (true) {
    gather_partials(1, 2);
    bool all_done = true;
    for (int f = 0; f < nf; ++f) if (!done[f]) {
      const double rtz2 = rtz1[f];
      rtz1[f] = s_red[f][1];
      const double rbn2 = sqrt(s_red[f][2] / a.vol);
      if (rbn2 <= a.tol || iter[f] >= a.maxit || !(rtz1[f] == rtz1[f])) done[f] = true;
      else { beta[f] = iter[f] == 0 ? 0.0 : rtz1[f] / rtz2; iter[f] += 1; all_done = false; }
    }
    __syncthreads();          // s_red is rewritten below
    if (all_done) break;
    // ---- P1: p = z + beta p ; w = (h1 A + h2 B) p element-locally ; p.w over unshared nodes
    double pap[3] = {0, 0, 0};
    for (int64_t e0 = (int64_t)blockIdx.x * EPB; e0 < a.E; e0 += (int64_t)gridDim.x * EPB) {
      const int64_t e = e0 + le;
      const bool active = tile_thread && e < a.E;
      const size_t eb = (size_t)(active ? e : 0) * NP;
      for (int f = 0; f < nf; ++f) {
        if (done[f]) continue;              // uniform across the grid
        double ru[NZ], rw[NZ];
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
          const size_t g = eb + k * NN + tid;
          double pv = 0.0;
          if (active) { pv = a.f[f].r[g] / (a.h1 * a.diagA[g] + a.h2 * a.diagB[g]) + beta[f] * a.f[f].p[g]; a.f[f].p[g] = pv; }
          ru[k] = pv; rw[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NZ; ++k) {
          if (tile_thread) s_u[le][tid] = ru[k];
          __syncthreads();
          double gr = 0, gs = 0, gt = 0;
          if (tile_thread) {
            double ur = 0, us = 0, ut = 0;
#pragma unroll
            for (int l = 0; l < N; ++l) { ur += sDt[l * N + i] * s_u[le][j * N + l]; us += sD[j * N + l] * s_u[le][l * N + i]; }
            if (DIM == 3) {
#pragma unroll
              for (int l = 0; l < N; ++l) ut += sD[k * N + l] * ru[l];
            }
            const double* Gp = a.G + (size_t)(active ? e : 0) * NG * NP + k * NN + tid;
            if (DIM == 3) {
              const double g11 = Gp[0], g22 = Gp[NP], g33 = Gp[2 * NP], g12 = Gp[3 * NP], g13 = Gp[4 * NP], g23 = Gp[5 * NP];
              gr = g11 * ur + g12 * us + g13 * ut; gs = g12 * ur + g22 * us + g23 * ut; gt = g13 * ur + g23 * us + g33 * ut;
            } else {
              const double g11 = Gp[0], g22 = Gp[NP], g12 = Gp[2 * NP];
              gr = g11 * ur + g12 * us; gs = g12 * ur + g22 * us;
            }
            s_gr[le][tid] = gr; s_gs[le][tid] = gs;
          }
          __syncthreads();
          if (tile_thread) {
            double acc = 0;
#pragma unroll
            for (int l = 0; l < N; ++l) acc += sD[l * N + i] * s_gr[le][j * N + l] + sD[l * N + j] * s_gs[le][l * N + i];
            rw[k] += acc;
            if (DIM == 3) {
#pragma unroll
              for (int l = 0; l < N; ++l) rw[l] += sD[k * N + l] * gt;
            }
          }
          __syncthreads();
        }
        if (active) {
#pragma unroll
          for (int k = 0; k < NZ; ++k) {
            const size_t g = eb + k * NN + tid;
            const double wv = a.h1 * rw[k] + a.h2 * a.bm1[g] * ru[k];
            a.f[f].w[g] = wv;
            if (a.mult[g] == 1.0) pap[f] += wv * a.f[f].mask[g] * ru[k];
          }
        }
      }
    }
    grid.sync();
    // ---- P2: dssum of w over the shared-node groups ; their p.w contribution
    for (size_t g = gtid; g < (size_t)a.ngs; g += gsize) {
      const int b = a.gs_off[g], e = a.gs_off[g + 1];
      const int i0 = a.gs_idx[b];
      for (int f = 0; f < nf; ++f) {
        if (done[f]) continue;
        double* w = a.f[f].w; double s = 0;
        for (int t = b; t < e; ++t) s += w[a.gs_idx[t]];
        for (int t = b; t < e; ++t) w[a.gs_idx[t]] = s;
        pap[f] += s * a.f[f].mask[i0] * a.f[f].p[i0];
      }
    }
    { double vv[9]; for (int f = 0; f < 3; ++f) { vv[3 * f] = pap[f]; vv[3 * f + 1] = 0; vv[3 * f + 2] = 0; }
      // only kind 0 is written in this phase: write per-field pap into its slot
      for (int f = 0; f < nf; ++f) {
        double t = vv[3 * f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) s_blk[f][wid] = t;
      }
      __syncthreads();
      if (lin < nf) { double t = 0; for (int w = 0; w < NW; ++w) t += s_blk[lin][w]; a.partial[(size_t)(lin * 3 + 0) * nb + blockIdx.x] = t; }
    }
    grid.sync();
    // ---- P3: alpha ; x += alpha p ; r -= alpha mask w ; new (r.z, r.r) sums
    gather_partials(0, 0);
    double alpha[3] = {0, 0, 0};
    for (int f = 0; f < nf; ++f) if (!done[f]) alpha[f] = rtz1[f] / s_red[f][0];
    __syncthreads();
    {
      double v[6] = {0, 0, 0, 0, 0, 0};
      for (size_t p = gtid; p < N1; p += gsize) {
        const double dg = a.h1 * a.diagA[p] + a.h2 * a.diagB[p], m = a.mult[p], mb = m * a.binv[p];
        for (int f = 0; f < nf; ++f) {
          if (done[f]) continue;
          a.f[f].x[p] += alpha[f] * a.f[f].p[p];
          const double ri = a.f[f].r[p] - alpha[f] * a.f[f].mask[p] * a.f[f].w[p];
          a.f[f].r[p] = ri;
          v[2 * f] += ri * (ri / dg) * m; v[2 * f + 1] += ri * ri * mb;
        }
      }
      for (int f = 0; f < nf; ++f) {
        if (done[f]) continue;
        for (int kind = 0; kind < 2; ++kind) {
          double t = v[2 * f + kind];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
          if (lane == 0) s_blk[2 * f + kind][wid] = t;
        }
      }
      __syncthreads();
      if (lin < 2 * nf && !done[lin / 2]) { double t = 0; for (int w = 0; w < NW; ++w) t += s_blk[lin][w]; a.partial[(size_t)((lin / 2) * 3 + 1 + (lin & 1)) * nb + blockIdx.x] = t; }
    }
    grid.sync();
  }
  // ---- exit: sol += x ; iteration counts
  for (int f = 0; f < nf; ++f) if (a.f[f].sol) for (size_t p = gtid; p < N1; p += gsize) a.f[f].sol[p] += a.f[f].x[p];
  if (blockIdx.x == 0 && lin < nf) { a.iters_out[lin] = iter[lin]; atomicAdd(a.iters_total, (unsigned long long)iter[lin]); }
}

template <int N, int DIM>
static bool cgp_dispatch(const DevMesh& dm, CgArgs& a, cudaStream_t st) {
  constexpr int NN = N * N, EPB = NN >= 100 ? 1 : (NN >= 64 ? 2 : 4), NT = (NN * EPB + 31) / 32 * 32;
  static int blocks_per_sm = -1, nsm = 0;
  if (blocks_per_sm < 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_cg_persistent<N, DIM>, NT, 0);
    int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) blocks_per_sm = 0;
  }
  if (blocks_per_sm <= 0) return false;
  int want = (int)((dm.E + EPB - 1) / EPB);
  // small (latency-bound) problems: 2 CTAs/SM keep the grid barrier cheap; HBM-bound sizes need the bandwidth of 8 CTAs/SM
  static int cap_env = -2; if (cap_env == -2) { const char* e = getenv("NLK_CGP_BLOCKS_PER_SM"); cap_env = e ? atoi(e) : -1; }
  const int cap = cap_env > 0 ? cap_env : ((size_t)dm.N1 > 1500000 ? 8 : 2);
  int grid = std::min(want, nsm * std::min(blocks_per_sm, cap));
  if (grid < 1) grid = 1;
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_cg_persistent<N, DIM>, dim3(grid), dim3(NT), args, 0, st);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  ++g_launches;
  return true;
}

bool launch_cg_persistent(const DevMesh& dm, CgArgs& a, cudaStream_t st) {
  bool ok = false;
#define CGP(N_, D_) ok = cgp_dispatch<N_, D_>(dm, a, st)
  switch (dm.n * 10 + dm.ndim) {
    case 42: CGP(4, 2); break;  case 43: CGP(4, 3); break;
    case 52: CGP(5, 2); break;  case 53: CGP(5, 3); break;
    case 62: CGP(6, 2); break;  case 63: CGP(6, 3); break;
    case 72: CGP(7, 2); break;  case 73: CGP(7, 3); break;
    case 82: CGP(8, 2); break;  case 83: CGP(8, 3); break;
    case 92: CGP(9, 2); break;  case 93: CGP(9, 3); break;
    case 102: CGP(10, 2); break; case 103: CGP(10, 3); break;
    case 112: CGP(11, 2); break; case 122: CGP(12, 2); break;
    default: ok = false;
  }
#undef CGP
  return ok;
}

}  // namespace nlk
