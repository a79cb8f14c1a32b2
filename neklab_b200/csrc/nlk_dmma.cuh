// FP64 tensor-core (mma.sync.m8n8k4.f64 -> SASS DMMA) building block for the lx1 = 8 element kernels: one warp applies an
// 8 x 8 operator along one direction of an 8 x 8 x 8 element tile held in shared memory ((8 x 8) x (8 x 64) product).
// FP64 has no tcgen05 path on sm_100a; DMMA is its tensor pipe.  Why tensor cores for memory-bound kernels: with scalar FMAs
// every FMA of a contraction needs a shared-memory (or constant-bank) operand, and ncu shows the LSU pipe -- not DRAM, not the
// FP64 pipe -- as the limiter of the element kernels; one DMMA consumes one 8-byte LDS per lane for 256 FMAs.
#pragma once

namespace nlk {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// out = (M along direction DIR) in, both tiles [k][j][i] with x-pitch PN; M given as A fragments (a0: K-block 0, a1: K-block 1)
template <int DIR, int PN, bool ACC = false>
__device__ __forceinline__ void dmma_contract8(double* __restrict__ out, const double* __restrict__ in, double a0, double a1, int lane) {
  const int kr = lane & 3, cq = lane >> 2;
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    // column c = cb*8 + cq enumerates the two uncontracted indices; contracted index runs with stride `cs`
    int base_b, cs;
    { const int c = cb * 8 + cq;
      if (DIR == 0) { base_b = PN * c; cs = 1; }
      else if (DIR == 1) { base_b = (c & 7) + PN * 8 * (c >> 3); cs = PN; }
      else { base_b = (c & 7) + PN * (c >> 3); cs = PN * 8; } }
    // D fragment: row o = lane>>2, columns cb*8 + 2*(lane&3) + {0,1}
    const int o = cq;
    int bd[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = cb * 8 + 2 * kr + h;
      if (DIR == 0) bd[h] = PN * c + o;
      else if (DIR == 1) bd[h] = (c & 7) + PN * 8 * (c >> 3) + PN * o;
      else bd[h] = (c & 7) + PN * (c >> 3) + PN * 8 * o;
    }
    double d0 = ACC ? out[bd[0]] : 0.0, d1 = ACC ? out[bd[1]] : 0.0;
    dmma884(d0, d1, a0, in[base_b + kr * cs]);
    dmma884(d0, d1, a1, in[base_b + (4 + kr) * cs]);
    out[bd[0]] = d0; out[bd[1]] = d1;
  }
  __syncwarp();
}

}  // namespace nlk
