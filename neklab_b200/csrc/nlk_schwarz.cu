// Fused additive-Schwarz branch of the pressure preconditioner (Nek hsmg_schwarz / hsmg_fdm / hsmg_extrude analogue, K10):
// host-side pull tables, multi-rank ghost exchange, and the orchestration of the preconditioner application
//   z = W sum_e R_e^T FDM_e R_e r  +  R_0^T A_0^-1 R_0 r .
// Kernels: k_swf_a / k_swf_b / k_coarse_part_w / k_swf_pack in nlk_kernels_tp.cu.  The unfused chain (embed -> dssum -> fdm
// -> dssum -> gather, nlk_solver.cu) stays as the set-up path of the overlap-count weights and as the NLK_NO_SWF=1 A/B switch;
// both give the same result to round-off (tests/test_gpu_kernels.py::test_precond_parity, test_fused_schwarz_equals_chain).
#include "nlk_ctx.hpp"
#include <algorithm>
#include <cstdlib>

namespace nlk {

struct SwfState {
  int32_t *t1 = nullptr, *t2 = nullptr;          // [E][2d][fs]
  double *zint = nullptr, *ZF = nullptr;         // [N2], [E*2d*fs]
  // multi-rank: per neighbour a contiguous range of ghost slots (same order on both sides: ascending global node id)
  int nghost = 0; std::vector<int> nb_off;       // size nneigh + 1
  int32_t *send_inner = nullptr, *send_face = nullptr;
  double *sendbuf = nullptr, *ghostA = nullptr, *ghostB = nullptr;
};

static const int NCCL_DOUBLE_ = 8;

// (e, i, j, k) of local node g and its face classification
struct NodeCls { int64_t e; int i, j, k, nb, f, s; int32_t inner; };
static NodeCls classify(const HostMesh& hm, int32_t g) {
  const int n = hm.n, d = hm.ndim, q = hm.q;
  NodeCls c{}; c.e = g / hm.np1; const int p = (int)(g - c.e * hm.np1);
  c.i = p % n; c.j = (p / n) % n; c.k = d == 3 ? p / (n * n) : 0;
  const bool bi = c.i == 0 || c.i == n - 1, bj = c.j == 0 || c.j == n - 1, bk = d == 3 && (c.k == 0 || c.k == n - 1);
  c.nb = (int)bi + (int)bj + (int)bk; c.f = -1; c.s = -1; c.inner = -1;
  if (c.nb != 1) return c;
  const int k0 = d == 3 ? 1 : 0;
  int ii = c.i, jj = c.j, kk = c.k;
  if (bi) { c.f = c.i == 0 ? 0 : 1; c.s = (c.j - 1) + q * (c.k - k0); ii = c.i == 0 ? 1 : n - 2; }
  else if (bj) { c.f = 2 + (c.j == 0 ? 0 : 1); c.s = (c.i - 1) + q * (c.k - k0); jj = c.j == 0 ? 1 : n - 2; }
  else { c.f = 4 + (c.k == 0 ? 0 : 1); c.s = (c.i - 1) + q * (c.j - 1); kk = c.k == 0 ? 1 : n - 2; }
  c.inner = (int32_t)(c.e * hm.np2 + (int64_t)((d == 3 ? (kk - 1) * q : 0) + (jj - 1)) * q + (ii - 1));
  return c;
}

void swf_release(nlk_ctx* c) { delete static_cast<SwfState*>(c->swf); c->swf = nullptr; }

int swf_setup(nlk_ctx* c) {
  if (getenv("NLK_NO_SWF")) return 0;
  const HostMesh& hm = c->mesh->hm; const DevMesh& dm = c->dm;
  const int d = hm.ndim, q = hm.q, nf = 2 * d, fs = d == 3 ? q * q : q;
  if (q < 2) return 0;
  if ((size_t)hm.E * nf * fs > (size_t)INT32_MAX || dm.N2 > (size_t)INT32_MAX) return 0;      // int32 tables
  auto* S = new SwfState(); c->swf = S;
  const size_t nt = (size_t)hm.E * nf * fs;
  std::vector<int32_t> t1(nt, -1), t2(nt, -1);
  // local pairs: groups of the gather-scatter CSR with exactly two copies, both face-interior
  const size_t ng = hm.gs_off.size() - 1;
  for (size_t g = 0; g < ng; ++g) {
    if (hm.gs_off[g + 1] - hm.gs_off[g] != 2) continue;
    const int32_t a = hm.gs_idx[hm.gs_off[g]], b = hm.gs_idx[hm.gs_off[g] + 1];
    const NodeCls ca = classify(hm, a), cb = classify(hm, b);
    if (ca.nb != 1 || cb.nb != 1) continue;
    const size_t sa = ((size_t)ca.e * nf + ca.f) * fs + ca.s, sb = ((size_t)cb.e * nf + cb.f) * fs + cb.s;
    t1[sa] = cb.inner; t2[sa] = (int32_t)sb;
    t1[sb] = ca.inner; t2[sb] = (int32_t)sa;
  }
  // remote pairs: interface nodes that are face-interior here (then they are on the other side too)
  std::vector<int32_t> send_inner, send_face;
  S->nb_off.assign(1, 0);
  for (const Neighbor& nb : hm.neigh) {
    for (size_t k = 0; k < nb.gids.size(); ++k) {
      const NodeCls cl = classify(hm, nb.rep[k]);
      if (cl.nb != 1) continue;
      const size_t sl = ((size_t)cl.e * nf + cl.f) * fs + cl.s;
      if (t1[sl] != -1) continue;                          // already paired locally (cannot happen on a conforming mesh)
      const int gh = (int)send_inner.size();
      t1[sl] = -2 - gh; t2[sl] = -2 - gh;
      send_inner.push_back(cl.inner); send_face.push_back((int32_t)sl);
    }
    S->nb_off.push_back((int)send_inner.size());
  }
  S->nghost = (int)send_inner.size();
  if (dev_upload(c, &S->t1, t1) || dev_upload(c, &S->t2, t2)) return 1;
  if (dev_alloc(c, &S->zint, dm.N2) || dev_alloc(c, &S->ZF, nt)) return 1;
  if (S->nghost > 0) {
    if (dev_upload(c, &S->send_inner, send_inner) || dev_upload(c, &S->send_face, send_face)) return 1;
    if (dev_alloc(c, &S->sendbuf, (size_t)S->nghost) || dev_alloc(c, &S->ghostA, (size_t)S->nghost) || dev_alloc(c, &S->ghostB, (size_t)S->nghost)) return 1;
  }
  return 0;
}

static int swf_exchange(nlk_ctx* c, SwfState& S, double* ghost) {
  if (c->nccl.nranks <= 1 || S.nghost == 0) return 0;
  int r = c->nccl.GroupStart();
  for (size_t i = 0; i < c->neigh.size() && !r; ++i) {
    const int off = S.nb_off[i], cnt = S.nb_off[i + 1] - off;
    if (cnt == 0) continue;
    r = c->nccl.Send(S.sendbuf + off, (size_t)cnt, NCCL_DOUBLE_, c->neigh[i].rank, c->nccl.comm, c->st);
    if (!r) r = c->nccl.Recv(ghost + off, (size_t)cnt, NCCL_DOUBLE_, c->neigh[i].rank, c->nccl.comm, c->st);
  }
  int r2 = c->nccl.GroupEnd();
  if (r || r2) { set_error(std::string("NCCL error in the Schwarz overlap exchange: ") + c->nccl.GetErrorString(r ? r : r2)); return 1; }
  return 0;
}

// the Schwarz branch alone: z = W sum_e R_e^T FDM_e R_e (in_mul r) [+ prolong(yc) when yc != nullptr]
int swf_apply(nlk_ctx* c, const double* r, const double* in_mul, const double* yc, double* z, cudaEvent_t wait_before_b, int phase) {
  SwfState& S = *static_cast<SwfState*>(c->swf);
  const DevMesh& dm = c->dm;
  if (phase != 2) {
    if (S.nghost > 0) { launch_swf_pack(r, in_mul, S.send_inner, S.nghost, S.sendbuf, c->st); if (swf_exchange(c, S, S.ghostA)) return 1; }
    if (!tp_swf_a(dm, r, in_mul, S.t1, S.ghostA, S.zint, S.ZF, c->st)) { set_error("fused Schwarz: no kernel instantiation for this lx1"); return 1; }
    if (S.nghost > 0) { launch_swf_pack(S.ZF, nullptr, S.send_face, S.nghost, S.sendbuf, c->st); if (swf_exchange(c, S, S.ghostB)) return 1; }
  }
  if (phase != 1) {
    if (wait_before_b) NLK_CUDA(cudaStreamWaitEvent(c->st, wait_before_b, 0));
    tp_swf_b(dm, S.zint, S.ZF, S.t2, S.ghostB, yc, z, c->st);
  }
  return 0;
}

}  // namespace nlk
