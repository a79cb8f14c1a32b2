// C-ABI: context life-cycle (uploads + device-side setup), nek_dvector operations, exptA operator, test hooks.
// Reference interfaces replaced: see include/nlk.h (file:line citations per entry point).
#include "nlk_ctx.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

using namespace nlk;

namespace nlk {
static int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { set_error("no usable CUDA device: libnlk has no CPU fallback"); return 1; }
  return 0;
}

// upload everything the kernels need and run the device-side parts of the setup (dssum-dependent quantities)
static int ctx_setup(nlk_ctx* c) {
  const HostMesh& hm = c->mesh->hm; DevMesh& dm = c->dm; const Basis& b = hm.b;
  const int d = hm.ndim, n = hm.n, q = hm.q, m = hm.m;
  dm.ndim = d; dm.n = n; dm.m = m; dm.q = q; dm.np1 = hm.np1; dm.np2 = hm.np2; dm.npd = hm.npd; dm.ng = d == 3 ? 6 : 3;
  dm.E = hm.E; dm.N1 = (size_t)hm.E * hm.np1; dm.N2 = (size_t)hm.E * hm.np2; dm.Nd = (size_t)hm.E * hm.npd;
  dm.has_outflow = hm.has_outflow; dm.nvert = hm.nvert;
  const size_t N1 = dm.N1, N2 = dm.N2;
  // 1-D operators
  if (dev_upload(c, &dm.D, b.D) || dev_upload(c, &dm.Dt, mat_transpose(b.D, n, n)) || dev_upload(c, &dm.I12, b.I12) ||
      dev_upload(c, &dm.I12t, mat_transpose(b.I12, q, n)) || dev_upload(c, &dm.D12, b.D12) || dev_upload(c, &dm.D12t, mat_transpose(b.D12, q, n)) ||
      dev_upload(c, &dm.I1d, b.I1d) || dev_upload(c, &dm.I1dt, mat_transpose(b.I1d, m, n)) || dev_upload(c, &dm.Dd, b.Dd) ||
      dev_upload(c, &dm.Ddt, mat_transpose(b.Dd, m, m))) return 1;
  { std::vector<double> wz(b.w2); wz.insert(wz.end(), b.z2.begin(), b.z2.end()); if (dev_upload(c, &dm.w2, wz)) return 1; }
  // geometry (packed [E][k][np])
  {
    std::vector<double> G((size_t)hm.E * dm.ng * hm.np1);
    const int order3[6] = {0, 1, 2, 3, 4, 5}, order2[3] = {0, 1, 3};
    for (int64_t e = 0; e < hm.E; ++e) for (int k = 0; k < dm.ng; ++k) {
      const std::vector<double>& src = hm.G[d == 3 ? order3[k] : order2[k]];
      std::copy(src.begin() + e * hm.np1, src.begin() + (e + 1) * hm.np1, G.begin() + ((size_t)e * dm.ng + k) * hm.np1);
    }
    if (dev_upload(c, &dm.G, G)) return 1;
    std::vector<double> R((size_t)hm.E * d * d * hm.np2), Rd((size_t)hm.E * d * d * hm.npd), Rj((size_t)hm.E * d * d * hm.np1);
    for (int64_t e = 0; e < hm.E; ++e) for (int k = 0; k < d * d; ++k) {
      std::copy(hm.rxw2[k].begin() + e * hm.np2, hm.rxw2[k].begin() + (e + 1) * hm.np2, R.begin() + ((size_t)e * d * d + k) * hm.np2);
      std::copy(hm.rxd[k].begin() + e * hm.npd, hm.rxd[k].begin() + (e + 1) * hm.npd, Rd.begin() + ((size_t)e * d * d + k) * hm.npd);
      for (int p = 0; p < hm.np1; ++p) Rj[((size_t)e * d * d + k) * hm.np1 + p] = hm.rx[k][e * hm.np1 + p] / hm.jac[e * hm.np1 + p];
    }
    if (dev_upload(c, &dm.rxw2, R) || dev_upload(c, &dm.rxd, Rd) || dev_upload(c, &dm.rxj, Rj)) return 1;
  }
  if (dev_upload(c, &dm.bm1, hm.bm1) || dev_upload(c, &dm.bm2, hm.bm2)) return 1;
  for (int k = 0; k < d; ++k) if (dev_upload(c, &dm.mask[k], hm.vmask[k])) return 1;
  if (dev_upload(c, &dm.mask[3], hm.tmask)) return 1;
  {
    std::vector<double> ml(N2), mu(N2), bi(N2), ones(N2, 1.0);
    for (size_t i = 0; i < N2; ++i) { ml[i] = std::sqrt(1.0 / hm.bm2[i]); mu[i] = std::sqrt(hm.bm2[i]); bi[i] = 1.0 / hm.bm2[i]; }
    if (dev_upload(c, &dm.ml, ml) || dev_upload(c, &dm.mu, mu) || dev_upload(c, &c->pw[5], bi) || dev_upload(c, &c->ones2, ones)) return 1;
    std::vector<double> dri(n);
    dri[0] = 1.0 / (b.z1[1] - b.z1[0]); dri[n - 1] = 1.0 / (b.z1[n - 1] - b.z1[n - 2]);
    for (int i = 1; i < n - 1; ++i) dri[i] = 2.0 / (b.z1[i + 1] - b.z1[i - 1]);
    if (dev_upload(c, &dm.dri, dri)) return 1;
  }
  if (dev_upload(c, &dm.gs_off, hm.gs_off) || dev_upload(c, &dm.gs_idx, hm.gs_idx)) return 1;
  dm.ngs = (int)hm.gs_off.size() - 1;
  {
    std::vector<int32_t> g2, g4, ro(1, 0), ri;
    for (int g = 0; g < dm.ngs; ++g) {
      const int b_ = hm.gs_off[g], n_ = hm.gs_off[g + 1] - b_;
      if (n_ == 2) { g2.push_back(hm.gs_idx[b_]); g2.push_back(hm.gs_idx[b_ + 1]); }
      else if (n_ == 4) { for (int t = 0; t < 4; ++t) g4.push_back(hm.gs_idx[b_ + t]); }
      else { for (int t = 0; t < n_; ++t) ri.push_back(hm.gs_idx[b_ + t]); ro.push_back((int32_t)ri.size()); }
    }
    dm.ngs2 = (int)g2.size() / 2; dm.ngs4 = (int)g4.size() / 4; dm.ngsr = (int)ro.size() - 1;
    if (dev_upload(c, &dm.gs2, g2) || dev_upload(c, &dm.gs4, g4) || dev_upload(c, &dm.gsr_off, ro) || dev_upload(c, &dm.gsr_idx, ri)) return 1;
  }
  if (dev_upload(c, &dm.vertex, hm.vertex_local) || dev_upload(c, &dm.vert_off, hm.vert_off) || dev_upload(c, &dm.vert_ec, hm.vert_ec)) return 1;
  for (int k = 0; k < d; ++k) if (dev_upload(c, &c->xyz[k], hm.xyz[k])) return 1;
  if (dev_upload(c, &c->d_lglel, hm.lglel)) return 1;
  // neighbours (multi-rank)
  for (const Neighbor& nb : hm.neigh) {
    DevNeighbor dn{}; dn.rank = nb.rank; dn.cnt = (int)nb.gids.size();
    // local copies of each shared node: look the gid up in the interface CSR
    std::vector<int32_t> off(1, 0), idx;
    for (size_t i = 0; i < nb.gids.size(); ++i) {
      auto it = std::lower_bound(hm.if_gids.begin(), hm.if_gids.end(), nb.gids[i]);
      size_t g = it - hm.if_gids.begin();
      for (int32_t u = hm.if_off[g]; u < hm.if_off[g + 1]; ++u) idx.push_back(hm.if_idx[u]);
      off.push_back((int32_t)idx.size());
    }
    if (dev_upload(c, &dn.rep, nb.rep) || dev_upload(c, &dn.cp_off, off) || dev_upload(c, &dn.cp_idx, idx)) return 1;
    if (dev_alloc(c, &dn.sendbuf, (size_t)3 * dn.cnt) || dev_alloc(c, &dn.recvbuf, (size_t)3 * dn.cnt)) return 1;
    c->neigh.push_back(dn);
  }
  // scalars / reducer
  if (dev_alloc(c, &c->d_sc, 1) || dev_alloc(c, &c->d_red, 512)) return 1;
  NLK_CUDA(cudaMallocHost((void**)&c->h_sc, sizeof(SolverScal)));
  NLK_CUDA(cudaMallocHost((void**)&c->h_red, sizeof(double) * 512));
  NLK_CUDA(cudaMallocHost((void**)&c->h_seq, sizeof(unsigned int) * 16)); c->h_seq[0] = 0; c->seq = 0;
  c->red.maxblocks = 2048;
  if (dev_alloc(c, &c->red.partial, (size_t)16 * c->red.maxblocks) || dev_alloc(c, &c->red.counter, 1)) return 1;
  // state + work arrays
  for (int k = 0; k < 3; ++k) {
    if (k < d) {
      if (dev_alloc(c, &c->U[k], N1) || dev_alloc(c, &c->vp[k], N1) || dev_alloc(c, &c->vlag[0][k], N1) || dev_alloc(c, &c->vlag[1][k], N1) ||
          dev_alloc(c, &c->exx1[k], N1) || dev_alloc(c, &c->exx2[k], N1) || dev_alloc(c, &c->bf[k], N1)) return 1;
    }
  }
  if (dev_alloc(c, &c->T, N1) || dev_alloc(c, &c->tp, N1) || dev_alloc(c, &c->prp, N2) || dev_alloc(c, &c->prlag, N2)) return 1;
  if (c->prm.ifheat) { if (dev_alloc(c, &c->tlag[0], N1) || dev_alloc(c, &c->tlag[1], N1) || dev_alloc(c, &c->vgradt1, N1) || dev_alloc(c, &c->vgradt2, N1) || dev_alloc(c, &c->bq, N1)) return 1; }
  for (int k = 0; k < 8; ++k) if (dev_alloc(c, &c->wk[k], N1)) return 1;
  if (dev_alloc(c, &c->cg_x, N1) || dev_alloc(c, &c->cg_r, N1) || dev_alloc(c, &c->cg_p, N1) || dev_alloc(c, &c->cg_w, N1)) return 1;
  if (dev_alloc(c, &c->cg_pap_partial, (size_t)hm.E + 1) || dev_alloc(c, &c->cg_pap_counter, 4)) return 1;
  for (int k = 0; k < 5; ++k) if (dev_alloc(c, &c->pw[k], N2)) return 1;
  // persistent cooperative PCG: single rank and small enough to be launch/latency-bound (NLK_NO_CGP=1 disables it)
  { size_t lim = 1500000; if (const char* e = getenv("NLK_CGP_MAX_POINTS")) lim = (size_t)atoll(e); c->use_cgp = hm.nranks <= 1 && N1 <= lim && !getenv("NLK_NO_CGP"); }
  if (c->use_cgp) {
    for (int k = 0; k < d; ++k) if (dev_alloc(c, &c->cgm_x[k], N1) || dev_alloc(c, &c->cgm_p[k], N1) || dev_alloc(c, &c->cgm_w[k], N1)) return 1;
    if (dev_alloc(c, &c->d_cg_iters, 4) || dev_alloc(c, &c->d_cg_total, 1)) return 1;
  }
  if (dev_alloc(c, &c->gm_V, (size_t)(c->prm.lgmres + 1) * N2) || dev_alloc(c, &c->gm_Z, (size_t)c->prm.lgmres * N2)) return 1;
  if (dev_alloc(c, &c->sw_w, N1) || dev_alloc(c, &c->sw_z, N1) || dev_alloc(c, &c->sw_t, N1)) return 1;
  if (c->prm.pr_proj > 0) { if (dev_alloc(c, &c->proj_X, (size_t)c->prm.pr_proj * N2) || dev_alloc(c, &c->proj_EX, (size_t)c->prm.pr_proj * N2) || dev_alloc(c, &c->proj_w, N2) || dev_alloc(c, &c->proj_xbar, N2)) return 1; }
  // ---- device-side setup: global volumes, binvm1, vmult, Jacobi diagonals
  {
    double vols[3] = {hm.volvm1, hm.volvm2, (double)N2};
    NLK_CUDA(cudaMemcpyAsync(c->d_red, vols, sizeof(vols), cudaMemcpyHostToDevice, c->st));
    if (ctx_allreduce(c, c->d_red, 3, false)) return 1;
    if (ctx_read_scalars(c, 3)) return 1;
    dm.volvm1 = c->h_red[0]; dm.volvm2 = c->h_red[1]; dm.N2_global = (int64_t)std::llround(c->h_red[2]);
  }
  if (dev_alloc(c, &dm.binvm1, N1) || dev_alloc(c, &dm.vmult, N1) || dev_alloc(c, &dm.diagA, N1) || dev_alloc(c, &dm.diagB, N1)) return 1;
  NLK_CUDA(cudaMemcpyAsync(dm.diagB, dm.bm1, N1 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  launch_fill(dm.vmult, N1, 1.0, c->st);
  {
    std::vector<double> dA(N1);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < hm.E; ++e) for (int p = 0; p < hm.np1; ++p) dA[(size_t)e * hm.np1 + p] = mesh_diag_local(hm, e, p);
    NLK_CUDA(cudaMemcpyAsync(dm.diagA, dA.data(), N1 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    NLK_CUDA(cudaStreamSynchronize(c->st));
  }
  if (ctx_gs(c, Ptr3{{dm.diagB, dm.vmult, dm.diagA}}, 3)) return 1;
  launch_recip(dm.binvm1, dm.diagB, N1, c->st);
  launch_recip(dm.vmult, dm.vmult, N1, c->st);
  // filter
  if (c->prm.filter_weight > 0) { std::vector<double> F; make_filter_matrix(b, c->prm.filter_weight, c->prm.filter_cutoff, F); if (dev_upload(c, &c->filterF, F)) return 1; }
  // ---- Schwarz + coarse preconditioner
  if (c->prm.precond >= 1) {
    const int np1 = hm.np1;
    // element lengths per direction; neighbour lengths through a face dssum
    std::vector<double> len((size_t)hm.E * d), wl(N1 * (size_t)d, 0.0);
    for (int64_t e = 0; e < hm.E; ++e) for (int k = 0; k < d; ++k) {
      double s = 0; int cnt = 0;
      int nz = d == 3 ? n : 1;
      for (int a2 = 0; a2 < nz; ++a2) for (int a1 = 0; a1 < n; ++a1) {
        // (a1,a2) enumerate the face nodes; compute lo/hi point indices along direction k
        int idx_lo[3], idx_hi[3]; int t = 0;
        for (int dd = 0; dd < 3; ++dd) { if (dd == k) { idx_lo[dd] = 0; idx_hi[dd] = n - 1; } else { int v = (t == 0 ? a1 : a2); idx_lo[dd] = idx_hi[dd] = v; ++t; } }
        if (d == 2) { idx_lo[2] = idx_hi[2] = 0; if (a2 > 0) continue; }
        int plo = (idx_lo[2] * n + idx_lo[1]) * n + idx_lo[0], phi = (idx_hi[2] * n + idx_hi[1]) * n + idx_hi[0];
        double dist = 0;
        for (int cc = 0; cc < d; ++cc) { double dx = hm.xyz[cc][e * np1 + phi] - hm.xyz[cc][e * np1 + plo]; dist += dx * dx; }
        s += std::sqrt(dist); ++cnt;
      }
      len[e * d + k] = s / cnt;
      // write the length on both faces normal to k (all face nodes)
      for (int p = 0; p < np1; ++p) {
        int ijk[3] = {p % n, (p / n) % n, d == 3 ? p / (n * n) : 0};
        if (ijk[k] == 0 || ijk[k] == n - 1) wl[(size_t)k * N1 + e * np1 + p] = len[e * d + k];
      }
    }
    double* dwl = nullptr;
    if (dev_upload(c, &dwl, wl)) return 1;
    if (ctx_gs(c, Ptr3{{dwl, dwl + N1, d == 3 ? dwl + 2 * N1 : nullptr}}, d)) return 1;
    NLK_CUDA(cudaMemcpyAsync(wl.data(), dwl, wl.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    NLK_CUDA(cudaStreamSynchronize(c->st));
    std::vector<double> S((size_t)hm.E * d * n * n), St((size_t)hm.E * d * n * n), dinv(N1, 0.0);
    const int mid = n / 2;
    static const int F_LO[3] = {3, 0, 4}, F_HI[3] = {1, 2, 5};
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < hm.E; ++e) {
      std::vector<double> lam((size_t)d * n); int nact[3] = {1, 1, 1};
      for (int k = 0; k < d; ++k) {
        int ijk_lo[3] = {mid, mid, d == 3 ? mid : 0}, ijk_hi[3] = {mid, mid, d == 3 ? mid : 0};
        ijk_lo[k] = 0; ijk_hi[k] = n - 1;
        double lm = len[e * d + k];
        double ll = wl[(size_t)k * N1 + e * np1 + (ijk_lo[2] * n + ijk_lo[1]) * n + ijk_lo[0]] - lm;
        double lr = wl[(size_t)k * N1 + e * np1 + (ijk_hi[2] * n + ijk_hi[1]) * n + ijk_hi[0]] - lm;
        if (std::fabs(ll) < 1e-12 * lm) ll = 0; if (std::fabs(lr) < 1e-12 * lm) lr = 0;
        auto bc_of = [&](double lnb, int f) { if (lnb > 0) return 0; const auto& cb = hm.cbc_v[e * hm.nfaces + f]; return (cb[0] == 'O' || cb[0] == 'o') ? 1 : 2; };
        int bcl = bc_of(ll, F_LO[k]), bcr = bc_of(lr, F_HI[k]);
        make_fdm_1d(b, lm, ll, lr, bcl, bcr, &S[((size_t)e * d + k) * n * n], &lam[(size_t)k * n], &nact[k]);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) St[((size_t)e * d + k) * n * n + j * n + i] = S[((size_t)e * d + k) * n * n + i * n + j];
      }
      double maxden = 0;
      for (int p = 0; p < np1; ++p) {
        int i = p % n, j = (p / n) % n, k = d == 3 ? p / (n * n) : 0;
        bool act = i < nact[0] && j < nact[1] && (d == 2 || k < nact[2]);
        if (act) { double den = lam[i] + lam[n + j] + (d == 3 ? lam[2 * n + k] : 0.0); maxden = std::max(maxden, std::fabs(den)); }
      }
      for (int p = 0; p < np1; ++p) {
        int i = p % n, j = (p / n) % n, k = d == 3 ? p / (n * n) : 0;
        bool act = i < nact[0] && j < nact[1] && (d == 2 || k < nact[2]);
        double den = lam[i] + lam[n + j] + (d == 3 ? lam[2 * n + k] : 0.0);
        dinv[(size_t)e * np1 + p] = (act && den > 1e-12 * maxden) ? 1.0 / den : 0.0;
      }
    }
    if (dev_upload(c, &dm.fdmS, S) || dev_upload(c, &dm.fdmSt, St) || dev_upload(c, &dm.fdmDinv, dinv)) return 1;
    // overlap-count weights
    if (dev_alloc(c, &dm.swt, N2)) return 1;
    launch_schwarz_embed(dm, c->ones2, nullptr, c->sw_w, c->st);
    if (ctx_gs(c, Ptr3{{c->sw_w, nullptr, nullptr}}, 1)) return 1;
    launch_schwarz_count(dm, c->sw_w, c->sw_z, c->sw_t, c->st);
    if (ctx_gs(c, Ptr3{{c->sw_t, nullptr, nullptr}}, 1)) return 1;
    launch_schwarz_gather_nowt(dm, c->sw_z, c->sw_t, dm.swt, c->st);
    launch_recip(dm.swt, dm.swt, N2, c->st);
    c->have_schwarz = true;
    if (swf_setup(c)) return 1;
    // coarse operator A0 = R0 E R0^T, column by column on the device; dense inverse on the host
    const int64_t nvt = hm.nvert;
    if (nvt <= 5000 && c->prm.precond != 2 && c->prm.precond != 4) {
      if (dev_alloc(c, &c->crs_part, (size_t)hm.E << d) || dev_alloc(c, &c->crs_r, nvt) || dev_alloc(c, &c->crs_y, nvt)) return 1;
      double* dA0 = nullptr; if (dev_alloc(c, &dA0, (size_t)nvt * nvt)) return 1;
      for (int64_t v = 0; v < nvt; ++v) {
        NLK_CUDA(cudaMemsetAsync(c->crs_y, 0, nvt * sizeof(double), c->st));
        launch_fill(c->crs_y + v, 1, 1.0, c->st);
        launch_coarse_prolong_add(dm, c->crs_y, c->pw[0], 0, c->st);
        if (apply_E(c, c->pw[0], c->pw[1], nullptr)) return 1;
        launch_coarse_restrict(dm, c->pw[1], nullptr, c->crs_part, dA0 + (size_t)v * nvt, c->st);
      }
      if (ctx_allreduce(c, dA0, (int)(nvt * nvt), false)) return 1;
      std::vector<double> A0((size_t)nvt * nvt);
      NLK_CUDA(cudaMemcpyAsync(A0.data(), dA0, A0.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
      NLK_CUDA(cudaStreamSynchronize(c->st));
      for (int64_t i = 0; i < nvt; ++i) for (int64_t j = 0; j < i; ++j) { double s = 0.5 * (A0[i * nvt + j] + A0[j * nvt + i]); A0[i * nvt + j] = A0[j * nvt + i] = s; }
      if (!hm.has_outflow) {   // singular (constant null space): shift along 1 1^T -- the solution changes by a constant, removed by ortho
        double tr = 0; for (int64_t i = 0; i < nvt; ++i) tr += A0[i * nvt + i];
        const double sft = tr / ((double)nvt * (double)nvt);
        for (size_t i = 0; i < A0.size(); ++i) A0[i] += sft;
      }
      if (spd_inverse((int)nvt, A0.data(), false)) return 1;
      dm.A0inv = dA0;
      NLK_CUDA(cudaMemcpyAsync(dA0, A0.data(), A0.size() * sizeof(double), cudaMemcpyHostToDevice, c->st));
      NLK_CUDA(cudaStreamSynchronize(c->st));
      c->have_coarse = true;
    } else if (c->prm.precond != 2) {
      c->crs_iters = c->prm.coarse_iters;
      if (coarse_setup_sparse(c)) return 1;
    }
  }
  NLK_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}
}  // namespace nlk

extern "C" {

int nlk_params_default(nlk_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->viscosity = 1.0; p->density = 1.0; p->torder = 3; p->vtol = 1e-9; p->ptol = 1e-7; p->ifheat = 0; p->conductivity = 1.0; p->rhocp = 1.0;
  p->ttol = 1e-9; p->filter_weight = 0.0; p->filter_cutoff = 1.0; p->cg_maxit = 1000; p->gmres_maxit = 100; p->lgmres = 30; p->precond = 1;
  p->pr_proj = 0; p->cfl_limit = 0.5; p->rst_mode = 0; p->coarse_iters = 8;
  return 0;
}

int nlk_comm_unique_id(char id[128]) {
  Nccl n; if (nccl_load(n)) return 1;
  NcclId u; int r = n.GetUniqueId(&u); if (r) { set_error("ncclGetUniqueId failed"); return 1; }
  std::memcpy(id, u.internal, 128); return 0;
}

int nlk_ctx_create(const nlk_mesh* m, const nlk_params* p, int32_t device, nlk_ctx** out) {
  if (!m || !p || !out) { set_error("null argument"); return 1; }
  if (check_device()) return 1;
  if (p->torder < 1 || p->torder > 3) { set_error("torder must be 1..3"); return 1; }
  if (p->lgmres < 1 || p->lgmres > 200) { set_error("lgmres out of range"); return 1; }
  if (p->pr_proj < 0 || p->pr_proj > 30) { set_error("pr_proj out of range [0,30]"); return 1; }
  NLK_CUDA(cudaSetDevice(device));
  nlk_ctx* c = new nlk_ctx(); c->mesh = m; c->prm = *p; c->device = device;
  NLK_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
  c->ph.on = getenv("NLK_PHASES") != nullptr;
  if (!getenv("NLK_NO_STREAM2")) {
    // highest priority: the coarse branch is a chain of tiny kernels whose blocks must slip in between the waves of the
    // Schwarz kernels on the main stream instead of queueing behind each of them
    int prio_lo = 0, prio_hi = 0;
    NLK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    NLK_CUDA(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, prio_hi));
    NLK_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming)); NLK_CUDA(cudaEventCreateWithFlags(&c->ev_crs, cudaEventDisableTiming));
  }
  *out = c;
  if (m->hm.nranks > 1) return 0;            // multi-rank: setup is finished by nlk_ctx_comm_init
  if (ctx_setup(c)) return 1;
  return 0;
}

int nlk_ctx_comm_init(nlk_ctx* c, const char id[128], int32_t rank, int32_t nranks) {
  if (nranks != c->mesh->hm.nranks || rank != c->mesh->hm.rank) { set_error("rank/nranks do not match the mesh partition"); return 1; }
  if (nranks > 1) {
    if (nccl_load(c->nccl)) return 1;
    NcclId u; std::memcpy(u.internal, id, 128);
    int r = c->nccl.CommInitRank(&c->nccl.comm, nranks, u, rank);
    if (r) { set_error(std::string("ncclCommInitRank failed: ") + c->nccl.GetErrorString(r)); return 1; }
    c->nccl.rank = rank; c->nccl.nranks = nranks;
    if (c->nccl.CommSplit && c->st2 && !getenv("NLK_NO_COMM2")) {
      if (c->nccl.CommSplit(c->nccl.comm, 0, rank, &c->nccl.comm2, nullptr) != 0) c->nccl.comm2 = nullptr;     // fall back to the single-stream path
    }
  }
  return ctx_setup(c);
}

int nlk_ctx_destroy(nlk_ctx* c) {
  if (!c) return 0;
  cudaStreamSynchronize(c->st);
  if (c->ph.on && c->ph.steps > 0) {
    static const char* names[nlk::PH_COUNT] = {"makef (convection, ext/bdf)", "velocity residual", "Helmholtz PCG", "pressure rhs", "pressure solve (total)",
                                               "  preconditioner", "  E apply", "  orthogonalisation + host", "velocity correction", "heat", "filter"};
    fprintf(stderr, "[nlk phases] %ld steps, rank %d\n", c->ph.steps, c->nccl.rank);
    for (int i = 0; i < nlk::PH_COUNT; ++i)
      if (c->ph.calls[i]) fprintf(stderr, "[nlk phases] %-32s %9.3f ms/step  (%6.1f calls/step)\n", names[i], c->ph.ms[i] / c->ph.steps, (double)c->ph.calls[i] / c->ph.steps);
    for (cudaEvent_t e : c->ph.pool) cudaEventDestroy(e);
  }
  if (c->crs_graph) cudaGraphExecDestroy(c->crs_graph);
  swf_release(c);
  for (void* p : c->allocs) cudaFree(p);
  if (c->h_sc) cudaFreeHost(c->h_sc);
  if (c->h_red) cudaFreeHost(c->h_red);
  if (c->h_seq) cudaFreeHost(c->h_seq);
  if (c->nccl.comm2) c->nccl.CommDestroy(c->nccl.comm2);
  if (c->nccl.comm) c->nccl.CommDestroy(c->nccl.comm);
  if (c->st2) { cudaStreamSynchronize(c->st2); cudaStreamDestroy(c->st2); cudaEventDestroy(c->ev_in); cudaEventDestroy(c->ev_crs); }
  cudaStreamDestroy(c->st);
  delete c->bank2;
  delete c; return 0;
}
// setup_nek(vtol=, ptol=) (src/neklab_nek_setup.f90:226-230): param(22) = vtol also becomes restol(:) / atol(:) of EVERY field, so the
// temperature Helmholtz tolerance follows the velocity one
int nlk_ctx_set_tol(nlk_ctx* c, double vtol, double ptol) { c->prm.vtol = vtol; c->prm.ptol = ptol; c->prm.ttol = vtol; return 0; }
int nlk_ctx_set_dt(nlk_ctx* c, double dt) { if (dt <= 0) { set_error("dt must be positive"); return 1; } c->dt = dt; return 0; }
int nlk_ctx_sync(nlk_ctx* c) { NLK_CUDA(cudaStreamSynchronize(c->st)); NLK_CUDA(cudaGetLastError()); return 0; }
void* nlk_ctx_stream(nlk_ctx* c) { return (void*)c->st; }

// neklab forcing registry (src/neklab_nek_forcing.f90): one slot per ipert in 0..lpert (lpert = 1 here)
static int forcing_slot_ok(int32_t ipert, const char* who) {
  if (ipert < 0 || ipert > 1) { set_error(std::string(who) + ": invalid value for ipert (0 = nonlinear solver, 1 = perturbation; lpert = 1)"); return 0; }
  return 1;
}
int nlk_set_neklab_forcing(nlk_ctx* c, const double* fx, const double* fy, const double* fz, int32_t ipert) {
  if (!forcing_slot_ok(ipert, "set_neklab_forcing")) return 1;
  const double* f[3] = {fx, fy, fz};
  for (int k = 0; k < c->dm.ndim; ++k) {
    if (!c->forcing[ipert][k]) { if (dev_alloc(c, &c->forcing[ipert][k], c->dm.N1)) return 1; }
    if (f[k]) NLK_CUDA(cudaMemcpyAsync(c->forcing[ipert][k], f[k], c->dm.N1 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    else NLK_CUDA(cudaMemsetAsync(c->forcing[ipert][k], 0, c->dm.N1 * sizeof(double), c->st));
  }
  NLK_CUDA(cudaStreamSynchronize(c->st));
  c->has_forcing[ipert] = true; return 0;
}
int nlk_get_neklab_forcing(nlk_ctx* c, double* fx, double* fy, double* fz, int32_t ipert) {
  if (!forcing_slot_ok(ipert, "get_neklab_forcing")) return 1;
  double* f[3] = {fx, fy, fz};
  for (int k = 0; k < c->dm.ndim; ++k) {
    if (!f[k]) continue;
    if (c->has_forcing[ipert]) NLK_CUDA(cudaMemcpyAsync(f[k], c->forcing[ipert][k], c->dm.N1 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    else std::memset(f[k], 0, c->dm.N1 * sizeof(double));
  }
  NLK_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}
int nlk_zero_neklab_forcing_ipert(nlk_ctx* c, int32_t ipert) {
  if (!forcing_slot_ok(ipert, "zero_neklab_forcing_ipert")) return 1;
  c->has_forcing[ipert] = false; return 0;            // a zero slot is skipped by the step (same arithmetic as adding zeros)
}
int nlk_zero_neklab_forcing(nlk_ctx* c) { c->has_forcing[0] = c->has_forcing[1] = false; return 0; }
int nlk_ctx_set_forcing(nlk_ctx* c, const double* fx, const double* fy, const double* fz) { return nlk_set_neklab_forcing(c, fx, fy, fz, 1); }

// ============================================================================================== nek_dvector
int nlk_vec_create(nlk_ctx* c, nlk_vec** out) {
  nlk_vec* v = new nlk_vec(); v->c = c;
  for (int k = 0; k < c->dm.ndim; ++k) if (dev_alloc(c, &v->v[k], c->dm.N1)) return 1;
  if (dev_alloc(c, &v->pr, c->dm.N2)) return 1;
  if (c->prm.ifheat && dev_alloc(c, &v->theta, c->dm.N1)) return 1;      // no temperature storage without ifheat (Krylov bases at 100k elements)
  *out = v; return 0;
}
int nlk_vec_destroy(nlk_vec* v) {
  if (!v) return 0;
  nlk_ctx* c = v->c;
  auto rel = [&](double* p) { if (!p) return; auto it = std::find(c->allocs.begin(), c->allocs.end(), (void*)p); if (it != c->allocs.end()) { cudaFree(p); c->allocs.erase(it); } };
  cudaStreamSynchronize(c->st);
  for (int k = 0; k < 3; ++k) { rel(v->v[k]); for (int s = 0; s < 2; ++s) rel(v->rv[s][k]); }
  rel(v->pr); rel(v->theta);
  for (int s = 0; s < 2; ++s) { rel(v->rpr[s]); rel(v->rth[s]); }
  delete v; return 0;
}
}  // extern "C"

namespace nlk {
int vec_alloc_rst(nlk_vec* v, int s) {
  nlk_ctx* c = v->c;
  if (v->rpr[s]) return 0;
  for (int k = 0; k < c->dm.ndim; ++k) if (dev_alloc(c, &v->rv[s][k], c->dm.N1)) return 1;
  if (dev_alloc(c, &v->rpr[s], c->dm.N2)) return 1;
  if (c->prm.ifheat && dev_alloc(c, &v->rth[s], c->dm.N1)) return 1;
  return 0;
}
void vec_release_rst(nlk_vec* v) {
  nlk_ctx* c = v->c;
  if (!v->rpr[0] && !v->rpr[1]) { v->nrst = 0; return; }
  auto rel = [&](double*& p) { if (!p) return; auto it = std::find(c->allocs.begin(), c->allocs.end(), (void*)p); if (it != c->allocs.end()) { cudaFree(p); c->allocs.erase(it); } p = nullptr; };
  cudaStreamSynchronize(c->st);
  for (int s = 0; s < 2; ++s) { for (int k = 0; k < 3; ++k) rel(v->rv[s][k]); rel(v->rpr[s]); rel(v->rth[s]); }
  v->nrst = 0;
}
static int copy_fields(nlk_ctx* c, double* const dv[3], double* dp, double* dt, const double* const sv[3], const double* sp, const double* stt) {
  const DevMesh& dm = c->dm;
  for (int k = 0; k < dm.ndim; ++k) NLK_CUDA(cudaMemcpyAsync(dv[k], sv[k], dm.N1 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  NLK_CUDA(cudaMemcpyAsync(dp, sp, dm.N2 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  if (c->prm.ifheat) NLK_CUDA(cudaMemcpyAsync(dt, stt, dm.N1 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  return 0;
}
}  // namespace nlk

extern "C" {

int nlk_vec_copy(nlk_vec* dst, const nlk_vec* src) {
  nlk_ctx* c = dst->c;
  if (copy_fields(c, dst->v, dst->pr, dst->theta, src->v, src->pr, src->theta)) return 1;
  dst->nrst = src->nrst;
  for (int s = 0; s < src->nrst; ++s) {
    if (vec_alloc_rst(dst, s)) return 1;
    if (copy_fields(c, dst->rv[s], dst->rpr[s], dst->rth[s], src->rv[s], src->rpr[s], src->rth[s])) return 1;
  }
  return 0;
}
int nlk_vec_zero(nlk_vec* v) {
  nlk_ctx* c = v->c; const DevMesh& dm = c->dm;
  for (int k = 0; k < dm.ndim; ++k) NLK_CUDA(cudaMemsetAsync(v->v[k], 0, dm.N1 * sizeof(double), c->st));
  NLK_CUDA(cudaMemsetAsync(v->pr, 0, dm.N2 * sizeof(double), c->st));
  if (v->theta) NLK_CUDA(cudaMemsetAsync(v->theta, 0, dm.N1 * sizeof(double), c->st));
  v->nrst = 0; return 0;
}
int nlk_vec_scal(nlk_vec* v, double a) {
  nlk_ctx* c = v->c; const DevMesh& dm = c->dm;
  auto sc = [&](double* const f[3], double* p, double* t) {
    for (int k = 0; k < dm.ndim; ++k) launch_lin(f[k], dm.N1, a, f[k], 0, nullptr, 0, nullptr, 0, nullptr, nullptr, c->st);
    launch_lin(p, dm.N2, a, p, 0, nullptr, 0, nullptr, 0, nullptr, nullptr, c->st);
    if (c->prm.ifheat) launch_lin(t, dm.N1, a, t, 0, nullptr, 0, nullptr, 0, nullptr, nullptr, c->st);
  };
  sc(v->v, v->pr, v->theta);
  for (int s = 0; s < v->nrst; ++s) sc(v->rv[s], v->rpr[s], v->rth[s]);
  return 0;
}
// self = alpha*x + beta*self ; rst slots of self get alpha * (x's CURRENT fields)  (real_vectors.f90:186-200)
int nlk_vec_axpby(double alpha, const nlk_vec* x, double beta, nlk_vec* self) {
  nlk_ctx* c = self->c; const DevMesh& dm = c->dm;
  auto ax = [&](double* const f[3], double* p, double* t) {
    for (int k = 0; k < dm.ndim; ++k) launch_lin(f[k], dm.N1, beta, f[k], alpha, x->v[k], 0, nullptr, 0, nullptr, nullptr, c->st);
    launch_lin(p, dm.N2, beta, p, alpha, x->pr, 0, nullptr, 0, nullptr, nullptr, c->st);
    if (c->prm.ifheat) launch_lin(t, dm.N1, beta, t, alpha, x->theta, 0, nullptr, 0, nullptr, nullptr, c->st);
  };
  ax(self->v, self->pr, self->theta);
  const int mode = c->prm.rst_mode;
  for (int s = 0; s < self->nrst; ++s) {
    if (mode == 1 && x->nrst > s) {      // experimental: consistent rst combination (NOT the reference behaviour)
      for (int k = 0; k < dm.ndim; ++k) launch_lin(self->rv[s][k], dm.N1, beta, self->rv[s][k], alpha, x->rv[s][k], 0, nullptr, 0, nullptr, nullptr, c->st);
      launch_lin(self->rpr[s], dm.N2, beta, self->rpr[s], alpha, x->rpr[s], 0, nullptr, 0, nullptr, nullptr, c->st);
      if (c->prm.ifheat) launch_lin(self->rth[s], dm.N1, beta, self->rth[s], alpha, x->rth[s], 0, nullptr, 0, nullptr, nullptr, c->st);
    } else ax(self->rv[s], self->rpr[s], self->rth[s]);
  }
  return 0;
}
int nlk_vec_dot(const nlk_vec* a, const nlk_vec* b, double* out) {
  nlk_ctx* c = a->c; const DevMesh& dm = c->dm;
  CPtr4 A{}, B{}; int np = 0;
  for (int k = 0; k < dm.ndim; ++k) { A.p[np] = a->v[k]; B.p[np] = b->v[k]; ++np; }
  if (c->prm.ifheat) { A.p[np] = a->theta; B.p[np] = b->theta; ++np; }
  launch_dot(dm.N1, A, B, np, dm.bm1, c->d_red + 310, c->red, c->st);
  if (ctx_allreduce(c, c->d_red + 310, 1, false)) return 1;
  NLK_CUDA(cudaMemcpyAsync(c->h_red + 310, c->d_red + 310, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  NLK_CUDA(cudaStreamSynchronize(c->st));
  *out = c->h_red[310]; return 0;
}
int nlk_vec_norm(const nlk_vec* a, double* out) { double d; if (nlk_vec_dot(a, a, &d)) return 1; *out = std::sqrt(d); return 0; }
int nlk_vec_size(const nlk_vec* v, int64_t* n) {
  const DevMesh& dm = v->c->dm;
  *n = (int64_t)dm.ndim * dm.N1 + dm.N2 + (v->c->prm.ifheat ? dm.N1 : 0); return 0;
}
int nlk_vec_save_rst(nlk_vec* self, const nlk_vec* st, int32_t irst) {
  nlk_ctx* c = self->c;
  if (irst < 1 || irst >= c->prm.torder) { set_error("save_rst: cannot save rst field irst for this temporal order"); return 1; }
  if (vec_alloc_rst(self, irst - 1)) return 1;
  if (copy_fields(c, self->rv[irst - 1], self->rpr[irst - 1], self->rth[irst - 1], st->v, st->pr, st->theta)) return 1;
  self->nrst = std::max(self->nrst, irst); return 0;
}
int nlk_vec_get_rst(const nlk_vec* self, nlk_vec* out, int32_t irst) {
  nlk_ctx* c = self->c;
  if (irst < 1 || irst > self->nrst) { set_error("get_rst: no rst field to retrieve"); return 1; }
  if (copy_fields(c, out->v, out->pr, out->theta, self->rv[irst - 1], self->rpr[irst - 1], self->rth[irst - 1])) return 1;
  out->nrst = 0; return 0;
}
int nlk_vec_nrst(const nlk_vec* v, int32_t* nrst) { *nrst = v->nrst; return 0; }
int nlk_vec_clear_rst(nlk_vec* v) { v->nrst = 0; return 0; }

int nlk_vec_upload(nlk_vec* v, const double* vx, const double* vy, const double* vz, const double* pr, const double* theta) {
  nlk_ctx* c = v->c; const DevMesh& dm = c->dm;
  const double* f[3] = {vx, vy, vz};
  for (int k = 0; k < dm.ndim; ++k) if (f[k]) NLK_CUDA(cudaMemcpyAsync(v->v[k], f[k], dm.N1 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  if (pr) NLK_CUDA(cudaMemcpyAsync(v->pr, pr, dm.N2 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  if (theta && v->theta) NLK_CUDA(cudaMemcpyAsync(v->theta, theta, dm.N1 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  NLK_CUDA(cudaStreamSynchronize(c->st));
  v->nrst = 0; return 0;
}
int nlk_vec_download(const nlk_vec* v, double* vx, double* vy, double* vz, double* pr, double* theta) {
  nlk_ctx* c = v->c; const DevMesh& dm = c->dm;
  double* f[3] = {vx, vy, vz};
  for (int k = 0; k < dm.ndim; ++k) if (f[k]) NLK_CUDA(cudaMemcpyAsync(f[k], v->v[k], dm.N1 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  if (pr) NLK_CUDA(cudaMemcpyAsync(pr, v->pr, dm.N2 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  if (theta && v->theta) NLK_CUDA(cudaMemcpyAsync(theta, v->theta, dm.N1 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  else if (theta) std::memset(theta, 0, dm.N1 * sizeof(double));
  NLK_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

// ---- nek_zvector (complex_vectors.f90): (re, im) pairs of nek_dvector; rst slots follow the real-vector rules
static int z_fields(nlk_ctx* c, const nlk_vec* v, int slot, double* out[5], size_t n_out[5]) {
  const DevMesh& dm = c->dm; int k = 0;
  for (int f = 0; f < dm.ndim; ++f) { out[k] = slot < 0 ? v->v[f] : v->rv[slot][f]; n_out[k++] = dm.N1; }
  out[k] = slot < 0 ? v->pr : v->rpr[slot]; n_out[k++] = dm.N2;
  if (c->prm.ifheat) { out[k] = slot < 0 ? v->theta : v->rth[slot]; n_out[k++] = dm.N1; }
  return k;
}
int nlk_zvec_scal(nlk_vec* re, nlk_vec* im, double ar, double ai) {
  nlk_ctx* c = re->c;
  if (re->nrst != im->nrst) { set_error("zscal: re/im carry different numbers of rst fields"); return 1; }
  for (int slot = -1; slot < re->nrst; ++slot) {
    double* R[5]; double* I[5]; size_t n[5];
    int nf = z_fields(c, re, slot, R, n); z_fields(c, im, slot, I, n);
    for (int f = 0; f < nf; ++f) {           // (re, im) <- (ar*re - ai*im, ar*im + ai*re), through one scratch field
      double* t = n[f] == c->dm.N1 ? c->wk[0] : c->pw[0];
      launch_lin(t, n[f], ar, R[f], -ai, I[f], 0, nullptr, 0, nullptr, nullptr, c->st);
      launch_lin(I[f], n[f], ar, I[f], ai, R[f], 0, nullptr, 0, nullptr, nullptr, c->st);
      NLK_CUDA(cudaMemcpyAsync(R[f], t, n[f] * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    }
  }
  return 0;
}
int nlk_zvec_axpby(double ar, double ai, const nlk_vec* xre, const nlk_vec* xim, double br, double bi, nlk_vec* sre, nlk_vec* sim) {
  // self = alpha*x + beta*self (complex); the reference does zscal(beta) then adds alpha*x built in a scratch zvector
  if (nlk_zvec_scal(sre, sim, br, bi)) return 1;
  if (nlk_vec_axpby(ar, xre, 1.0, sre) || nlk_vec_axpby(-ai, xim, 1.0, sre)) return 1;
  if (nlk_vec_axpby(ar, xim, 1.0, sim) || nlk_vec_axpby(ai, xre, 1.0, sim)) return 1;
  return 0;
}
int nlk_zvec_dot(const nlk_vec* sre, const nlk_vec* sim, const nlk_vec* xre, const nlk_vec* xim, double* out_re, double* out_im) {
  double a, b, cc, d;                          // conj(s).x = (sre.xre + sim.xim) + i (sre.xim - sim.xre)
  if (nlk_vec_dot(sre, xre, &a) || nlk_vec_dot(sim, xim, &b) || nlk_vec_dot(sre, xim, &cc) || nlk_vec_dot(sim, xre, &d)) return 1;
  *out_re = a + b; *out_im = cc - d; return 0;
}

// seeded, C0, BC-satisfying random vector (nek_drand: real_vectors.f90:52-123 -- random_number is compiler-specific,
// so the generator is a documented replacement: smooth field + splitmix64 noise, then dssum*vmult, mask, normalise)
int nlk_vec_rand(nlk_vec* v, int32_t ifnorm, uint64_t seed) {
  nlk_ctx* c = v->c; const DevMesh& dm = c->dm;
  if (nlk_vec_zero(v)) return 1;
  for (int k = 0; k < dm.ndim; ++k) launch_rand_field_impl(v->v[k], c->xyz[0], c->xyz[1], dm.ndim == 3 ? c->xyz[2] : nullptr, c->d_lglel, dm.np1, dm.N1, seed, k, c->st);
  if (ctx_gs(c, Ptr3{{v->v[0], v->v[1], v->v[2]}}, dm.ndim)) return 1;
  for (int k = 0; k < dm.ndim; ++k) launch_axpy_mm(v->v[k], dm.N1, nullptr, 1.0, v->v[k], dm.vmult, dm.mask[k], c->st);
  if (c->prm.ifheat) {
    launch_rand_field_impl(v->theta, c->xyz[0], c->xyz[1], dm.ndim == 3 ? c->xyz[2] : nullptr, c->d_lglel, dm.np1, dm.N1, seed, 7, c->st);
    if (ctx_gs(c, Ptr3{{v->theta, nullptr, nullptr}}, 1)) return 1;
    launch_axpy_mm(v->theta, dm.N1, nullptr, 1.0, v->theta, dm.vmult, dm.mask[3], c->st);
  }
  if (ifnorm) { double nr; if (nlk_vec_norm(v, &nr)) return 1; if (nlk_vec_scal(v, 1.0 / nr)) return 1; }
  v->nrst = 0; return 0;
}

// ============================================================================================== exptA
int nlk_exptA_create(nlk_ctx* c, double tau, const nlk_vec* bf, nlk_op** out) {
  nlk_op* op = new nlk_op(); op->c = c; op->tau = tau;
  if (nlk_vec_create(c, &op->baseflow)) return 1;
  if (nlk_vec_copy(op->baseflow, bf)) return 1;
  *out = op; return 0;
}
int nlk_exptA_destroy(nlk_op* op) { if (!op) return 0; nlk_vec_destroy(op->baseflow); delete op; return 0; }
int nlk_exptA_set_tau(nlk_op* op, double tau) { op->tau = tau; return 0; }
}  // extern "C"

namespace nlk {
static int push_baseflow(nlk_op* op) {            // vec2nek(vx,vy,vz,pr,t, self%baseflow)
  nlk_ctx* c = op->c; const DevMesh& dm = c->dm;
  for (int k = 0; k < dm.ndim; ++k) NLK_CUDA(cudaMemcpyAsync(c->U[k], op->baseflow->v[k], dm.N1 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  if (op->baseflow->theta) NLK_CUDA(cudaMemcpyAsync(c->T, op->baseflow->theta, dm.N1 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
  return 0;
}
static int state_from_vec(nlk_ctx* c, const double* const v[3], const double* pr, const double* th) {
  return copy_fields(c, c->vp, c->prp, c->tp, v, pr, th);
}

// exptA_proj_linop%proj (exponential_propagator_proj.f90:135-173) on d velocity fields
int exptA_project(nlk_op* op, double* const v[3]) {
  nlk_ctx* c = op->c; const DevMesh& dm = c->dm;
  for (int k = 0; k < dm.ndim; ++k)
    launch_planar_proj(v[k], dm.bm1, op->proj_cv, op->proj_sv, op->proj_off, op->proj_idx, op->proj_gid, op->proj_ngroups, dm.N1, op->proj_coef, c->st);
  return 0;
}
struct EvPair {            // timing events released on every return path
  cudaEvent_t a = nullptr, b = nullptr;
  EvPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EvPair() { cudaEventDestroy(a); cudaEventDestroy(b); }
};
// exptA_matvec / exptA_rmatvec (src/linops/exponential_propagator.f90:15-107) incl. get_rst / compute_rst (:109-142)
int exptA_apply(nlk_op* op, const nlk_vec* in, nlk_vec* out, bool transpose) {
  nlk_ctx* c = op->c;
  if (in == out) { set_error("exptA: vec_in and vec_out must differ"); return 1; }
  const int nrst = c->prm.torder - 1;
  if (sync_cg_counter(c)) return 1;
  EvPair ev; cudaEvent_t e0 = ev.a, e1 = ev.b;
  long l0 = g_launches; long cg0 = c->cg_iters, gm0 = c->gmres_iters, st0 = c->steps;
  cudaEventRecord(e0, c->st);
  if (push_baseflow(op)) return 1;
  if (step_setup(c, op->tau, transpose)) return 1;
  if (state_from_vec(c, in->v, in->pr, in->theta)) return 1;
  if (op->proj_dir && exptA_project(op, c->vp)) return 1;                               // exptA_proj: project out unwanted wavenumbers (:46-47)
  if (c->prm.step_variant & 4) NLK_CUDA(cudaMemsetAsync(c->prp, 0, c->dm.N2 * sizeof(double), c->st));
  if (reset_history_pub(c)) return 1;
  for (int istep = 1; istep <= c->nsteps; ++istep) {
    if (step_advance(c, istep)) return 1;
    if (istep <= nrst && in->nrst > 0 && c->prm.rst_mode != 2) {
      if (istep > in->nrst) { set_error("exptA: input vector has fewer rst fields than the temporal order needs"); return 1; }
      if (state_from_vec(c, in->rv[istep - 1], in->rpr[istep - 1], in->rth[istep - 1])) return 1;
    }
  }
  if (op->proj_dir && exptA_project(op, c->vp)) return 1;                               // (:64-65)
  if (copy_fields(c, out->v, out->pr, out->theta, c->vp, c->prp, c->tp)) return 1;       // nek2vec (intent(out): nrst reset)
  out->nrst = 0;
  for (int k = 1; k <= nrst; ++k) {
    if (step_advance(c, c->nsteps + k)) return 1;
    if (vec_alloc_rst(out, k - 1)) return 1;
    if (copy_fields(c, out->rv[k - 1], out->rpr[k - 1], out->rth[k - 1], c->vp, c->prp, c->tp)) return 1;
    out->nrst = std::max(out->nrst, k);
  }
  cudaEventRecord(e1, c->st);
  NLK_CUDA(cudaStreamSynchronize(c->st));
  if (sync_cg_counter(c)) return 1;
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  op->stats.nsteps = c->nsteps; op->stats.dt = c->dt; op->stats.ms_total = ms; op->stats.launches = g_launches - l0;
  op->stats.cg_iters = c->cg_iters - cg0; op->stats.gmres_iters = c->gmres_iters - gm0; op->stats.steps = c->steps - st0; op->stats.matvecs += 1;
  return 0;
}
}  // namespace nlk

extern "C" {
int nlk_exptA_init(nlk_op* op) { if (push_baseflow(op)) return 1; return step_setup(op->c, op->tau, false); }
int nlk_exptA_matvec(nlk_op* op, const nlk_vec* in, nlk_vec* out) { return exptA_apply(op, in, out, false); }
int nlk_exptA_rmatvec(nlk_op* op, const nlk_vec* in, nlk_vec* out) { return exptA_apply(op, in, out, true); }
int nlk_exptA_set_baseflow(nlk_op* op, const nlk_vec* bf) { return nlk_vec_copy(op->baseflow, bf); }
int nlk_exptA_set_projection(nlk_op* op, double alpha, int32_t idir) {
  nlk_ctx* c = op->c; const HostMesh& hm = c->mesh->hm; const int d = hm.ndim;
  if (idir == 0) { op->proj_dir = 0; return 0; }
  if (idir < 1 || idir > d) { set_error("exptA_proj: idir must be 1..ndim (0 = off)"); return 1; }
  if (hm.nranks > 1) { set_error("exptA_proj_linop is single-rank only"); return 1; }
  const size_t N1 = (size_t)hm.E * hm.np1; const int ax = idir - 1;
  // plane groups: points sharing their transverse coordinates (Nek gtpp_gs_setup on a tensor-product box mesh)
  double ext = 0; for (int k = 0; k < d; ++k) { auto mm = std::minmax_element(hm.xyz[k].begin(), hm.xyz[k].end()); ext = std::max(ext, *mm.second - *mm.first); }
  const double tol = 1e-8 * (ext > 0 ? ext : 1.0);
  std::vector<std::array<int64_t, 3>> key(N1);
  for (size_t i = 0; i < N1; ++i) { int t = 0; key[i] = {0, 0, (int64_t)i}; for (int k = 0; k < d; ++k) if (k != ax) key[i][t++] = (int64_t)std::llround(hm.xyz[k][i] / tol); }
  std::sort(key.begin(), key.end());
  std::vector<int32_t> off(1, 0), idx(N1), gid(N1);
  for (size_t i = 0; i < N1; ++i) {
    if (i > 0 && (key[i][0] != key[i - 1][0] || key[i][1] != key[i - 1][1])) off.push_back((int32_t)i);
    idx[i] = (int32_t)key[i][2]; gid[key[i][2]] = (int32_t)off.size() - 1;
  }
  off.push_back((int32_t)N1);
  std::vector<double> cv(N1), sv(N1);
  for (size_t i = 0; i < N1; ++i) { cv[i] = std::cos(alpha * hm.xyz[ax][i]); sv[i] = std::sin(alpha * hm.xyz[ax][i]); }
  op->proj_ngroups = (int64_t)off.size() - 1;
  if (dev_upload(c, &op->proj_off, off) || dev_upload(c, &op->proj_idx, idx) || dev_upload(c, &op->proj_gid, gid) || dev_upload(c, &op->proj_cv, cv) ||
      dev_upload(c, &op->proj_sv, sv) || dev_alloc(c, &op->proj_coef, (size_t)2 * op->proj_ngroups)) return 1;
  op->proj_dir = idir; op->proj_alpha = alpha; return 0;
}
int nlk_exptA_apply_projection(nlk_op* op, nlk_vec* v) {
  if (!op->proj_dir) { set_error("exptA_proj: no projection set"); return 1; }
  return exptA_project(op, v->v);
}
int nlk_nek2vec(nlk_ctx* c, nlk_vec* out) { if (copy_fields(c, out->v, out->pr, out->theta, c->vp, c->prp, c->tp)) return 1; out->nrst = 0; return 0; }
int nlk_vec2nek(nlk_ctx* c, const nlk_vec* in) { return state_from_vec(c, in->v, in->pr, in->theta); }

int nlk_nonlinear_map(nlk_ctx* c, double tau, double cfl_limit, const nlk_vec* in, nlk_vec* out) {
  if (in == out) { set_error("nonlinear_map: vec_in and vec_out must differ"); return 1; }
  c->nonlinear = true; c->adjoint = false;
  int rc = 0;
  do {
    if ((rc = step_setup_cfl(c, tau, cfl_limit, CPtr3{{in->v[0], in->v[1], in->v[2]}}))) break;
    if ((rc = state_from_vec(c, in->v, in->pr, in->theta))) break;
    if ((rc = reset_history_pub(c))) break;
    for (int istep = 1; istep <= c->nsteps && !rc; ++istep) rc = step_advance(c, istep);
    if (rc) break;
    if ((rc = copy_fields(c, out->v, out->pr, out->theta, c->vp, c->prp, c->tp))) break;
    out->nrst = 0;
    rc = nlk_vec_axpby(-1.0, in, 1.0, out);                       // vec_out%sub(vec_in)
  } while (0);
  c->nonlinear = false;
  if (!rc) rc = sync_cg_counter(c);
  return rc;
}

// LightKrylov `newton` with gmres_rdp on the fixed-point Jacobian and neklab's tolerance schedulers
int nlk_newton_fixed_point(nlk_ctx* c, double tau, nlk_vec* X, double tol, int32_t tol_mode, int32_t maxiter, int32_t gmres_kdim,
                           double* rnorm_hist, int32_t* niter, int32_t* info) {
  const double mintol = 10.0 * 1e-15, maxtol = 1.0e-4;
  const double vtol0 = c->prm.vtol, ptol0 = c->prm.ptol, ttol0 = c->prm.ttol, cfl0 = c->prm.cfl_limit;
  nlk_vec *r = nullptr, *dx = nullptr; nlk_op* J = nullptr;
  if (nlk_vec_create(c, &r) || nlk_vec_create(c, &dx) || nlk_exptA_create(c, tau, X, &J)) return 1;
  double tol_s = std::max(tol, mintol);
  *info = 1; int it = 0; int rc = 0;
  for (it = 0; it <= maxiter; ++it) {
    // scheduler (nek_constant_tol | nek_dynamic_tol); the first evaluation uses the target tolerance
    if (it > 0 && tol_mode == 2) {
      double target = std::min(std::max(tol, mintol), maxtol);
      double t = std::max(0.1 * rnorm_hist[it - 1], target);
      if (t < 10 * target) t = target;
      tol_s = std::min(t, maxtol);
    }
    c->prm.vtol = tol_s * 0.1; c->prm.ptol = tol_s * 0.1; c->prm.ttol = c->prm.vtol;          // nonlinear_map: vtol = ptol = atol*0.1 (restol(:) = vtol)
    if ((rc = nlk_nonlinear_map(c, tau, 0.4, X, r))) break;
    double rn; if ((rc = nlk_vec_norm(r, &rn))) break;
    rnorm_hist[it] = rn;
    if (rn < tol) { *info = 0; break; }
    if (it == maxiter) break;
    // Jacobian = exptA(X) - I with vtol = ptol = atol*0.5, cfl 0.5
    c->prm.vtol = tol_s * 0.5; c->prm.ptol = tol_s * 0.5; c->prm.ttol = c->prm.vtol; c->prm.cfl_limit = 0.5;
    if ((rc = nlk_exptA_set_baseflow(J, X))) break;
    if ((rc = nlk_vec_zero(dx))) break;
    int ginfo = 0;
    if ((rc = nlk_gmres(J, 1, r, dx, gmres_kdim, tol_s, std::sqrt(1e-15), 10, 0, &ginfo))) break;
    int nr = X->nrst; X->nrst = 0;
    if ((rc = nlk_vec_axpby(-1.0, dx, 1.0, X))) break;             // X <- X - dX
    X->nrst = nr;
  }
  *niter = it;
  c->prm.vtol = vtol0; c->prm.ptol = ptol0; c->prm.ttol = ttol0; c->prm.cfl_limit = cfl0;
  nlk_exptA_destroy(J); nlk_vec_destroy(r); nlk_vec_destroy(dx);
  return rc;
}

int nlk_exptA_time_steps(nlk_op* op, const nlk_vec* in, int32_t nwarm, int32_t nsteps, double* ms_timed) {
  nlk_ctx* c = op->c;
  if (push_baseflow(op)) return 1;
  if (step_setup(c, op->tau, false)) return 1;
  if (state_from_vec(c, in->v, in->pr, in->theta)) return 1;
  if (reset_history_pub(c)) return 1;
  for (int i = 1; i <= nwarm; ++i) if (step_advance(c, i)) return 1;
  if (sync_cg_counter(c)) return 1;
  EvPair ev; cudaEvent_t e0 = ev.a, e1 = ev.b;
  long l0 = g_launches; long cg0 = c->cg_iters, gm0 = c->gmres_iters, st0 = c->steps;
  cudaEventRecord(e0, c->st);
  for (int i = nwarm + 1; i <= nwarm + nsteps; ++i) if (step_advance(c, i)) return 1;
  cudaEventRecord(e1, c->st);
  NLK_CUDA(cudaStreamSynchronize(c->st));
  if (sync_cg_counter(c)) return 1;
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  *ms_timed = ms;
  op->stats.nsteps = c->nsteps; op->stats.dt = c->dt; op->stats.ms_total = ms; op->stats.launches = g_launches - l0;
  op->stats.cg_iters = c->cg_iters - cg0; op->stats.gmres_iters = c->gmres_iters - gm0; op->stats.steps = c->steps - st0;
  return 0;
}
// resolvent_linop (src/linops/resolvent.f90): evaluate_rhs (:80-112) / evaluate_imaginary_part (:136-166) are the same loop --
// time-harmonic forcing Re(exp(+-i omega t) f) written into the perturbation slot of the forcing registry before every step --
// started from zero (x0 == nullptr) or from x0.  The forcing is combined on the device (no host copy of the fields per step).
static int resolvent_integrate(nlk_op* op, double tau, double omega, const nlk_vec* fre, const nlk_vec* fim, bool adjoint, const nlk_vec* x0, nlk_vec* out) {
  nlk_ctx* c = op->c; const DevMesh& dm = c->dm;
  if (push_baseflow(op)) return 1;
  if (step_setup(c, tau, adjoint)) return 1;                      // exptA%init(): setup_linear_solver(endtime = tau) -> nsteps, dt
  if (x0) { if (state_from_vec(c, x0->v, x0->pr, x0->theta)) return 1; }
  else {                                                         // opzero(vxp, vyp, vzp); the stale prp/tp of the reference are zeroed too
    for (int k = 0; k < dm.ndim; ++k) NLK_CUDA(cudaMemsetAsync(c->vp[k], 0, dm.N1 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(c->prp, 0, dm.N2 * sizeof(double), c->st));
    NLK_CUDA(cudaMemsetAsync(c->tp, 0, dm.N1 * sizeof(double), c->st));
  }
  if (reset_history_pub(c)) return 1;
  for (int k = 0; k < dm.ndim; ++k) if (!c->forcing[1][k]) { if (dev_alloc(c, &c->forcing[1][k], dm.N1)) return 1; }
  const double sign = adjoint ? -1.0 : 1.0;
  int rc = 0;
  for (int istep = 1; istep <= c->nsteps && !rc; ++istep) {
    const double th = sign * omega * (double)(istep - 1) * c->dt;  // `time` before nek_advance
    for (int k = 0; k < dm.ndim; ++k)                              // Re((cos + i sin)(f_re + i f_im))
      launch_lin(c->forcing[1][k], dm.N1, std::cos(th), fre->v[k], -std::sin(th), fim->v[k], 0, nullptr, 0, nullptr, nullptr, c->st);
    c->has_forcing[1] = true;
    rc = step_advance(c, istep);
  }
  c->has_forcing[0] = c->has_forcing[1] = false;                   // zero_neklab_forcing()
  c->adjoint = false;
  if (rc) return 1;
  if (copy_fields(c, out->v, out->pr, out->theta, c->vp, c->prp, c->tp)) return 1;
  out->nrst = 0;
  return sync_cg_counter(c);
}
}  // extern "C"
namespace nlk { int resolvent_apply(nlk_op* op, double omega, const nlk_vec* fre, const nlk_vec* fim, nlk_vec* ore, nlk_vec* oim, bool adjoint, double rtol, int32_t* info); }
int nlk::resolvent_apply(nlk_op* op, double omega, const nlk_vec* fre, const nlk_vec* fim, nlk_vec* ore, nlk_vec* oim, bool adjoint, double rtol, int32_t* info) {
  nlk_ctx* c = op->c;
  if (ore == oim || ore == fre || ore == fim || oim == fre || oim == fim) { set_error("resolvent: input and output vectors must differ"); return 1; }
  const double tau = omega == 0.0 ? 1.0 : 2.0 * M_PI / std::fabs(omega);
  const double tau0 = op->tau;
  nlk_vec* b = nullptr; if (nlk_vec_create(c, &b)) return 1;
  int rc = 0;
  do {
    op->tau = tau;
    if ((rc = resolvent_integrate(op, tau, omega, fre, fim, adjoint, nullptr, b))) break;           // b = forced response from rest over one period
    if ((rc = nlk_vec_scal(b, -1.0))) break;                                                        // (I - exptA) x = b  <=>  (exptA - I) x = -b
    if ((rc = nlk_vec_zero(ore))) break;
    if ((rc = nlk_gmres(op, 1, b, ore, 64, 1.0e-12, rtol, 10, adjoint ? 1 : 0, info))) break;       // gmres_dp_opts(kdim = 64), atol 1e-12, rtol 1e-6
    ore->nrst = 0;
    op->tau = tau / 4.0;
    rc = resolvent_integrate(op, tau / 4.0, omega, fre, fim, adjoint, ore, oim);                    // quarter period from the periodic state
  } while (0);
  op->tau = tau0;
  nlk_vec_destroy(b);
  return rc;
}
extern "C" {
// resolvent_matvec / resolvent_rmatvec (src/linops/resolvent.f90:17-75); rtol <= 0 selects the reference's 1e-6
int nlk_resolvent_matvec(nlk_op* op, double omega, const nlk_vec* fre, const nlk_vec* fim, nlk_vec* ore, nlk_vec* oim, int32_t adjoint, double rtol, int32_t* info) {
  int32_t dummy = 0;
  return nlk::resolvent_apply(op, omega, fre, fim, ore, oim, adjoint != 0, rtol > 0 ? rtol : 1.0e-6, info ? info : &dummy);
}
// one leg of the above, for tests: the forced integration over tau from x0 (or from rest when x0 is NULL)
int nlk_resolvent_integrate(nlk_op* op, double tau, double omega, const nlk_vec* fre, const nlk_vec* fim, int32_t adjoint, const nlk_vec* x0, nlk_vec* out) {
  if (out == fre || out == fim || out == x0) { set_error("resolvent: input and output vectors must differ"); return 1; }
  return resolvent_integrate(op, tau, omega, fre, fim, adjoint != 0, x0, out);
}
// ---- periodic orbits: nek_upo_jacobian (src/systems/periodic_orbit.f90:46-181) on nek_ext_dvector = (fields, T)
// compute_fdot (src/systems/neklab_systems.f90:202-223): (F_dt(X) - X)/dt with ONE first-order nonlinear step from the base
// state held in bank2 (the perturbation part of that nek_advance is skipped: nothing reads it afterwards)
static int upo_fdot(nlk_ctx* c, nlk_vec* vec) {
  const DevMesh& dm = c->dm;
  bank_swap(c);
  int rc = copy_fields(c, vec->v, vec->pr, vec->theta, c->vp, c->prp, c->tp);
  if (!rc) rc = reset_history_pub(c);
  c->nonlinear = true; const bool adj = c->adjoint; c->adjoint = false;
  if (!rc) rc = step_advance(c, 1);
  c->nonlinear = false; c->adjoint = adj;
  if (!rc) {
    const double s = 1.0 / c->dt;
    for (int k = 0; k < dm.ndim; ++k) launch_lin(vec->v[k], dm.N1, s, c->vp[k], -s, vec->v[k], 0, nullptr, 0, nullptr, nullptr, c->st);
    launch_lin(vec->pr, dm.N2, s, c->prp, -s, vec->pr, 0, nullptr, 0, nullptr, nullptr, c->st);
    vec->nrst = 0;
  }
  bank_swap(c);
  return rc;
}
int nlk_upo_jacobian(nlk_ctx* c, const nlk_vec* X, double TX, const nlk_vec* in, double Tin, nlk_vec* out, double* Tout, int32_t transpose) {
  if (in == out || X == out) { set_error("upo jacobian: vec_in / X and vec_out must differ"); return 1; }
  if (bank2_ensure(c)) return 1;
  const DevMesh& dm = c->dm; StateBank* b = c->bank2;
  const int nrst = c->prm.torder - 1;
  const double atol = c->prm.vtol;                                        // param(22)
  const double fac = transpose ? 0.5 : 0.1;                               // :68-69 / :137-138
  nlk_vec* vec = nullptr; if (nlk_vec_create(c, &vec)) return 1;
  double* Usave[3] = {c->U[0], c->U[1], c->U[2]};
  int rc = 0;
  auto base_from_X = [&]() { return copy_fields(c, b->vp, b->prp, b->tp, X->v, X->pr, X->theta); };   // abs_ext_vec2nek(vx, vy, vz, pr, t, self%X)
  do {
    c->prm.vtol = atol * fac; c->prm.ptol = atol * fac;
    if ((rc = base_from_X())) break;
    if ((rc = step_setup_cfl(c, TX, 0.4, CPtr3{{X->v[0], X->v[1], X->v[2]}}))) break;               // endtime = get_period_abs(self%X), cfl_limit = 0.4
    c->adjoint = transpose != 0;
    if ((rc = state_from_vec(c, in->v, in->pr, in->theta))) break;                                    // ext_vec2nek(vxp, ..., vec_in)
    if ((rc = reset_history_pub(c))) break;
    bank_swap(c); rc = reset_history_pub(c); bank_swap(c); if (rc) break;
    for (int k = 0; k < dm.ndim; ++k) c->U[k] = b->vp[k];                                            // the linearisation point is the moving base state
    for (int istep = 1; istep <= c->nsteps && !rc; ++istep) {
      rc = coupled_advance(c, istep);
      if (!rc && istep <= nrst && in->nrst > 0 && c->prm.rst_mode != 2) {                            // jac_get_rst (:216-228)
        if (istep > in->nrst) { set_error("upo jacobian: input vector has fewer rst fields than the temporal order needs"); rc = 1; break; }
        rc = state_from_vec(c, in->rv[istep - 1], in->rpr[istep - 1], in->rth[istep - 1]);
      }
    }
    if (rc) break;
    if ((rc = copy_fields(c, out->v, out->pr, out->theta, c->vp, c->prp, c->tp))) break;              // nek2ext_vec(vec_out, vxp, ...)
    out->nrst = 0;
    for (int k = 1; k <= nrst && !rc; ++k) {                                                          // jac_compute_rst (:183-214): the base flow keeps moving too
      if ((rc = coupled_advance(c, c->nsteps + k))) break;
      if ((rc = vec_alloc_rst(out, k - 1))) break;
      rc = copy_fields(c, out->rv[k - 1], out->rpr[k - 1], out->rth[k - 1], c->vp, c->prp, c->tp);
      out->nrst = std::max(out->nrst, k);
    }
    if (rc) break;
    if ((rc = nlk_vec_axpby(-1.0, in, 1.0, out))) break;                                              // [exp(tau J) - I] dx
    if ((rc = upo_fdot(c, vec))) break;                                                               // f'(X(T)) from where the base trajectory stands (:93-97)
    if ((rc = nlk_vec_axpby(Tin, vec, 1.0, out))) break;                                              // + f'(X(T)) dT
    if ((rc = base_from_X())) break;                                                                  // phase condition at X(0) (:99-103)
    if ((rc = upo_fdot(c, vec))) break;
    double d; if ((rc = nlk_vec_dot(in, vec, &d))) break;                                             // vec%T = 0: no period term
    *Tout = d;
  } while (0);
  for (int k = 0; k < 3; ++k) c->U[k] = Usave[k];
  c->adjoint = false;
  c->prm.vtol = atol; c->prm.ptol = atol;                                                             // "Reset tolerances": param(22) = param(21) = atol (:106-107)
  nlk_vec_destroy(vec);
  if (!rc) rc = sync_cg_counter(c);
  return rc;
}
int nlk_exptA_stats(const nlk_op* op, nlk_stats* out) { *out = op->stats; out->nsteps = op->c->nsteps; out->dt = op->c->dt; return 0; }

// ============================================================================================== test hooks
static int up(nlk_ctx* c, double* d, const double* h, size_t n) { NLK_CUDA(cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, c->st)); return 0; }
static int down(nlk_ctx* c, double* h, const double* d, size_t n) { NLK_CUDA(cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->st)); NLK_CUDA(cudaStreamSynchronize(c->st)); return 0; }

int nlk_test_axhelm(nlk_ctx* c, const double* u, double h1, double h2, double* w) {
  if (up(c, c->wk[0], u, c->dm.N1)) return 1;
  launch_axhelm(c->dm, c->wk[0], c->wk[1], h1, h2, c->st);
  return down(c, w, c->wk[1], c->dm.N1);
}
int nlk_test_dssum(nlk_ctx* c, double* u) {
  if (up(c, c->wk[0], u, c->dm.N1)) return 1;
  if (ctx_gs(c, Ptr3{{c->wk[0], nullptr, nullptr}}, 1)) return 1;
  return down(c, u, c->wk[0], c->dm.N1);
}
int nlk_test_opdiv(nlk_ctx* c, const double* ux, const double* uy, const double* uz, double* p) {
  const double* u[3] = {ux, uy, uz};
  for (int k = 0; k < c->dm.ndim; ++k) if (up(c, c->wk[k], u[k], c->dm.N1)) return 1;
  launch_opdiv(c->dm, CPtr3{{c->wk[0], c->wk[1], c->wk[2]}}, c->pw[0], 1.0, c->st);
  return down(c, p, c->pw[0], c->dm.N2);
}
int nlk_test_opgradt(nlk_ctx* c, const double* p, double* wx, double* wy, double* wz) {
  if (up(c, c->pw[0], p, c->dm.N2)) return 1;
  launch_opgradt(c->dm, c->pw[0], Ptr3{{c->wk[0], c->wk[1], c->wk[2]}}, c->st);
  double* w[3] = {wx, wy, wz};
  for (int k = 0; k < c->dm.ndim; ++k) if (down(c, w[k], c->wk[k], c->dm.N1)) return 1;
  return 0;
}
int nlk_test_cdabdtp(nlk_ctx* c, const double* p, double* ep) {
  if (up(c, c->pw[3], p, c->dm.N2)) return 1;
  if (apply_E(c, c->pw[3], c->pw[4], nullptr)) return 1;
  return down(c, ep, c->pw[4], c->dm.N2);
}
int nlk_test_convect(nlk_ctx* c, const double* u, const double* cx, const double* cy, const double* cz, double* out) {
  const double* cc[3] = {cx, cy, cz};
  if (up(c, c->wk[3], u, c->dm.N1)) return 1;
  for (int k = 0; k < c->dm.ndim; ++k) if (up(c, c->wk[k], cc[k], c->dm.N1)) return 1;
  launch_convect(c->dm, CPtr4{{c->wk[3], nullptr, nullptr, nullptr}}, 1, CPtr3{{c->wk[0], c->wk[1], c->wk[2]}}, Ptr4{{c->wk[4], nullptr, nullptr, nullptr}}, 1.0, 0, c->st);
  return down(c, out, c->wk[4], c->dm.N1);
}
int nlk_test_convect_adj(nlk_ctx* c, const double* const* U, const double* const* cf, double* const* out) {
  const int d = c->dm.ndim;
  for (int k = 0; k < d; ++k) { if (up(c, c->wk[k], U[k], c->dm.N1)) return 1; if (up(c, c->wk[3 + k], cf[k], c->dm.N1)) return 1; }
  launch_convect_adj(c->dm, CPtr3{{c->wk[0], c->wk[1], c->wk[2]}}, CPtr3{{c->wk[3], c->wk[4], c->wk[5]}}, Ptr3{{c->cg_x, c->cg_r, c->cg_p}}, 1.0, 0, c->st);
  double* o[3] = {c->cg_x, c->cg_r, c->cg_p};
  for (int k = 0; k < d; ++k) if (down(c, out[k], o[k], c->dm.N1)) return 1;
  return 0;
}
int nlk_test_helmholtz(nlk_ctx* c, const double* f, double h1, double h2, int32_t comp, double tol, double* x, int32_t* iters) {
  if (up(c, c->wk[7], f, c->dm.N1)) return 1;
  int it = 0;
  if (c->use_cgp) {                               // persistent cooperative path (small single-rank problems)
    NLK_CUDA(cudaMemsetAsync(c->cg_x, 0, c->dm.N1 * sizeof(double), c->st));
    double* rhs[1] = {c->wk[7]}; const double* mk[1] = {c->dm.mask[comp]}; double* sol[1] = {c->cg_x};
    if (helmholtz_solve_multi(c, 1, rhs, h1, h2, mk, tol, sol)) return 1;
    if (c->use_cgp) { NLK_CUDA(cudaMemcpyAsync(&it, c->d_cg_iters, sizeof(int), cudaMemcpyDeviceToHost, c->st)); NLK_CUDA(cudaStreamSynchronize(c->st)); }
  } else if (helmholtz_solve(c, c->wk[7], h1, h2, c->dm.mask[comp], tol, c->cg_x, &it)) return 1;
  if (iters) *iters = it;
  return down(c, x, c->cg_x, c->dm.N1);
}
int nlk_test_pressure(nlk_ctx* c, const double* rhs, double tol, double* x, int32_t* iters) {
  if (up(c, c->pw[4], rhs, c->dm.N2)) return 1;
  int it = 0;
  if (pressure_solve(c, c->pw[4], tol, c->pw[3], &it)) return 1;
  if (iters) *iters = it;
  return down(c, x, c->pw[3], c->dm.N2);
}
int nlk_test_precond(nlk_ctx* c, const double* r, double* z) {
  if (up(c, c->pw[3], r, c->dm.N2)) return 1;
  if (apply_precond(c, c->pw[3], c->pw[4], nullptr)) return 1;
  return down(c, z, c->pw[4], c->dm.N2);
}
int nlk_test_cfl(nlk_ctx* c, const double* ux, const double* uy, const double* uz, double dt, double* cfl) {
  const double* u[3] = {ux, uy, uz};
  for (int k = 0; k < c->dm.ndim; ++k) if (up(c, c->wk[k], u[k], c->dm.N1)) return 1;
  launch_cfl(c->dm, CPtr3{{c->wk[0], c->wk[1], c->wk[2]}}, c->d_red, c->red, c->st);
  if (ctx_allreduce(c, c->d_red, 1, true)) return 1;
  if (ctx_read_scalars(c, 1)) return 1;
  *cfl = dt * c->h_red[0]; return 0;
}

int nlk_bench_kernel(nlk_ctx* c, int32_t which, int32_t nrep, double* ms_per_launch, double* algo_bytes) {
  const DevMesh& dm = c->dm; const int d = dm.ndim;
  EvPair ev; cudaEvent_t e0 = ev.a, e1 = ev.b;
  auto run = [&]() -> int {
    switch (which) {
      case 0: launch_axhelm(dm, c->wk[0], c->wk[1], 1.0, 1.0, c->st); break;
      case 1: return ctx_gs(c, Ptr3{{c->wk[0], nullptr, nullptr}}, 1);
      case 2: return apply_E(c, c->pw[3], c->pw[4], nullptr);
      case 3: launch_convect(dm, CPtr4{{c->wk[3], c->wk[4], c->wk[5], nullptr}}, d, CPtr3{{c->wk[0], c->wk[1], c->wk[2]}}, Ptr4{{c->bf[0], c->bf[1], c->bf[2], nullptr}}, 1.0, 0, c->st); break;
      case 4: return apply_precond(c, c->pw[3], c->pw[4], nullptr);
      case 5: { CPtr4 A{{c->wk[0], c->wk[1], c->wk[2], nullptr}}; launch_dot(dm.N1, A, A, d, dm.bm1, c->d_red, c->red, c->st); break; }
      case 6: if (!c->coarse_sparse) { set_error("bench 6: the context has no sparse coarse level"); return 1; } return coarse_solve_sparse(c, c->crs_r, c->crs_y);
      case 7: if (!c->have_schwarz) { set_error("bench 7: no Schwarz level"); return 1; }      // Schwarz branch alone
              if (c->swf) return swf_apply(c, c->pw[3], nullptr, nullptr, c->pw[4]);
              launch_schwarz_embed(dm, c->pw[3], nullptr, c->sw_w, c->st); if (ctx_gs(c, Ptr3{{c->sw_w, nullptr, nullptr}}, 1)) return 1;
              launch_schwarz_fdm(dm, c->sw_w, c->sw_z, c->sw_t, c->st); if (ctx_gs(c, Ptr3{{c->sw_t, nullptr, nullptr}}, 1)) return 1;
              launch_schwarz_gather(dm, c->sw_z, c->sw_t, c->pw[4], c->st); break;
      case 8: { int slot; if (cg_weights(c, dm.mask[0], 0.02, 180.0, &slot)) return 1;
                launch_axhelm_cg(dm, c->cg_p, c->cg_r, c->cg_w, c->cg_hd[slot], 0.02, 180.0, c->d_sc, c->cg_pap_partial, c->cg_pap_counter, 1, c->st); break; }
      case 9: { int slot; if (cg_weights(c, dm.mask[0], 0.02, 180.0, &slot)) return 1;
                launch_cg_update_reduce(dm, c->cg_x, c->cg_r, c->cg_p, c->cg_w, c->cg_wa[slot], c->cg_wb[slot], c->d_sc, c->red, 0, 1, c->st); break; }
      case 10: launch_opgradt(dm, c->pw[3], Ptr3{{c->wk[0], c->wk[1], c->wk[2]}}, c->st); break;
      case 11: launch_opdiv_fused(dm, CPtr3{{c->wk[0], c->wk[1], c->wk[2]}}, c->pw[4], 1.0, dm.binvm1, nullptr, c->st); break;
      case 12: return ctx_gs(c, Ptr3{{c->wk[0], c->wk[1], c->wk[2]}}, d);
      default: set_error("unknown bench kernel"); return 1;
    }
    return 0;
  };
  if (which == 8 || which == 9) {      // the PCG kernels read their scalars from the device: fresh state, deferred finalisation (no early exit)
    launch_cg_init(dm, c->d_sc, 0.0, 1 << 30, c->st);
    NLK_CUDA(cudaMemcpyAsync(c->cg_r, c->wk[0], dm.N1 * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    NLK_CUDA(cudaMemsetAsync(c->cg_p, 0, dm.N1 * sizeof(double), c->st));
  }
  for (int i = 0; i < 3; ++i) if (run()) return 1;
  cudaEventRecord(e0, c->st);
  for (int i = 0; i < nrep; ++i) if (run()) return 1;
  cudaEventRecord(e1, c->st);
  NLK_CUDA(cudaStreamSynchronize(c->st));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_launch = ms / nrep;
  double N1 = (double)dm.N1, N2 = (double)dm.N2;
  double bytes = 0;
  switch (which) {
    case 0: bytes = (2 + dm.ng + 1) * 8.0 * N1; break;                      // u, w, g-factors, bm1
    case 1: bytes = 16.0 * N1 * 0 + (double)c->mesh->hm.gs_idx.size() * (16.0 + 4.0); break;
    case 2: bytes = 2 * 8.0 * N2 + 2.0 * d * d * 8.0 * N2 + d * (4 * 8.0) * N1; break;
    case 3: bytes = (3.0 * d + d * d) * 8.0 * N1; break;
    case 4: bytes = 2 * 8.0 * N2 + 6 * 8.0 * N1; break;
    case 5: bytes = (d + 1) * 8.0 * N1; break;
    case 6: bytes = (double)c->crs_iters * (12.0 * (double)c->crs_nnz + 9 * 8.0 * (double)dm.nvert); break;   // CSR values + indices per SpMV, vectors
    case 7: bytes = 2 * 8.0 * N2 + 6 * 8.0 * N1; break;                       // (same figure for the chain and the fused path: the chain's work arrays)
    case 8: bytes = (5 + dm.ng + 1) * 8.0 * N1; break;                       // r, hd, p in; p, w out; g-factors; bm1
    case 9: bytes = 8 * 8.0 * N1; break;                                     // x, r, p, w, wa, wb in; x, r out
    case 10: bytes = d * 8.0 * N1 + (1 + d * d) * 8.0 * N2; break;
    case 11: bytes = 2.0 * d * 8.0 * N1 + (1 + d * d) * 8.0 * N2; break;     // u_c and binvm1*mask per component in; metrics; p out
    case 12: bytes = d * (double)c->mesh->hm.gs_idx.size() * 16.0 + (double)c->mesh->hm.gs_idx.size() * 4.0; break;
  }
  if (algo_bytes) *algo_bytes = bytes;
  return 0;
}

}  // extern "C"
