// Host mesh setup: metrics, mass, stiffness factors, mesh-2 / dealias metrics, direct-stiffness numbering
// from .ma2 vertex ids, Dirichlet masks, partition rule, gather-scatter plans (local CSR + neighbour lists).
// Restates Nek5000 geom1/geom2 (coef.f), map12, set_dealias_rx (convect.f), set_vert/setvert2d/3d (navier8.f),
// bcmask (bdry.f), assign_gllnid -- all un-vendored upstream.
// In-tree consumers: src/vectors/real_vectors.f90:100-113 (opdssum/vmult/masks), :217-224 (bm1).
#include "nlk_host.hpp"
#include "../../include/nlk.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

namespace nlk {

// ---- small tensor helpers on one element (x fastest) -------------------------------------------------
// apply M (mo x mi) along direction dir of u with dims (d0=x, d1=y, d2=z)
static void apply_dir(const double* M, int mo, int mi, const double* u, double* out, int dir, int n0, int n1, int n2) {
  int dims[3] = {n0, n1, n2};
  int od[3] = {n0, n1, n2}; od[dir] = mo;
  for (int k = 0; k < od[2]; ++k) for (int j = 0; j < od[1]; ++j) for (int i = 0; i < od[0]; ++i) {
    int idx[3] = {i, j, k};
    double s = 0;
    for (int l = 0; l < mi; ++l) {
      int id[3] = {i, j, k}; id[dir] = l;
      s += M[idx[dir] * mi + l] * u[(id[2] * dims[1] + id[1]) * dims[0] + id[0]];
    }
    out[(k * od[1] + j) * od[0] + i] = s;
  }
}

// (M (x) M [(x) M]) u : n^d -> mo^d
static void tensor_apply(const double* M, int mo, int mi, const double* u, double* out, int ndim, std::vector<double>& t1, std::vector<double>& t2) {
  if (ndim == 2) {
    t1.resize((size_t)mo * mi);
    apply_dir(M, mo, mi, u, t1.data(), 0, mi, mi, 1);
    apply_dir(M, mo, mi, t1.data(), out, 1, mo, mi, 1);
  } else {
    t1.resize((size_t)mo * mi * mi); t2.resize((size_t)mo * mo * mi);
    apply_dir(M, mo, mi, u, t1.data(), 0, mi, mi, mi);
    apply_dir(M, mo, mi, t1.data(), t2.data(), 1, mo, mi, mi);
    apply_dir(M, mo, mi, t2.data(), out, 2, mo, mo, mi);
  }
}

static inline bool is_dirichlet_v(const char* c) {
  return (c[0] == 'v' || c[0] == 'V' || c[0] == 'W' || c[0] == 'w') && (c[1] == ' ' || c[1] == 'l' || c[1] == 'L' || c[1] == 0);
}
// 'SYM': symmetry plane whose normal is the element's own reference axis of that face (structured, unrotated meshes);
// 'SYx' / 'SYy' / 'SYz': symmetry plane with PHYSICAL normal x / y / z -- what a front end writes for unstructured meshes
// whose elements are rotated (api.resolve_sym does it from the coordinates).  Returns the velocity component to mask, -1 if
// the code is not a symmetry code.  (Nek's general 'SYM' rotates into face-normal coordinates; only axis-aligned planes
// occur in the reference's configs.)
static inline int sym_axis(const char* c, int ref_axis) {
  if (c[0] != 'S' || c[1] != 'Y') return -1;
  if (c[2] == 'M') return ref_axis;
  if (c[2] == 'x' || c[2] == 'X') return 0;
  if (c[2] == 'y' || c[2] == 'Y') return 1;
  if (c[2] == 'z' || c[2] == 'Z') return 2;
  return -1;
}
static inline bool is_outflow(const char* c) { return (c[0] == 'O' || c[0] == 'o'); }
static inline bool is_dirichlet_t(const char* c) { return (c[0] == 't' || c[0] == 'T') && (c[1] == ' ' || c[1] == 0); }

// preprocessor face f (0-based) -> (normal axis, side)
static const int FACE_AXIS[6] = {1, 0, 1, 0, 2, 2};
static const int FACE_SIDE[6] = {0, 1, 1, 0, 0, 1};

struct Key2 { int64_t a, b; int64_t e; int32_t le; };
struct Key4 { int64_t v[4]; int64_t e; int32_t lf; };

int build_mesh(HostMesh& hm, int ndim, int lx1, int lxd, int64_t nelg, int64_t nel, const double* xm1, const double* ym1,
               const double* zm1, const int64_t* vertex_all, const char* cbc_v, const char* cbc_t,
               const int32_t* gllnid, int rank, int nranks) {
  if (ndim != 2 && ndim != 3) { set_error("ndim must be 2 or 3"); return 1; }
  if (lx1 < 3 || lx1 > 16) { set_error("lx1 out of range [3,16]"); return 1; }
  if (lxd < lx1) { set_error("lxd must be >= lx1"); return 1; }
  if (nranks > 64) { set_error("at most 64 ranks"); return 1; }
  hm.ndim = ndim; hm.n = lx1; hm.m = lxd; hm.q = lx1 - 2;
  hm.Eg = nelg; hm.rank = rank; hm.nranks = nranks;
  const int n = lx1, m = lxd, q = lx1 - 2, d = ndim;
  const int nz = d == 3 ? n : 1;
  hm.np1 = d == 3 ? n * n * n : n * n;
  hm.np2 = d == 3 ? q * q * q : q * q;
  hm.npd = d == 3 ? m * m * m : m * m;
  hm.nv = 1 << d; hm.nfaces = 2 * d;
  const int np1 = hm.np1, np2 = hm.np2, npd = hm.npd, nv = hm.nv, nf = hm.nfaces;
  hm.b.build(lx1, lxd);
  const Basis& b = hm.b;
  // ---- local elements
  hm.lglel.clear();
  for (int64_t e = 0; e < nelg; ++e) if (!gllnid || gllnid[e] == rank) hm.lglel.push_back(e);
  if ((int64_t)hm.lglel.size() != nel) { set_error("nel does not match the number of elements owned by this rank"); return 1; }
  hm.E = nel;
  const int64_t E = nel;
  const double* xin[3] = {xm1, ym1, zm1};
  for (int c = 0; c < d; ++c) { if (!xin[c]) { set_error("missing coordinate array"); return 1; } hm.xyz[c].assign(xin[c], xin[c] + (size_t)E * np1); }
  hm.vertex_all.assign(vertex_all, vertex_all + (size_t)nelg * nv);
  hm.vertex_local.resize((size_t)E * nv);
  for (int64_t e = 0; e < E; ++e) for (int c = 0; c < nv; ++c) hm.vertex_local[e * nv + c] = vertex_all[hm.lglel[e] * nv + c];
  hm.cbc_v.resize((size_t)E * nf); hm.cbc_t.resize((size_t)E * nf); hm.has_tbc = cbc_t != nullptr;
  for (int64_t e = 0; e < E; ++e) for (int f = 0; f < nf; ++f) {
    const char* s = cbc_v + ((size_t)hm.lglel[e] * nf + f) * 3;
    hm.cbc_v[e * nf + f] = {s[0], s[1], s[2]};
    if (cbc_t) { const char* t = cbc_t + ((size_t)hm.lglel[e] * nf + f) * 3; hm.cbc_t[e * nf + f] = {t[0], t[1], t[2]}; }
    else hm.cbc_t[e * nf + f] = {'E', ' ', ' '};
  }
  hm.has_outflow = false;
  for (int64_t e = 0; e < nelg; ++e) for (int f = 0; f < nf; ++f) if (is_outflow(cbc_v + ((size_t)e * nf + f) * 3)) hm.has_outflow = true;
  // symmetry planes must be coordinate planes of the masked component: check every local symmetry face geometrically
  for (int64_t e = 0; e < E; ++e) for (int f = 0; f < nf; ++f) {
    const char* s = cbc_v + ((size_t)hm.lglel[e] * nf + f) * 3;
    const int ax = FACE_AXIS[f], side = FACE_SIDE[f], sa = sym_axis(s, ax);
    if (sa < 0) continue;
    if (sa >= d) { set_error("symmetry code names an axis outside the mesh dimension"); return 1; }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    const int nz = d == 3 ? n : 1;
    for (int k = 0; k < nz; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
      const size_t g = (size_t)e * np1 + ((size_t)k * n + j) * n + i;
      const int idx = ax == 0 ? i : (ax == 1 ? j : k);
      if (idx != (side ? n - 1 : 0)) continue;
      for (int c = 0; c < d; ++c) { lo[c] = std::min(lo[c], hm.xyz[c][g]); hi[c] = std::max(hi[c], hm.xyz[c][g]); }
    }
    double ext = 0; for (int c = 0; c < d; ++c) ext = std::max(ext, hi[c] - lo[c]);
    if (hi[sa] - lo[sa] > 1e-8 * std::max(ext, 1e-300)) {
      set_error("symmetry face is not a coordinate plane of the masked component: write 'SYx'/'SYy'/'SYz' (physical normal) in the "
                "boundary codes (neklab_b200.api.resolve_sym) -- plain 'SYM' means the element's own reference axis");
      return 1;
    }
  }

  // ---- geometry
  const size_t N1 = (size_t)E * np1, N2 = (size_t)E * np2, Nd = (size_t)E * npd;
  hm.jac.assign(N1, 0); hm.bm1.assign(N1, 0); hm.bm2.assign(N2, 0); hm.bm2inv.assign(N2, 0);
  for (int k = 0; k < d * d; ++k) { hm.rx[k].assign(N1, 0); hm.rxw2[k].assign(N2, 0); hm.rxd[k].assign(Nd, 0); }
  const int ng = d == 3 ? 6 : 3;
  for (int k = 0; k < 6; ++k) hm.G[k].clear();
  if (d == 3) for (int k = 0; k < 6; ++k) hm.G[k].assign(N1, 0);
  else { hm.G[0].assign(N1, 0); hm.G[1].assign(N1, 0); hm.G[3].assign(N1, 0); }
  (void)ng;
  int bad_jac = 0;
  double vol1 = 0, vol2 = 0;
#pragma omp parallel
  {
    std::vector<double> xr(9 * (size_t)np1), t1, t2, tmp2(np2), tmpd(npd), jac2(np2);
    int bad_local = 0; double v1 = 0, v2 = 0;
#pragma omp for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
      // xr[(c*3+k)] = d x_c / d r_k
      for (int c = 0; c < d; ++c) for (int k = 0; k < d; ++k)
        apply_dir(b.D.data(), n, n, hm.xyz[c].data() + e * np1, xr.data() + (size_t)(c * 3 + k) * np1, k, n, n, nz);
      for (int p = 0; p < np1; ++p) {
        double J, r[9];
        if (d == 2) {
          double x_r = xr[(0 * 3 + 0) * (size_t)np1 + p], x_s = xr[(0 * 3 + 1) * (size_t)np1 + p];
          double y_r = xr[(1 * 3 + 0) * (size_t)np1 + p], y_s = xr[(1 * 3 + 1) * (size_t)np1 + p];
          J = x_r * y_s - x_s * y_r;
          r[0] = y_s; r[1] = -x_s; r[2] = -y_r; r[3] = x_r;          // rx ry ; sx sy
        } else {
          auto X = [&](int c, int k) { return xr[(size_t)(c * 3 + k) * np1 + p]; };
          double x_r = X(0, 0), x_s = X(0, 1), x_t = X(0, 2), y_r = X(1, 0), y_s = X(1, 1), y_t = X(1, 2), z_r = X(2, 0), z_s = X(2, 1), z_t = X(2, 2);
          J = x_r * (y_s * z_t - y_t * z_s) - x_s * (y_r * z_t - y_t * z_r) + x_t * (y_r * z_s - y_s * z_r);
          r[0] = y_s * z_t - y_t * z_s; r[1] = x_t * z_s - x_s * z_t; r[2] = x_s * y_t - x_t * y_s;
          r[3] = y_t * z_r - y_r * z_t; r[4] = x_r * z_t - x_t * z_r; r[5] = x_t * y_r - x_r * y_t;
          r[6] = y_r * z_s - y_s * z_r; r[7] = x_s * z_r - x_r * z_s; r[8] = x_r * y_s - x_s * y_r;
        }
        if (!(J > 0)) bad_local = 1;
        int i = p % n, j = (p / n) % n, k = p / (n * n);
        double W = b.w1[i] * b.w1[j] * (d == 3 ? b.w1[k] : 1.0);
        size_t g = (size_t)e * np1 + p;
        hm.jac[g] = J; hm.bm1[g] = J * W; v1 += J * W;
        for (int kk = 0; kk < d * d; ++kk) hm.rx[kk][g] = r[kk];
        auto dotr = [&](int a, int bb) { double s = 0; for (int c = 0; c < d; ++c) s += r[a * d + c] * r[bb * d + c]; return s * W / J; };
        hm.G[0][g] = dotr(0, 0); hm.G[1][g] = dotr(1, 1); hm.G[3][g] = dotr(0, 1);
        if (d == 3) { hm.G[2][g] = dotr(2, 2); hm.G[4][g] = dotr(0, 2); hm.G[5][g] = dotr(1, 2); }
      }
      // mesh 2 and dealias metrics
      tensor_apply(b.I12.data(), q, n, hm.jac.data() + e * np1, jac2.data(), d, t1, t2);
      for (int p = 0; p < np2; ++p) {
        int i = p % q, j = (p / q) % q, k = p / (q * q);
        double W = b.w2[i] * b.w2[j] * (d == 3 ? b.w2[k] : 1.0);
        hm.bm2[e * np2 + p] = jac2[p] * W; hm.bm2inv[e * np2 + p] = 1.0 / (jac2[p] * W); v2 += jac2[p] * W;
      }
      for (int kk = 0; kk < d * d; ++kk) {
        tensor_apply(b.I12.data(), q, n, hm.rx[kk].data() + e * np1, tmp2.data(), d, t1, t2);
        for (int p = 0; p < np2; ++p) {
          int i = p % q, j = (p / q) % q, k = p / (q * q);
          hm.rxw2[kk][e * np2 + p] = tmp2[p] * b.w2[i] * b.w2[j] * (d == 3 ? b.w2[k] : 1.0);
        }
        tensor_apply(b.I1d.data(), m, n, hm.rx[kk].data() + e * np1, tmpd.data(), d, t1, t2);
        for (int p = 0; p < npd; ++p) {
          int i = p % m, j = (p / m) % m, k = p / (m * m);
          hm.rxd[kk][e * npd + p] = tmpd[p] * b.wd[i] * b.wd[j] * (d == 3 ? b.wd[k] : 1.0);
        }
      }
    }
#pragma omp critical
    { bad_jac |= bad_local; vol1 += v1; vol2 += v2; }
  }
  if (bad_jac) { set_error("non-positive Jacobian (check corner ordering / coordinates)"); return 1; }
  hm.volvm1 = vol1; hm.volvm2 = vol2;

  // ---- global entity tables (vertices given; edges, faces by sorted vertex tuples), identical ids to the oracle
  int64_t nvert = 0;
  for (int64_t i = 0; i < nelg * nv; ++i) nvert = std::max(nvert, vertex_all[i]);
  hm.nvert = nvert;
  struct EdgeDef { int a, b, ax; };
  std::vector<EdgeDef> edefs;
  for (int c = 0; c < nv; ++c) {
    int i = c & 1, j = (c >> 1) & 1, k = (c >> 2) & 1;
    if (i == 0) edefs.push_back({c, c | 1, 0});
    if (j == 0) edefs.push_back({c, c | 2, 1});
    if (d == 3 && k == 0) edefs.push_back({c, c | 4, 2});
  }
  const int ne_loc = (int)edefs.size();
  std::vector<Key2> ek((size_t)nelg * ne_loc);
  for (int64_t e = 0; e < nelg; ++e) for (int le = 0; le < ne_loc; ++le) {
    int64_t va = vertex_all[e * nv + edefs[le].a], vb = vertex_all[e * nv + edefs[le].b];
    ek[e * ne_loc + le] = {std::min(va, vb), std::max(va, vb), e, le};
  }
  std::sort(ek.begin(), ek.end(), [](const Key2& x, const Key2& y) { return x.a != y.a ? x.a < y.a : (x.b != y.b ? x.b < y.b : (x.e != y.e ? x.e < y.e : x.le < y.le)); });
  std::vector<int64_t> edge_id((size_t)nelg * ne_loc);
  int64_t nedge = 0;
  for (size_t i = 0; i < ek.size(); ++i) {
    if (i > 0 && (ek[i].a != ek[i - 1].a || ek[i].b != ek[i - 1].b)) ++nedge;
    edge_id[ek[i].e * ne_loc + ek[i].le] = nedge;
  }
  if (!ek.empty()) ++nedge;
  // faces (3-D): fixed axis ax, side; in-face axes (a,b) = remaining axes ascending; corners [c00,c10,c01,c11]
  struct FaceDef { int ax, side, cs[4]; };
  std::vector<FaceDef> fdefs;
  std::vector<int64_t> face_id;
  int64_t nface = 0;
  if (d == 3) {
    for (int ax = 0; ax < 3; ++ax) for (int side = 0; side < 2; ++side) {
      FaceDef f; f.ax = ax; f.side = side;
      int oth[2], t = 0; for (int dd = 0; dd < 3; ++dd) if (dd != ax) oth[t++] = dd;
      int ci = 0;
      for (int bb = 0; bb < 2; ++bb) for (int aa = 0; aa < 2; ++aa) {
        int ijk[3]; ijk[ax] = side; ijk[oth[0]] = aa; ijk[oth[1]] = bb;
        f.cs[ci++] = ijk[0] | (ijk[1] << 1) | (ijk[2] << 2);
      }
      fdefs.push_back(f);
    }
    std::vector<Key4> fk((size_t)nelg * 6);
    for (int64_t e = 0; e < nelg; ++e) for (int lf = 0; lf < 6; ++lf) {
      Key4 k; for (int t = 0; t < 4; ++t) k.v[t] = vertex_all[e * nv + fdefs[lf].cs[t]];
      std::sort(k.v, k.v + 4); k.e = e; k.lf = lf; fk[e * 6 + lf] = k;
    }
    auto less4 = [](const Key4& x, const Key4& y) { for (int t = 0; t < 4; ++t) if (x.v[t] != y.v[t]) return x.v[t] < y.v[t]; return x.e != y.e ? x.e < y.e : x.lf < y.lf; };
    std::sort(fk.begin(), fk.end(), less4);
    face_id.resize((size_t)nelg * 6);
    for (size_t i = 0; i < fk.size(); ++i) {
      if (i > 0 && std::memcmp(fk[i].v, fk[i - 1].v, sizeof(int64_t) * 4) != 0) ++nface;
      face_id[fk[i].e * 6 + fk[i].lf] = nface;
    }
    if (!fk.empty()) ++nface;
  }
  // rank masks and Dirichlet bits per entity (global loops)
  std::vector<uint64_t> vrank(nvert + 1, 0), erank(nedge, 0), frank(nface, 0);
  std::vector<uint8_t> vbits(nvert + 1, 0), ebits(nedge, 0), fbits(nface, 0);
  for (int64_t e = 0; e < nelg; ++e) {
    uint64_t rb = 1ull << (gllnid ? gllnid[e] : 0);
    for (int c = 0; c < nv; ++c) vrank[vertex_all[e * nv + c]] |= rb;
    for (int le = 0; le < ne_loc; ++le) erank[edge_id[e * ne_loc + le]] |= rb;
    if (d == 3) for (int lf = 0; lf < 6; ++lf) frank[face_id[e * 6 + lf]] |= rb;
    for (int f = 0; f < nf; ++f) {
      const char* cv = cbc_v + ((size_t)e * nf + f) * 3;
      uint8_t bits = 0;
      int ax = FACE_AXIS[f], side = FACE_SIDE[f];
      if (is_dirichlet_v(cv)) bits |= 7;
      { const int sa = sym_axis(cv, ax); if (sa >= 0 && sa < d) bits |= (uint8_t)(1 << sa); }
      if (cbc_t && is_dirichlet_t(cbc_t + ((size_t)e * nf + f) * 3)) bits |= 8;
      if (!bits) continue;
      for (int c = 0; c < nv; ++c) if (((c >> ax) & 1) == side) vbits[vertex_all[e * nv + c]] |= bits;
      for (int le = 0; le < ne_loc; ++le) {
        int a = edefs[le].a, bb = edefs[le].b;
        if (((a >> ax) & 1) == side && ((bb >> ax) & 1) == side) ebits[edge_id[e * ne_loc + le]] |= bits;
      }
      if (d == 3) for (int lf = 0; lf < 6; ++lf) if (fdefs[lf].ax == ax && fdefs[lf].side == side) fbits[face_id[e * 6 + lf]] |= bits;
    }
  }
  // ---- local numbering, masks, interface sharer masks
  hm.glo.assign(N1, 0);
  for (int c = 0; c < 3; ++c) hm.vmask[c].assign(c < d ? N1 : 0, 1.0);
  hm.tmask.assign(N1, 1.0);
  std::vector<uint64_t> sharers(N1, 0);        // rank mask for surface nodes (0 for interiors)
  const int ni = n - 2;
  const int64_t base_edge = nvert, base_face = nvert + nedge * ni, base_int = base_face + nface * (int64_t)ni * ni;
  const int64_t nint = d == 3 ? (int64_t)ni * ni * ni : (int64_t)ni * ni;
  auto setnode = [&](int64_t e, int i, int j, int k, int64_t gid, uint64_t rmask, uint8_t bits) {
    size_t g = (size_t)e * np1 + ((size_t)k * n + j) * n + i;
    hm.glo[g] = gid; sharers[g] = rmask;
    for (int c = 0; c < d; ++c) if (bits & (1 << c)) hm.vmask[c][g] = 0.0;
    if (bits & 8) hm.tmask[g] = 0.0;
  };
  for (int64_t e = 0; e < E; ++e) {
    const int64_t eg = hm.lglel[e];
    const int64_t* v = vertex_all + eg * nv;
    for (int c = 0; c < nv; ++c) {
      int i = (c & 1) * (n - 1), j = ((c >> 1) & 1) * (n - 1), k = ((c >> 2) & 1) * (nz - 1);
      setnode(e, i, j, k, v[c], vrank[v[c]], vbits[v[c]]);
    }
    for (int le = 0; le < ne_loc; ++le) {
      const EdgeDef& ed = edefs[le];
      int64_t id = edge_id[eg * ne_loc + le];
      bool fwd = v[ed.a] < v[ed.b];
      int ci = (ed.a & 1) * (n - 1), cj = ((ed.a >> 1) & 1) * (n - 1), ck = ((ed.a >> 2) & 1) * (nz - 1);
      for (int pos = 1; pos < n - 1; ++pos) {
        int p = fwd ? pos - 1 : (n - 1 - pos) - 1;
        int64_t gid = base_edge + id * ni + p + 1;
        int i = ci, j = cj, k = ck;
        if (ed.ax == 0) i = pos; else if (ed.ax == 1) j = pos; else k = pos;
        setnode(e, i, j, k, gid, erank[id], ebits[id]);
      }
    }
    if (d == 3) for (int lf = 0; lf < 6; ++lf) {
      const FaceDef& fd = fdefs[lf];
      int64_t id = face_id[eg * 6 + lf];
      int64_t fv[4]; for (int t = 0; t < 4; ++t) fv[t] = v[fd.cs[t]];
      int cmin = 0; for (int t = 1; t < 4; ++t) if (fv[t] < fv[cmin]) cmin = t;
      int amin = cmin & 1, bmin = (cmin >> 1) & 1;
      int64_t na = fv[(1 - amin) | (bmin << 1)], nb = fv[amin | ((1 - bmin) << 1)];
      bool swap = nb < na;
      int oth[2], t = 0; for (int dd = 0; dd < 3; ++dd) if (dd != fd.ax) oth[t++] = dd;
      for (int B = 1; B < n - 1; ++B) for (int A = 1; A < n - 1; ++A) {
        int ia = amin == 0 ? A - 1 : (n - 1 - A) - 1;
        int ib = bmin == 0 ? B - 1 : (n - 1 - B) - 1;
        int p = swap ? ib + ni * ia : ia + ni * ib;
        int64_t gid = base_face + id * (int64_t)ni * ni + p + 1;
        int ijk[3]; ijk[fd.ax] = fd.side * (n - 1); ijk[oth[0]] = A; ijk[oth[1]] = B;
        setnode(e, ijk[0], ijk[1], ijk[2], gid, frank[id], fbits[id]);
      }
    }
    // interiors
    if (d == 3) {
      for (int k = 1; k < n - 1; ++k) for (int j = 1; j < n - 1; ++j) for (int i = 1; i < n - 1; ++i)
        setnode(e, i, j, k, base_int + eg * nint + ((int64_t)(k - 1) * ni + (j - 1)) * ni + (i - 1) + 1, 0, 0);
    } else {
      for (int j = 1; j < n - 1; ++j) for (int i = 1; i < n - 1; ++i)
        setnode(e, i, j, 0, base_int + eg * nint + (int64_t)(j - 1) * ni + (i - 1) + 1, 0, 0);
    }
  }
  // ---- local gather-scatter CSR over surface nodes
  std::vector<std::pair<int64_t, int32_t>> surf;
  surf.reserve((size_t)E * (np1 - nint));
  for (int64_t e = 0; e < E; ++e) for (int p = 0; p < np1; ++p) {
    int i = p % n, j = (p / n) % n, k = p / (n * n);
    bool interior = i > 0 && i < n - 1 && j > 0 && j < n - 1 && (d == 2 || (k > 0 && k < n - 1));
    if (!interior) surf.push_back({hm.glo[e * np1 + p], (int32_t)(e * np1 + p)});
  }
  if ((size_t)E * np1 > (size_t)INT32_MAX) { set_error("local mesh too large for int32 indices"); return 1; }
  std::sort(surf.begin(), surf.end());
  hm.gs_off.assign(1, 0); hm.gs_idx.clear(); hm.gs_first.clear();
  int64_t ndistinct_surf = 0;
  std::map<int, std::vector<std::pair<int64_t, int32_t>>> nb_first;      // neighbour rank -> (gid, rep)
  std::vector<std::pair<int64_t, std::pair<size_t, size_t>>> if_groups;   // interface gid -> [begin,end) in surf
  for (size_t s = 0; s < surf.size();) {
    size_t t = s; while (t < surf.size() && surf[t].first == surf[s].first) ++t;
    ++ndistinct_surf;
    if (t - s >= 2) {
      for (size_t u = s; u < t; ++u) hm.gs_idx.push_back(surf[u].second);
      hm.gs_off.push_back((int32_t)hm.gs_idx.size());
      hm.gs_first.push_back(surf[s].second);
    }
    uint64_t rm = sharers[surf[s].second] & ~(1ull << rank);
    if (rm) {
      if_groups.push_back({surf[s].first, {s, t}});
      for (int r = 0; r < nranks; ++r) if (rm & (1ull << r)) nb_first[r].push_back({surf[s].first, surf[s].second});
    }
    s = t;
  }
  hm.nglob_local = ndistinct_surf + E * nint;
  // Order the shared groups by the local address of their first copy (copies inside a group stay in ascending local order, so
  // every sum is taken in the same order as before): consecutive threads of k_gs then walk the owner element's surface in
  // memory order, and the partner copies sit on the neighbours' matching faces -- each 32-byte sector of a face is touched by
  // one or two adjacent warps instead of by warps scattered over the global-id order (r01 ncu: 6.8x the useful DRAM bytes).
  {
    const size_t ng = hm.gs_off.size() - 1;
    std::vector<int32_t> ord(ng); for (size_t i = 0; i < ng; ++i) ord[i] = (int32_t)i;
    std::sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return hm.gs_idx[hm.gs_off[x]] < hm.gs_idx[hm.gs_off[y]]; });
    std::vector<int32_t> noff(1, 0), nidx; nidx.reserve(hm.gs_idx.size()); std::vector<int32_t> nfirst; nfirst.reserve(ng);
    for (size_t i = 0; i < ng; ++i) {
      const int32_t g = ord[i];
      for (int32_t u = hm.gs_off[g]; u < hm.gs_off[g + 1]; ++u) nidx.push_back(hm.gs_idx[u]);
      noff.push_back((int32_t)nidx.size()); nfirst.push_back(hm.gs_first[g]);
    }
    hm.gs_off.swap(noff); hm.gs_idx.swap(nidx); hm.gs_first.swap(nfirst);
  }
  hm.neigh.clear();
  for (auto& kv : nb_first) {
    Neighbor nb; nb.rank = kv.first;
    for (auto& pr : kv.second) { nb.gids.push_back(pr.first); nb.rep.push_back(pr.second); }
    hm.neigh.push_back(std::move(nb));
  }
  hm.if_gids.clear(); hm.if_off.assign(1, 0); hm.if_idx.clear();
  for (auto& g : if_groups) {
    hm.if_gids.push_back(g.first);
    for (size_t u = g.second.first; u < g.second.second; ++u) hm.if_idx.push_back(surf[u].second);
    hm.if_off.push_back((int32_t)hm.if_idx.size());
  }
  // ---- vertex incidence CSR (coarse-grid restriction)
  {
    std::vector<std::pair<int64_t, int32_t>> inc((size_t)E * nv);
    for (int64_t e = 0; e < E; ++e) for (int c = 0; c < nv; ++c) inc[e * nv + c] = {hm.vertex_local[e * nv + c], (int32_t)(e * nv + c)};
    std::sort(inc.begin(), inc.end());
    hm.vert_off.assign(nvert + 1, 0); hm.vert_ec.resize(inc.size());
    for (size_t i = 0; i < inc.size(); ++i) { hm.vert_off[inc[i].first]++; hm.vert_ec[i] = inc[i].second; }   // ids are 1-based: count at [vid]
    // exclusive scan: vert_off[v-1]..vert_off[v]
    int32_t run = 0;
    for (int64_t vtx = 1; vtx <= nvert; ++vtx) { int32_t c = hm.vert_off[vtx]; hm.vert_off[vtx - 1] = run; run += c; }
    hm.vert_off[nvert] = run;
  }
  // single-rank conveniences: binvm1, vmult on the host (multi-rank: recomputed on device with the exchange)
  hm.binvm1.assign(N1, 0); hm.vmult.assign(N1, 0);
  for (size_t g = 0; g < N1; ++g) { hm.binvm1[g] = hm.bm1[g]; hm.vmult[g] = 1.0; }
  for (size_t gi = 0; gi + 1 < hm.gs_off.size(); ++gi) {
    double sb = 0, sm = 0;
    for (int32_t u = hm.gs_off[gi]; u < hm.gs_off[gi + 1]; ++u) { sb += hm.bm1[hm.gs_idx[u]]; sm += 1.0; }
    for (int32_t u = hm.gs_off[gi]; u < hm.gs_off[gi + 1]; ++u) { hm.binvm1[hm.gs_idx[u]] = sb; hm.vmult[hm.gs_idx[u]] = sm; }
  }
  for (size_t g = 0; g < N1; ++g) { hm.binvm1[g] = 1.0 / hm.binvm1[g]; hm.vmult[g] = 1.0 / hm.vmult[g]; }
  return 0;
}

}  // namespace nlk

// ======================================================================= C-ABI (host-only part)
using namespace nlk;
struct nlk_mesh { HostMesh hm; };

extern "C" {

const char* nlk_last_error(void) { return get_error(); }
int nlk_version(void) { return 100; }

int nlk_partition(const int64_t* pid, int64_t nelg, int32_t nranks, int32_t* gllnid) {
  if (nranks <= 0 || (nranks & (nranks - 1))) { set_error("nranks must be a power of two"); return 1; }
  int64_t mx = 0; for (int64_t e = 0; e < nelg; ++e) mx = std::max(mx, pid[e]);
  int64_t npstar = 1; while (npstar < mx + 1) npstar *= 2;
  if (nranks > npstar) { set_error("more ranks than partition leaves"); return 1; }
  int64_t div = npstar / nranks;
  for (int64_t e = 0; e < nelg; ++e) gllnid[e] = (int32_t)(pid[e] / div);
  return 0;
}

int nlk_mesh_create(const nlk_mesh_desc* d, nlk_mesh** out) {
  if (!d || !out) { set_error("null argument"); return 1; }
  nlk_mesh* m = new nlk_mesh();
  int rc = build_mesh(m->hm, d->ndim, d->lx1, d->lxd, d->nelg, d->nel, d->xm1, d->ym1, d->zm1, d->vertex, d->cbc_v, d->cbc_t,
                      d->gllnid, d->rank, d->nranks <= 0 ? 1 : d->nranks);
  if (rc) { delete m; return rc; }
  *out = m; return 0;
}
int nlk_mesh_destroy(nlk_mesh* m) { delete m; return 0; }

int nlk_mesh_info(const nlk_mesh* m, nlk_mesh_info_t* o) {
  const HostMesh& h = m->hm;
  o->ndim = h.ndim; o->lx1 = h.n; o->lx2 = h.q; o->lxd = h.m; o->nel = h.E; o->nelg = h.Eg; o->np1 = h.np1; o->np2 = h.np2;
  o->nglob_local = h.nglob_local; o->nshared_local = (int64_t)h.gs_off.size() - 1; o->nvert = h.nvert;
  o->has_outflow = h.has_outflow; o->nneigh = (int32_t)h.neigh.size(); o->volvm1 = h.volvm1; o->volvm2 = h.volvm2;
  return 0;
}
int nlk_mesh_glo_num(const nlk_mesh* m, int64_t* glo) { std::copy(m->hm.glo.begin(), m->hm.glo.end(), glo); return 0; }

int nlk_mesh_field(const nlk_mesh* m, const char* name, double* out) {
  const HostMesh& h = m->hm; std::string s(name);
  const std::vector<double>* v = nullptr;
  if (s == "bm1") v = &h.bm1; else if (s == "jac") v = &h.jac; else if (s == "binvm1") v = &h.binvm1; else if (s == "vmult") v = &h.vmult;
  else if (s == "vmask0") v = &h.vmask[0]; else if (s == "vmask1") v = &h.vmask[1]; else if (s == "vmask2") v = &h.vmask[2];
  else if (s == "tmask") v = &h.tmask; else if (s == "bm2") v = &h.bm2;
  else if (s == "g11") v = &h.G[0]; else if (s == "g22") v = &h.G[1]; else if (s == "g33") v = &h.G[2];
  else if (s == "g12") v = &h.G[3]; else if (s == "g13") v = &h.G[4]; else if (s == "g23") v = &h.G[5];
  if (!v || v->empty()) { set_error("unknown or empty mesh field: " + s); return 1; }
  std::copy(v->begin(), v->end(), out); return 0;
}

int nlk_mesh_neighbor(const nlk_mesh* m, int32_t k, int32_t* rank, int64_t* count, int64_t* gids) {
  const HostMesh& h = m->hm;
  if (k < 0 || k >= (int)h.neigh.size()) { set_error("neighbour index out of range"); return 1; }
  *rank = h.neigh[k].rank; *count = (int64_t)h.neigh[k].gids.size();
  if (gids) std::copy(h.neigh[k].gids.begin(), h.neigh[k].gids.end(), gids);
  return 0;
}

int nlk_mesh_basis(const nlk_mesh* m, const char* name, double* out, int64_t cap) {
  const Basis& b = m->hm.b; std::string s(name);
  const std::vector<double>* v = nullptr;
  if (s == "z1") v = &b.z1; else if (s == "w1") v = &b.w1; else if (s == "z2") v = &b.z2; else if (s == "w2") v = &b.w2;
  else if (s == "zd") v = &b.zd; else if (s == "wd") v = &b.wd; else if (s == "D") v = &b.D; else if (s == "I12") v = &b.I12;
  else if (s == "D12") v = &b.D12; else if (s == "I1d") v = &b.I1d; else if (s == "Dd") v = &b.Dd; else if (s == "I21") v = &b.I21;
  if (!v) { set_error("unknown basis array: " + s); return -1; }
  if ((int64_t)v->size() > cap) { set_error("buffer too small"); return -1; }
  std::copy(v->begin(), v->end(), out); return (int)v->size();
}

}  // extern "C"
