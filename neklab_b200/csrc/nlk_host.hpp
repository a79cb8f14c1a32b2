// Host-side structures of libnlk: 1-D spectral-element operators and mesh setup.
// Restates Nek5000 start-up objects (speclib.f / coef.f / navier8.f set_vert; un-vendored upstream,
// SURVEY.md App. A.1, A.4).  No CUDA here: everything in this header is testable on a CPU box.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <array>

namespace nlk {

void set_error(const std::string& s);
const char* get_error();

struct Basis {
  int n = 0, m = 0, q = 0;                 // lx1, lxd, lx2
  std::vector<double> z1, w1, z2, w2, zd, wd;
  std::vector<double> D;                   // n x n   (dxm1)   D[i*n+j] = l_j'(z1_i)
  std::vector<double> I12, D12;            // q x n   (ixm12, dxm12)
  std::vector<double> I21;                 // n x q
  std::vector<double> I1d;                 // m x n   GLL -> fine GL
  std::vector<double> Dd;                  // m x m
  void build(int lx1, int lxd);
};

void gll_nodes(int n, std::vector<double>& z, std::vector<double>& w);
void gl_nodes(int n, std::vector<double>& z, std::vector<double>& w);
void interp_matrix(const std::vector<double>& xto, const std::vector<double>& xfrom, std::vector<double>& M);
void deriv_matrix(const std::vector<double>& x, std::vector<double>& D);

struct Neighbor {
  int rank;
  std::vector<int64_t> gids;               // shared global node ids, ascending (same order on both sides)
  std::vector<int32_t> rep;                // representative local index (into [E*np1]) for each gid
};

struct HostMesh {
  int ndim = 2, n = 0, m = 0, q = 0;
  int64_t E = 0, Eg = 0;
  int np1 = 0, np2 = 0, npd = 0, nv = 0, nfaces = 0;
  int rank = 0, nranks = 1;
  Basis b;
  std::vector<int64_t> lglel;              // local -> global element (0-based)
  std::vector<double> xyz[3];
  std::vector<int64_t> vertex_local;       // E*nv (global vertex ids, 1-based)
  std::vector<int64_t> vertex_all;         // Eg*nv (all elements; coarse-operator sparsity and colouring)
  std::vector<std::array<char, 3>> cbc_v, cbc_t;
  bool has_tbc = false;
  // geometry (all element-major, np1/np2/npd per element)
  std::vector<double> jac, bm1, binvm1, vmult, bm2, bm2inv;
  std::vector<double> rx[9];               // rx[k*ndim+c] = J * d r_k / d x_c      (mesh 1)
  std::vector<double> G[6];                // g11 g22 g33 g12 g13 g23 (Nek g1..g6 order)
  std::vector<double> rxw2[9];             // W2 * rx interpolated to mesh 2
  std::vector<double> rxd[9];              // dealias metrics with GL weights folded in
  std::vector<double> vmask[3], tmask;
  double volvm1 = 0, volvm2 = 0;           // local
  bool has_outflow = false;                // global property
  int64_t nvert = 0;
  // numbering
  std::vector<int64_t> glo;                // E*np1, 1-based global node ids
  std::vector<int32_t> gs_off, gs_idx;     // CSR over local shared nodes (multiplicity >= 2 locally)
  std::vector<int32_t> gs_first;           // representative local index of every shared group
  int64_t nglob_local = 0;
  std::vector<Neighbor> neigh;
  // interface unpack CSR: for every interface global node (any neighbour), its local copies
  std::vector<int64_t> if_gids;
  std::vector<int32_t> if_off, if_idx;
  // Schwarz/FDM setup data (pressure preconditioner)
  std::vector<double> fdmS;                // E * ndim * n * n
  std::vector<double> fdmDinv;             // E * np1 (inverse eigenvalue sums, 0 where inactive)
  std::vector<double> schwarz_wt;          // E * np2
  // coarse
  std::vector<int32_t> vert_off, vert_ec;  // CSR vertex -> (e*nv + c) local incidences
  std::vector<int64_t> vert_ids;           // (unused for single rank) global vertex ids present locally
};

int build_mesh(HostMesh& hm, int ndim, int lx1, int lxd, int64_t nelg, int64_t nel, const double* xm1, const double* ym1,
               const double* zm1, const int64_t* vertex_all, const char* cbc_v, const char* cbc_t,
               const int32_t* gllnid, int rank, int nranks);

// dense helpers (nlk_dense.cpp)
int dense_eig(int n, const double* A, double* wr, double* wi, double* VR);
int sym_eig_jacobi(int n, double* A /* in: sym, out: destroyed */, double* w, double* V /* columns = eigenvectors, row-major V[i*n+j] */);
int spd_inverse(int n, double* A /* in/out row-major */, bool singular_ok);

}  // namespace nlk
