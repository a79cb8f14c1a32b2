// Plane Poiseuille flow, linear stability through the C++ host layer (include/neklab.hpp) over the C-ABI of libnlk --
// the compiled-host counterpart of the reference's `examples/poiseuille/stability/direct_alpha_1/poiseuille.usr` userchk:
//
//     exptA = exptA_linop(1.0_dp, bf) ; call exptA%init()
//     call linear_stability_analysis_fixed_point(exptA, kdim, nev)
//
//   g++ -std=c++17 -O2 -Iinclude examples/channel_eigs.cpp -Lneklab_b200 -lnlk -Wl,-rpath,$PWD/neklab_b200 -o channel_eigs
//   ./channel_eigs [Re=7500] [kdim=100] [nev=2] [outdir=.]
//
// Mesh: nx x ny elements on [0, 2 pi] x [-1, 1], periodic in x, cosine-graded towards the walls, lx1 = 8, lxd = 12.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "neklab.hpp"

// Gauss-Lobatto-Legendre nodes on [-1, 1]: Newton iteration on (1 - x^2) P'_{n-1}(x)
static std::vector<double> gll_nodes(int n) {
  const int N = n - 1;
  const double pi = std::acos(-1.0);
  std::vector<double> x(n);
  for (int i = 0; i < n; ++i) x[i] = -std::cos(pi * i / N);
  for (int i = 1; i < N; ++i) {
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x[i];
      for (int k = 1; k < N; ++k) { const double p2 = ((2 * k + 1) * x[i] * p1 - k * p0) / (k + 1); p0 = p1; p1 = p2; }
      // p1 = P_N, p0 = P_{N-1};  P'_N = N (P_{N-1} - x P_N) / (1 - x^2);  q = (1 - x^2) P'_N has q' = -N (N + 1) P_N
      const double q = N * (p0 - x[i] * p1), dq = -N * (N + 1.0) * p1;
      const double dx = -q / dq;
      x[i] += dx;
      if (std::fabs(dx) < 1e-15) break;
    }
  }
  return x;
}

int main(int argc, char** argv) {
  const double Re = argc > 1 ? std::atof(argv[1]) : 7500.0;
  const int kdim = argc > 2 ? std::atoi(argv[2]) : 100;
  const int nev = argc > 3 ? std::atoi(argv[3]) : 2;
  const std::string outdir = argc > 4 ? argv[4] : ".";
  const int nx = 8, ny = 10, n = 8, lxd = 12;
  const double pi = std::acos(-1.0), Lx = 2.0 * pi;
  const std::int64_t E = (std::int64_t)nx * ny;
  const int np = n * n;
  try {
    // ---- mesh description (what userchk would take from xm1/ym1, the .ma2 vertex ids and cbc)
    const std::vector<double> z = gll_nodes(n);
    std::vector<double> yb(ny + 1);
    for (int j = 0; j <= ny; ++j) yb[j] = -std::cos(pi * j / ny);
    std::vector<double> xm1(E * np), ym1(E * np);
    std::vector<std::int64_t> vertex(E * 4);
    std::vector<char> cbc(E * 4 * 3);
    for (int ej = 0; ej < ny; ++ej) for (int ei = 0; ei < nx; ++ei) {
      const std::int64_t e = (std::int64_t)ej * nx + ei;
      const double x0 = Lx * ei / nx, hx = Lx / nx, y0 = yb[ej], hy = yb[ej + 1] - yb[ej];
      for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
        xm1[e * np + j * n + i] = x0 + 0.5 * (z[i] + 1.0) * hx;
        ym1[e * np + j * n + i] = y0 + 0.5 * (z[j] + 1.0) * hy;
      }
      auto vid = [&](int i, int j) { return (std::int64_t)1 + (i % nx) + (std::int64_t)nx * j; };     // x-periodic vertex ids
      vertex[e * 4 + 0] = vid(ei, ej); vertex[e * 4 + 1] = vid(ei + 1, ej); vertex[e * 4 + 2] = vid(ei, ej + 1); vertex[e * 4 + 3] = vid(ei + 1, ej + 1);
      const char* code[4] = {ej == 0 ? "W  " : "E  ", ei == nx - 1 ? "P  " : "E  ", ej == ny - 1 ? "W  " : "E  ", ei == 0 ? "P  " : "E  "};
      for (int f = 0; f < 4; ++f) std::memcpy(&cbc[(e * 4 + f) * 3], code[f], 3);
    }
    nlk_mesh_desc d{};
    d.ndim = 2; d.lx1 = n; d.lxd = lxd; d.nelg = E; d.nel = E;
    d.xm1 = xm1.data(); d.ym1 = ym1.data(); d.zm1 = nullptr; d.vertex = vertex.data(); d.cbc_v = cbc.data(); d.cbc_t = nullptr;
    d.gllnid = nullptr; d.rank = 0; d.nranks = 1;
    neklab::mesh m(d);
    const nlk_mesh_info_t info = m.info();
    std::printf("mesh: %lld elements, lx1 = %d, %lld unique nodes, pressure operator %s\n", (long long)info.nel, info.lx1,
                (long long)info.nglob_local, info.has_outflow ? "regular" : "singular (no outflow)");

    // ---- solver state: poiseuille.par (viscosity = -7500, bdf2, residualtol 1e-10)
    nlk_params p = neklab::context::default_params();
    p.viscosity = 1.0 / Re; p.torder = 2; p.vtol = 1e-10; p.ptol = 1e-10; p.pr_proj = 20; p.gmres_maxit = 400;
    neklab::context ctx(m, p, 0);

    // ---- base flow U = (1 - y^2, 0): nek2vec(bf, vx, vy, vz, pr, t)
    std::vector<double> ux(E * np), uy(E * np, 0.0);
    for (std::size_t i = 0; i < ux.size(); ++i) ux[i] = 1.0 - ym1[i] * ym1[i];
    neklab::nek_dvector bf(ctx);
    bf.nek2vec(ux.data(), uy.data(), nullptr, nullptr, nullptr);

    // ---- exptA = exptA_linop(1.0_dp, bf); call exptA%init(); call linear_stability_analysis_fixed_point(exptA, kdim, nev)
    neklab::exptA_linop exptA(1.0, bf);
    exptA.init();
    const nlk_stats s0 = exptA.stats();
    std::printf("exptA: tau = %.3f, dt = %.6f, nsteps = %d\n", exptA.tau, s0.dt, s0.nsteps);
    const neklab::eigs_result r = neklab::linear_stability_analysis_fixed_point(exptA, kdim, nev, false, outdir, 1e-7);
    for (int i = 0; i < nev; ++i)
      std::printf("eig %d: mu = %+.10f %+.10f i   sigma = %+.10f %+.10f i   residual %.2e\n", i, r.mu[i].real(), r.mu[i].imag(),
                  r.eigvals[i].real(), r.eigvals[i].imag(), r.residuals[i]);
    std::printf("niter = %d, info = %d\n", r.niter, r.info);
    return r.info == 0 ? 0 : 2;
  } catch (const neklab::error& e) {
    std::fprintf(stderr, "neklab error: %s\n", e.what());          // stop_error: no CPU fallback exists
    return 1;
  }
}
