// 1-D spectral-element operators (Nek speclib.f analogue: zwgll, zwgl, dgll, igllm ...).
// Published Legendre/Lagrange formulas; consumers in the reference: src/linops/neklab_linops.f90:343-362.
#include "nlk_host.hpp"
#include <cmath>
#include <mutex>

namespace nlk {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
const char* get_error() { return g_err.c_str(); }

static void legendre(int n, double x, double& p, double& dp) {
  if (n == 0) { p = 1; dp = 0; return; }
  double p0 = 1, p1 = x, d0 = 0, d1 = 1;
  for (int k = 1; k < n; ++k) {
    double p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1);
    double d2 = d0 + (2 * k + 1) * p1;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  p = p1; dp = d1;
}

void gll_nodes(int n, std::vector<double>& z, std::vector<double>& w) {
  const int N = n - 1;
  z.assign(n, 0.0); w.assign(n, 0.0);
  for (int i = 0; i < n; ++i) z[i] = -std::cos(M_PI * i / N);
  for (int i = 1; i < n - 1; ++i) {
    double x = z[i];
    for (int it = 0; it < 100; ++it) {
      double p, dp; legendre(N, x, p, dp);
      double ddp = (2 * x * dp - N * (N + 1) * p) / (1 - x * x);
      double dx = -dp / ddp;
      x += dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    z[i] = x;
  }
  z[0] = -1; z[n - 1] = 1;
  for (int i = 0; i < n / 2; ++i) { double a = 0.5 * (z[i] - z[n - 1 - i]); z[i] = a; z[n - 1 - i] = -a; }
  if (n % 2) z[n / 2] = 0.0;
  for (int i = 0; i < n; ++i) { double p, dp; legendre(N, z[i], p, dp); w[i] = 2.0 / (N * (N + 1) * p * p); }
}

void gl_nodes(int n, std::vector<double>& z, std::vector<double>& w) {
  z.assign(n, 0.0); w.assign(n, 0.0);
  for (int i = 0; i < n; ++i) {
    double x = -std::cos(M_PI * (i + 0.5) / n);
    for (int it = 0; it < 100; ++it) {
      double p, dp; legendre(n, x, p, dp);
      double dx = -p / dp; x += dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    z[i] = x;
  }
  for (int i = 0; i < n / 2; ++i) { double a = 0.5 * (z[i] - z[n - 1 - i]); z[i] = a; z[n - 1 - i] = -a; }
  if (n % 2) z[n / 2] = 0.0;
  for (int i = 0; i < n; ++i) { double p, dp; legendre(n, z[i], p, dp); w[i] = 2.0 / ((1 - z[i] * z[i]) * dp * dp); }
}

static void bary(const std::vector<double>& x, std::vector<double>& bw) {
  int n = (int)x.size(); bw.assign(n, 1.0);
  for (int j = 0; j < n; ++j) { for (int k = 0; k < n; ++k) if (k != j) bw[j] *= (x[j] - x[k]); bw[j] = 1.0 / bw[j]; }
}

void interp_matrix(const std::vector<double>& xto, const std::vector<double>& xfrom, std::vector<double>& M) {
  int mo = (int)xto.size(), ni = (int)xfrom.size();
  std::vector<double> bw; bary(xfrom, bw);
  M.assign((size_t)mo * ni, 0.0);
  for (int i = 0; i < mo; ++i) {
    int hit = -1;
    for (int j = 0; j < ni; ++j) if (xto[i] == xfrom[j]) hit = j;
    if (hit >= 0) { M[(size_t)i * ni + hit] = 1.0; continue; }
    double s = 0;
    for (int j = 0; j < ni; ++j) { double t = bw[j] / (xto[i] - xfrom[j]); M[(size_t)i * ni + j] = t; s += t; }
    for (int j = 0; j < ni; ++j) M[(size_t)i * ni + j] /= s;
  }
}

void deriv_matrix(const std::vector<double>& x, std::vector<double>& D) {
  int n = (int)x.size();
  std::vector<double> bw; bary(x, bw);
  D.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) if (i != j) { D[(size_t)i * n + j] = (bw[j] / bw[i]) / (x[i] - x[j]); s += D[(size_t)i * n + j]; }
    D[(size_t)i * n + i] = -s;
  }
}

static void matmul(const std::vector<double>& A, int ar, int ac, const std::vector<double>& B, int bc, std::vector<double>& C) {
  C.assign((size_t)ar * bc, 0.0);
  for (int i = 0; i < ar; ++i) for (int k = 0; k < ac; ++k) { double a = A[(size_t)i * ac + k]; for (int j = 0; j < bc; ++j) C[(size_t)i * bc + j] += a * B[(size_t)k * bc + j]; }
}

void Basis::build(int lx1, int lxd) {
  n = lx1; m = lxd; q = lx1 - 2;
  gll_nodes(n, z1, w1); gl_nodes(q, z2, w2); gl_nodes(m, zd, wd);
  deriv_matrix(z1, D);
  interp_matrix(z2, z1, I12);
  matmul(I12, q, n, D, n, D12);
  interp_matrix(z1, z2, I21);
  interp_matrix(zd, z1, I1d);
  deriv_matrix(zd, Dd);
}

}  // namespace nlk
