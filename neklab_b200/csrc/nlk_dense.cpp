// Small dense host kernels (LAPACK-free): real nonsymmetric eigenproblem (Householder Hessenberg reduction +
// shifted QR, the classical EISPACK orthes/hqr2 algorithm), cyclic Jacobi for symmetric matrices, SPD inverse.
// Replaces the LAPACK calls LightKrylov makes through stdlib_linalg for the k x k Hessenberg/Rayleigh matrices
// (reference call sites: src/neklab_otd.f90:229,248).  k <= kdim (128..512): host cost is negligible.
#include "nlk_host.hpp"
#include "../../include/nlk.h"
#include <algorithm>
#include <cmath>
#include <vector>

namespace nlk {

namespace {
struct EigWork {
  int n;
  std::vector<double> H, V, d, e, ort;
  double& h(int i, int j) { return H[(size_t)i * n + j]; }
  double& v(int i, int j) { return V[(size_t)i * n + j]; }
};

void cdiv(double xr, double xi, double yr, double yi, double& cr, double& ci) {
  double r, dd;
  if (std::fabs(yr) > std::fabs(yi)) { r = yi / yr; dd = yr + r * yi; cr = (xr + r * xi) / dd; ci = (xi - r * xr) / dd; }
  else { r = yr / yi; dd = yi + r * yr; cr = (r * xr + xi) / dd; ci = (r * xi - xr) / dd; }
}

void orthes(EigWork& w) {
  const int n = w.n, low = 0, high = n - 1;
  for (int m = low + 1; m <= high - 1; ++m) {
    double scale = 0.0;
    for (int i = m; i <= high; ++i) scale += std::fabs(w.h(i, m - 1));
    if (scale != 0.0) {
      double hh = 0.0;
      for (int i = high; i >= m; --i) { w.ort[i] = w.h(i, m - 1) / scale; hh += w.ort[i] * w.ort[i]; }
      double g = std::sqrt(hh);
      if (w.ort[m] > 0) g = -g;
      hh -= w.ort[m] * g; w.ort[m] -= g;
      for (int j = m; j < n; ++j) {
        double f = 0.0;
        for (int i = high; i >= m; --i) f += w.ort[i] * w.h(i, j);
        f /= hh;
        for (int i = m; i <= high; ++i) w.h(i, j) -= f * w.ort[i];
      }
      for (int i = 0; i <= high; ++i) {
        double f = 0.0;
        for (int j = high; j >= m; --j) f += w.ort[j] * w.h(i, j);
        f /= hh;
        for (int j = m; j <= high; ++j) w.h(i, j) -= f * w.ort[j];
      }
      w.ort[m] = scale * w.ort[m];
      w.h(m, m - 1) = scale * g;
    }
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) w.v(i, j) = (i == j ? 1.0 : 0.0);
  for (int m = high - 1; m >= low + 1; --m) {
    if (w.h(m, m - 1) != 0.0) {
      for (int i = m + 1; i <= high; ++i) w.ort[i] = w.h(i, m - 1);
      for (int j = m; j <= high; ++j) {
        double g = 0.0;
        for (int i = m; i <= high; ++i) g += w.ort[i] * w.v(i, j);
        g = (g / w.ort[m]) / w.h(m, m - 1);
        for (int i = m; i <= high; ++i) w.v(i, j) += g * w.ort[i];
      }
    }
  }
}

int hqr2(EigWork& w) {
  const int nn = w.n;
  int n = nn - 1;
  const int low = 0, high = nn - 1;
  const double eps = std::pow(2.0, -52.0);
  double exshift = 0.0, p = 0, q = 0, r = 0, s = 0, z = 0, t, ww, x, y;
  double norm = 0.0;
  for (int i = 0; i < nn; ++i) for (int j = std::max(i - 1, 0); j < nn; ++j) norm += std::fabs(w.h(i, j));
  int iter = 0, total = 0;
  while (n >= low) {
    int l = n;
    while (l > low) {
      s = std::fabs(w.h(l - 1, l - 1)) + std::fabs(w.h(l, l));
      if (s == 0.0) s = norm;
      if (std::fabs(w.h(l, l - 1)) < eps * s) break;
      --l;
    }
    if (l == n) {                       // one root
      w.h(n, n) += exshift; w.d[n] = w.h(n, n); w.e[n] = 0.0; --n; iter = 0;
    } else if (l == n - 1) {            // two roots
      ww = w.h(n, n - 1) * w.h(n - 1, n);
      p = (w.h(n - 1, n - 1) - w.h(n, n)) / 2.0;
      q = p * p + ww;
      z = std::sqrt(std::fabs(q));
      w.h(n, n) += exshift; w.h(n - 1, n - 1) += exshift;
      x = w.h(n, n);
      if (q >= 0) {                     // real pair
        z = p >= 0 ? p + z : p - z;
        w.d[n - 1] = x + z; w.d[n] = w.d[n - 1];
        if (z != 0.0) w.d[n] = x - ww / z;
        w.e[n - 1] = 0.0; w.e[n] = 0.0;
        x = w.h(n, n - 1);
        s = std::fabs(x) + std::fabs(z);
        p = x / s; q = z / s;
        r = std::sqrt(p * p + q * q);
        p /= r; q /= r;
        for (int j = n - 1; j < nn; ++j) { z = w.h(n - 1, j); w.h(n - 1, j) = q * z + p * w.h(n, j); w.h(n, j) = q * w.h(n, j) - p * z; }
        for (int i = 0; i <= n; ++i) { z = w.h(i, n - 1); w.h(i, n - 1) = q * z + p * w.h(i, n); w.h(i, n) = q * w.h(i, n) - p * z; }
        for (int i = low; i <= high; ++i) { z = w.v(i, n - 1); w.v(i, n - 1) = q * z + p * w.v(i, n); w.v(i, n) = q * w.v(i, n) - p * z; }
      } else {                          // complex pair
        w.d[n - 1] = x + p; w.d[n] = x + p; w.e[n - 1] = z; w.e[n] = -z;
      }
      n -= 2; iter = 0;
    } else {
      x = w.h(n, n); y = 0.0; ww = 0.0;
      if (l < n) { y = w.h(n - 1, n - 1); ww = w.h(n, n - 1) * w.h(n - 1, n); }
      if (iter == 10) {                 // Wilkinson's original ad hoc shift
        exshift += x;
        for (int i = low; i <= n; ++i) w.h(i, i) -= x;
        s = std::fabs(w.h(n, n - 1)) + std::fabs(w.h(n - 1, n - 2));
        x = y = 0.75 * s; ww = -0.4375 * s * s;
      }
      if (iter == 30) {                 // MATLAB's new ad hoc shift
        s = (y - x) / 2.0; s = s * s + ww;
        if (s > 0) {
          s = std::sqrt(s); if (y < x) s = -s;
          s = x - ww / ((y - x) / 2.0 + s);
          for (int i = low; i <= n; ++i) w.h(i, i) -= s;
          exshift += s; x = y = ww = 0.964;
        }
      }
      ++iter; ++total;
      if (total > 300 * nn) return 1;
      int m = n - 2;
      while (m >= l) {
        z = w.h(m, m); r = x - z; s = y - z;
        p = (r * s - ww) / w.h(m + 1, m) + w.h(m, m + 1);
        q = w.h(m + 1, m + 1) - z - r - s;
        r = w.h(m + 2, m + 1);
        s = std::fabs(p) + std::fabs(q) + std::fabs(r);
        p /= s; q /= s; r /= s;
        if (m == l) break;
        if (std::fabs(w.h(m, m - 1)) * (std::fabs(q) + std::fabs(r)) <
            eps * (std::fabs(p) * (std::fabs(w.h(m - 1, m - 1)) + std::fabs(z) + std::fabs(w.h(m + 1, m + 1))))) break;
        --m;
      }
      for (int i = m + 2; i <= n; ++i) { w.h(i, i - 2) = 0.0; if (i > m + 2) w.h(i, i - 3) = 0.0; }
      for (int k = m; k <= n - 1; ++k) {
        bool notlast = (k != n - 1);
        if (k != m) {
          p = w.h(k, k - 1); q = w.h(k + 1, k - 1); r = notlast ? w.h(k + 2, k - 1) : 0.0;
          x = std::fabs(p) + std::fabs(q) + std::fabs(r);
          if (x == 0.0) continue;
          p /= x; q /= x; r /= x;
        }
        s = std::sqrt(p * p + q * q + r * r);
        if (p < 0) s = -s;
        if (s != 0) {
          if (k != m) w.h(k, k - 1) = -s * x;
          else if (l != m) w.h(k, k - 1) = -w.h(k, k - 1);
          p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
          for (int j = k; j < nn; ++j) {
            p = w.h(k, j) + q * w.h(k + 1, j);
            if (notlast) { p += r * w.h(k + 2, j); w.h(k + 2, j) -= p * z; }
            w.h(k, j) -= p * x; w.h(k + 1, j) -= p * y;
          }
          for (int i = 0; i <= std::min(n, k + 3); ++i) {
            p = x * w.h(i, k) + y * w.h(i, k + 1);
            if (notlast) { p += z * w.h(i, k + 2); w.h(i, k + 2) -= p * r; }
            w.h(i, k) -= p; w.h(i, k + 1) -= p * q;
          }
          for (int i = low; i <= high; ++i) {
            p = x * w.v(i, k) + y * w.v(i, k + 1);
            if (notlast) { p += z * w.v(i, k + 2); w.v(i, k + 2) -= p * r; }
            w.v(i, k) -= p; w.v(i, k + 1) -= p * q;
          }
        }
      }
    }
  }
  // back-substitute for the vectors of the upper-triangular (quasi) form
  if (norm == 0.0) return 0;
  for (n = nn - 1; n >= 0; --n) {
    p = w.d[n]; q = w.e[n];
    if (q == 0) {                       // real vector
      int l = n;
      w.h(n, n) = 1.0;
      for (int i = n - 1; i >= 0; --i) {
        ww = w.h(i, i) - p; r = 0.0;
        for (int j = l; j <= n; ++j) r += w.h(i, j) * w.h(j, n);
        if (w.e[i] < 0.0) { z = ww; s = r; }
        else {
          l = i;
          if (w.e[i] == 0.0) {
            w.h(i, n) = (ww != 0.0) ? -r / ww : -r / (eps * norm);
          } else {
            x = w.h(i, i + 1); y = w.h(i + 1, i);
            q = (w.d[i] - p) * (w.d[i] - p) + w.e[i] * w.e[i];
            t = (x * s - z * r) / q;
            w.h(i, n) = t;
            w.h(i + 1, n) = (std::fabs(x) > std::fabs(z)) ? (-r - ww * t) / x : (-s - y * t) / z;
          }
          t = std::fabs(w.h(i, n));
          if ((eps * t) * t > 1) for (int j = i; j <= n; ++j) w.h(j, n) /= t;
        }
      }
    } else if (q < 0) {                 // complex vector (last of the pair)
      int l = n - 1;
      if (std::fabs(w.h(n, n - 1)) > std::fabs(w.h(n - 1, n))) {
        w.h(n - 1, n - 1) = q / w.h(n, n - 1);
        w.h(n - 1, n) = -(w.h(n, n) - p) / w.h(n, n - 1);
      } else {
        double cr, ci; cdiv(0.0, -w.h(n - 1, n), w.h(n - 1, n - 1) - p, q, cr, ci);
        w.h(n - 1, n - 1) = cr; w.h(n - 1, n) = ci;
      }
      w.h(n, n - 1) = 0.0; w.h(n, n) = 1.0;
      for (int i = n - 2; i >= 0; --i) {
        double ra = 0.0, sa = 0.0, vr, vi;
        for (int j = l; j <= n; ++j) { ra += w.h(i, j) * w.h(j, n - 1); sa += w.h(i, j) * w.h(j, n); }
        ww = w.h(i, i) - p;
        if (w.e[i] < 0.0) { z = ww; r = ra; s = sa; }
        else {
          l = i;
          if (w.e[i] == 0) {
            double cr, ci; cdiv(-ra, -sa, ww, q, cr, ci);
            w.h(i, n - 1) = cr; w.h(i, n) = ci;
          } else {
            x = w.h(i, i + 1); y = w.h(i + 1, i);
            vr = (w.d[i] - p) * (w.d[i] - p) + w.e[i] * w.e[i] - q * q;
            vi = (w.d[i] - p) * 2.0 * q;
            if (vr == 0.0 && vi == 0.0) vr = eps * norm * (std::fabs(ww) + std::fabs(q) + std::fabs(x) + std::fabs(y) + std::fabs(z));
            double cr, ci; cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, cr, ci);
            w.h(i, n - 1) = cr; w.h(i, n) = ci;
            if (std::fabs(x) > (std::fabs(z) + std::fabs(q))) {
              w.h(i + 1, n - 1) = (-ra - ww * w.h(i, n - 1) + q * w.h(i, n)) / x;
              w.h(i + 1, n) = (-sa - ww * w.h(i, n) - q * w.h(i, n - 1)) / x;
            } else {
              cdiv(-r - y * w.h(i, n - 1), -s - y * w.h(i, n), z, q, cr, ci);
              w.h(i + 1, n - 1) = cr; w.h(i + 1, n) = ci;
            }
          }
          t = std::max(std::fabs(w.h(i, n - 1)), std::fabs(w.h(i, n)));
          if ((eps * t) * t > 1) for (int j = i; j <= n; ++j) { w.h(j, n - 1) /= t; w.h(j, n) /= t; }
        }
      }
    }
  }
  for (int j = nn - 1; j >= low; --j) {
    for (int i = low; i <= high; ++i) {
      z = 0.0;
      for (int k = low; k <= std::min(j, high); ++k) z += w.v(i, k) * w.h(k, j);
      w.v(i, j) = z;
    }
  }
  return 0;
}
}  // namespace

// eigenvalues (wr, wi) and right eigenvectors VR (row-major n x n).  Complex pairs (wi[j] > 0 then wi[j+1] < 0):
// column j holds the real part and column j+1 the imaginary part of the eigenvector of wr[j] + i wi[j]
// (LAPACK dgeev convention).  Vectors are normalised to unit 2-norm.
int dense_eig(int n, const double* A, double* wr, double* wi, double* VR) {
  if (n <= 0) return 0;
  EigWork w; w.n = n; w.H.assign(A, A + (size_t)n * n); w.V.assign((size_t)n * n, 0.0); w.d.assign(n, 0); w.e.assign(n, 0); w.ort.assign(n, 0);
  orthes(w);
  if (hqr2(w)) { set_error("dense_eig: QR iteration did not converge"); return 1; }
  for (int j = 0; j < n; ++j) { wr[j] = w.d[j]; wi[j] = w.e[j]; }
  if (VR) {
    for (int j = 0; j < n; ++j) {
      if (wi[j] == 0.0) {
        double s = 0; for (int i = 0; i < n; ++i) s += w.v(i, j) * w.v(i, j);
        s = s > 0 ? 1.0 / std::sqrt(s) : 1.0;
        for (int i = 0; i < n; ++i) VR[(size_t)i * n + j] = w.v(i, j) * s;
      } else if (wi[j] > 0 && j + 1 < n) {
        double s = 0; for (int i = 0; i < n; ++i) s += w.v(i, j) * w.v(i, j) + w.v(i, j + 1) * w.v(i, j + 1);
        s = s > 0 ? 1.0 / std::sqrt(s) : 1.0;
        for (int i = 0; i < n; ++i) { VR[(size_t)i * n + j] = w.v(i, j) * s; VR[(size_t)i * n + j + 1] = w.v(i, j + 1) * s; }
        ++j;
      }
    }
  }
  return 0;
}

// cyclic Jacobi for a symmetric matrix; eigenvalues ascending in w, eigenvectors in the columns of V.
int sym_eig_jacobi(int n, double* A, double* w, double* V) {
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, dg = 0;
    for (int i = 0; i < n; ++i) { dg += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-32 * (dg + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
      double apq = A[p * n + q];
      if (apq == 0.0) continue;
      double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < n; ++k) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; ++k) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; ++k) { double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  std::vector<int> ord(n); for (int i = 0; i < n; ++i) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
  std::vector<double> Vt(V, V + (size_t)n * n);
  for (int j = 0; j < n; ++j) { w[j] = A[ord[j] * n + ord[j]]; for (int i = 0; i < n; ++i) V[i * n + j] = Vt[i * n + ord[j]]; }
  return 0;
}

// in-place inverse of a symmetric positive (semi-)definite matrix via Cholesky; with singular_ok, pivots below
// 1e-12*max are treated as a null direction (pseudo-inverse restricted to the remaining unknowns).
int spd_inverse(int n, double* A, bool singular_ok) {
  std::vector<double> L((size_t)n * n, 0.0);
  std::vector<char> dead(n, 0);
  double dmax = 0; for (int i = 0; i < n; ++i) dmax = std::max(dmax, A[(size_t)i * n + i]);
  for (int j = 0; j < n; ++j) {
    double s = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) s -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
    if (s <= 1e-12 * dmax) {
      if (!singular_ok) { set_error("spd_inverse: matrix not positive definite"); return 1; }
      dead[j] = 1; L[(size_t)j * n + j] = 1.0;
      for (int i = j + 1; i < n; ++i) L[(size_t)i * n + j] = 0.0;
      continue;
    }
    double ljj = std::sqrt(s); L[(size_t)j * n + j] = ljj;
#pragma omp parallel for schedule(static)
    for (int i = j + 1; i < n; ++i) {
      double t = A[(size_t)i * n + j];
      const double* li = &L[(size_t)i * n]; const double* lj = &L[(size_t)j * n];
      for (int k = 0; k < j; ++k) t -= li[k] * lj[k];
      L[(size_t)i * n + j] = t / ljj;
    }
  }
  // invert L (lower triangular) column by column: Linv
  std::vector<double> Li((size_t)n * n, 0.0);
#pragma omp parallel for schedule(dynamic, 8)
  for (int c = 0; c < n; ++c) {
    if (dead[c]) continue;
    Li[(size_t)c * n + c] = 1.0 / L[(size_t)c * n + c];
    for (int i = c + 1; i < n; ++i) {
      if (dead[i]) continue;
      double t = 0; const double* li = &L[(size_t)i * n];
      for (int k = c; k < i; ++k) t += li[k] * Li[(size_t)k * n + c];
      Li[(size_t)i * n + c] = -t / L[(size_t)i * n + i];
    }
  }
  // A^-1 = Linv^T Linv
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) {
    double t = 0;
    if (!dead[i] && !dead[j]) for (int k = i; k < n; ++k) t += Li[(size_t)k * n + i] * Li[(size_t)k * n + j];
    A[(size_t)i * n + j] = t; A[(size_t)j * n + i] = t;
  }
  return 0;
}

}  // namespace nlk

extern "C" int nlk_dense_eig(int32_t n, const double* A, double* wr, double* wi, double* VR) { return nlk::dense_eig(n, A, wr, wi, VR); }
