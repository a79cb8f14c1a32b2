// neklab.hpp -- C++17 host-side mirror of neklab's operator / vector interface over the C-ABI of libnlk (include/nlk.h).
//
// neklab is compiled (Fortran) code; its toolchain is absent from this image, so this header plays the role of the
// ISO_C_BINDING shim sketched in INTEGRATION.md: the same type names, method names, argument meaning and error behaviour
// as the reference, as thin RAII wrappers around opaque device-resident handles.
//
//   reference (Fortran)                                               here
//   type nek_dvector            src/vectors/neklab_vectors.f90:26-113  neklab::nek_dvector
//   type exptA_linop            src/linops/neklab_linops.f90:35-75     neklab::exptA_linop
//   linear_stability_analysis_fixed_point  src/neklab_analysis.f90:38  neklab::linear_stability_analysis_fixed_point
//   transient_growth_analysis_fixed_point  src/neklab_analysis.f90:107 neklab::transient_growth_analysis_fixed_point
//   newton_fixed_point_iteration           src/neklab_analysis.f90:158 neklab::newton_fixed_point_iteration
//   stop_error / nek_stop_error            src/neklab_nek_setup.f90:406 neklab::error (exception carrying nlk_last_error())
//
// Header-only; link with -lnlk.  No CUDA or torch types appear here.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "nlk.h"

namespace neklab {

struct error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void check(int rc, const char* where) {
  if (rc != 0) throw error(std::string(where) + ": " + nlk_last_error());
}

// ---------------------------------------------------------------------------------------------------- mesh / context
// What Nek5000 derives from SIZE + .re2/.ma2 at start-up (geometry, numbering, masks); host only.
class mesh {
 public:
  explicit mesh(const nlk_mesh_desc& d) { check(nlk_mesh_create(&d, &h_), "nlk_mesh_create"); }
  ~mesh() { if (h_) nlk_mesh_destroy(h_); }
  mesh(const mesh&) = delete;
  mesh& operator=(const mesh&) = delete;
  nlk_mesh_info_t info() const { nlk_mesh_info_t i{}; check(nlk_mesh_info(h_, &i), "nlk_mesh_info"); return i; }
  const nlk_mesh* handle() const { return h_; }

 private:
  nlk_mesh* h_ = nullptr;
};

// Nek's COMMON-block solver state (base flow, perturbation, lag arrays, dt/nsteps, tolerances) on one GPU.  Like the
// reference it is not re-entrant: one context per process and device.
class context {
 public:
  static nlk_params default_params() { nlk_params p{}; nlk_params_default(&p); return p; }
  context(const mesh& m, const nlk_params& p, int device = 0) { check(nlk_ctx_create(m.handle(), &p, device, &h_), "nlk_ctx_create"); }
  ~context() { if (h_) nlk_ctx_destroy(h_); }
  context(const context&) = delete;
  context& operator=(const context&) = delete;
  void set_tol(double vtol, double ptol) { check(nlk_ctx_set_tol(h_, vtol, ptol), "nlk_ctx_set_tol"); }   // setup_nek(vtol=, ptol=)
  void set_dt(double dt) { check(nlk_ctx_set_dt(h_, dt), "nlk_ctx_set_dt"); }                                // zero base flow: preset dt
  void set_forcing(const double* fx, const double* fy, const double* fz) { check(nlk_ctx_set_forcing(h_, fx, fy, fz), "nlk_ctx_set_forcing"); }
  void comm_init(const char id[128], int rank, int nranks) { check(nlk_ctx_comm_init(h_, id, rank, nranks), "nlk_ctx_comm_init"); }
  void sync() { check(nlk_ctx_sync(h_), "nlk_ctx_sync"); }
  nlk_ctx* handle() const { return h_; }

 private:
  nlk_ctx* h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------- nek_dvector
// abstract_vector_rdp surface with the reference's semantics: the inner product is bm1-weighted over velocity (+ theta) and
// excludes the pressure, scal/axpby include it; copies are deep (Fortran value semantics: `allocate(X(k), source=x)`).
class nek_dvector {
 public:
  explicit nek_dvector(context& c) : c_(&c) { check(nlk_vec_create(c.handle(), &h_), "nlk_vec_create"); }
  nek_dvector(const nek_dvector& o) : c_(o.c_) {
    check(nlk_vec_create(c_->handle(), &h_), "nlk_vec_create");
    check(nlk_vec_copy(h_, o.h_), "nlk_vec_copy");
  }
  nek_dvector(nek_dvector&& o) noexcept : c_(o.c_), h_(o.h_) { o.h_ = nullptr; }
  nek_dvector& operator=(const nek_dvector& o) { if (this != &o) check(nlk_vec_copy(h_, o.h_), "nlk_vec_copy"); return *this; }
  nek_dvector& operator=(nek_dvector&& o) noexcept { std::swap(h_, o.h_); std::swap(c_, o.c_); return *this; }
  ~nek_dvector() { if (h_) nlk_vec_destroy(h_); }

  void zero() { check(nlk_vec_zero(h_), "nek_dzero"); }                                                    // real_vectors.f90:37-50
  void rand(bool ifnorm = false, std::uint64_t seed = 12345) { check(nlk_vec_rand(h_, ifnorm, seed), "nek_drand"); }   // :52-123
  void scal(double alpha) { check(nlk_vec_scal(h_, alpha), "nek_dscal"); }                                  // :125-160
  void axpby(double alpha, const nek_dvector& vec, double beta) { check(nlk_vec_axpby(alpha, vec.h_, beta, h_), "nek_daxpby"); }   // :162-206
  double dot(const nek_dvector& vec) const { double r = 0; check(nlk_vec_dot(h_, vec.h_, &r), "nek_ddot"); return r; }            // :208-233
  double norm() const { double r = 0; check(nlk_vec_norm(h_, &r), "nlk_vec_norm"); return r; }
  std::int64_t get_size() const { std::int64_t n = 0; check(nlk_vec_size(h_, &n), "nek_dsize"); return n; }                      // :235-247
  void save_rst(const nek_dvector& state, int irst) { check(nlk_vec_save_rst(h_, state.h_, irst), "dsave_rst"); }                // :249-291
  void get_rst(nek_dvector& out, int irst) const { check(nlk_vec_get_rst(h_, out.h_, irst), "dget_rst"); }                       // :293-333
  int nrst() const { std::int32_t n = 0; check(nlk_vec_nrst(h_, &n), "nlk_vec_nrst"); return n; }
  bool has_rst_fields() const { return nrst() > 0; }                                                                               // :335-338
  void clear_rst_fields() { check(nlk_vec_clear_rst(h_), "dclear_rst_fields"); }                                                  // :340-346
  // nek2vec / vec2nek (src/neklab_utils.f90:84-134): host arrays in Nek's (lx1,ly1,lz1,lelv) layout; null pointers are skipped
  void nek2vec(const double* vx, const double* vy, const double* vz, const double* pr, const double* t) { check(nlk_vec_upload(h_, vx, vy, vz, pr, t), "nek2vec"); }
  void vec2nek(double* vx, double* vy, double* vz, double* pr, double* t) const { check(nlk_vec_download(h_, vx, vy, vz, pr, t), "vec2nek"); }
  nlk_vec* handle() const { return h_; }
  context& ctx() const { return *c_; }

 private:
  context* c_;
  nlk_vec* h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------- exptA_linop
// abstract_exptA_linop_rdp: tau from the parent type, baseflow component; `exptA_linop(1.0_dp, bf); call exptA%init()`.
class exptA_linop {
 public:
  exptA_linop(double tau_, const nek_dvector& baseflow) : tau(tau_), c_(&baseflow.ctx()) {
    check(nlk_exptA_create(c_->handle(), tau_, baseflow.handle(), &h_), "exptA_linop");
  }
  ~exptA_linop() { if (h_) nlk_exptA_destroy(h_); }
  exptA_linop(const exptA_linop&) = delete;
  exptA_linop& operator=(const exptA_linop&) = delete;
  void init() { check(nlk_exptA_init(h_), "init_exptA"); }                                                                       // exponential_propagator.f90:4-13
  void matvec(const nek_dvector& vec_in, nek_dvector& vec_out) { sync_tau(); check(nlk_exptA_matvec(h_, vec_in.handle(), vec_out.handle()), "exptA_matvec"); }     // :15-60
  void rmatvec(const nek_dvector& vec_in, nek_dvector& vec_out) { sync_tau(); check(nlk_exptA_rmatvec(h_, vec_in.handle(), vec_out.handle()), "exptA_rmatvec"); }  // :62-107
  void set_baseflow(const nek_dvector& bf) { check(nlk_exptA_set_baseflow(h_, bf.handle()), "nlk_exptA_set_baseflow"); }
  nlk_stats stats() const { nlk_stats s{}; check(nlk_exptA_stats(h_, &s), "nlk_exptA_stats"); return s; }
  nlk_op* handle() const { return h_; }
  context& ctx() const { return *c_; }
  double tau;                                   // public component, as in the reference (apply_exptA sets A%tau, neklab_linops.f90:224-266)

 private:
  void sync_tau() { if (tau != tau_set_) { check(nlk_exptA_set_tau(h_, tau), "nlk_exptA_set_tau"); tau_set_ = tau; } }
  context* c_;
  nlk_op* h_ = nullptr;
  double tau_set_ = std::nan("");
};

// ---------------------------------------------------------------------------------------------------- nek_zvector / resolvent_linop
// nek_zvector (src/vectors/neklab_vectors.f90:219-237): complex vector as a (re, im) pair of nek_dvector
class nek_zvector {
 public:
  explicit nek_zvector(context& c) : re(c), im(c) {}
  void zero() { re.zero(); im.zero(); }
  void scal(std::complex<double> a) { check(nlk_zvec_scal(re.handle(), im.handle(), a.real(), a.imag()), "nek_zscal"); }
  void axpby(std::complex<double> a, const nek_zvector& v, std::complex<double> b) {
    check(nlk_zvec_axpby(a.real(), a.imag(), v.re.handle(), v.im.handle(), b.real(), b.imag(), re.handle(), im.handle()), "nek_zaxpby");
  }
  std::complex<double> dot(const nek_zvector& v) const {
    double a = 0, b = 0; check(nlk_zvec_dot(re.handle(), im.handle(), v.re.handle(), v.im.handle(), &a, &b), "nek_zdot"); return {a, b};
  }
  nek_dvector re, im;
};

// resolvent_linop (src/linops/neklab_linops.f90:198-205; src/linops/resolvent.f90): components omega and baseflow
class resolvent_linop {
 public:
  resolvent_linop(double omega_, const nek_dvector& baseflow) : omega(omega_), A_(1.0, baseflow) {}
  void matvec(const nek_zvector& vec_in, nek_zvector& vec_out) { apply(vec_in, vec_out, false); }      // resolvent.f90:17-46
  void rmatvec(const nek_zvector& vec_in, nek_zvector& vec_out) { apply(vec_in, vec_out, true); }      // :48-75
  double omega;
  double rtol = 0.0;                             // <= 0: the reference's GMRES tolerance 1e-6 (:122)
  int info = 0;

 private:
  void apply(const nek_zvector& i, nek_zvector& o, bool adjoint) {
    std::int32_t inf = 0;
    check(nlk_resolvent_matvec(A_.handle(), omega, i.re.handle(), i.im.handle(), o.re.handle(), o.im.handle(), adjoint ? 1 : 0, rtol, &inf), "resolvent_matvec");
    info = inf;
  }
  exptA_linop A_;
};

// ---------------------------------------------------------------------------------------------------- periodic orbits
// nek_ext_dvector (src/vectors/real_extended_vectors.f90): nek_dvector + period T (dot adds T*T', :243; axpby combines T, :193)
class nek_ext_dvector {
 public:
  explicit nek_ext_dvector(context& c, double T_ = 0.0) : vec(c), T(T_) {}
  void zero() { vec.zero(); T = 0.0; }
  void scal(double a) { vec.scal(a); T *= a; }
  void axpby(double alpha, const nek_ext_dvector& v, double beta) { vec.axpby(alpha, v.vec, beta); T = beta * T + alpha * v.T; }
  double dot(const nek_ext_dvector& v) const { return vec.dot(v.vec) + T * v.T; }
  double norm() const { return std::sqrt(dot(*this)); }
  std::int64_t get_size() const { return vec.get_size() + 1; }
  nek_dvector vec;
  double T;
};

// nek_upo_jacobian (src/systems/neklab_systems.f90:157-164; periodic_orbit.f90:46-181): linearisation about the orbit X = (X(0), T),
// base flow and perturbation advanced together on the device
class nek_upo_jacobian {
 public:
  nek_upo_jacobian(context& c, const nek_ext_dvector& X) : c_(&c), X_(&X) {}
  void matvec(const nek_ext_dvector& vec_in, nek_ext_dvector& vec_out) { apply(vec_in, vec_out, false); }
  void rmatvec(const nek_ext_dvector& vec_in, nek_ext_dvector& vec_out) { apply(vec_in, vec_out, true); }

 private:
  void apply(const nek_ext_dvector& i, nek_ext_dvector& o, bool transpose) {
    double T = 0;
    check(nlk_upo_jacobian(c_->handle(), X_->vec.handle(), X_->T, i.vec.handle(), i.T, o.vec.handle(), &T, transpose ? 1 : 0), "nek_upo_jacobian");
    o.T = T;
  }
  context* c_;
  const nek_ext_dvector* X_;
};

// ---------------------------------------------------------------------------------------------------- analysis drivers
struct eigs_result {
  std::vector<std::complex<double>> mu;        // Ritz values of exp(tau L)
  std::vector<std::complex<double>> eigvals;   // log(mu) / tau              (src/neklab_analysis.f90:84)
  std::vector<double> residuals;
  std::vector<nek_dvector> eigvecs;            // re, im, re, im, ...  (2*nev) when requested
  int niter = 0, info = 0;
};

namespace detail {
struct eigs_log { std::FILE* f; double tol; };
inline void eigs_cb(std::int32_t it, std::int32_t k, const double* re, const double* im, const double* res, void* user) {
  auto* l = static_cast<eigs_log*>(user);
  if (!l || !l->f) return;
  int b = 0;
  for (int i = 1; i < k; ++i) if (std::hypot(re[i], im[i]) > std::hypot(re[b], im[b])) b = i;
  std::fprintf(l->f, "%6d %18.10E %18.10E %18.10E %18.10E %s\n", it, re[b], im[b], std::hypot(re[b], im[b]), res[b], res[b] < l->tol ? "T" : "F");
}
}  // namespace detail

// src/neklab_analysis.f90:38-105.  Writes `eigs_output.txt` (the 6 columns test/lib/neklabTestCase.py:413-455 parses:
// iter, Re, Im, modulus, residual, T/F) into outdir when it is non-empty.  tol = 0 selects LightKrylov's default rtol_dp.
inline eigs_result linear_stability_analysis_fixed_point(exptA_linop& A, int kdim, int nev, bool adjoint = false, const std::string& outdir = "",
                                                         double tol = 0.0, const nek_dvector* x0 = nullptr, bool want_vectors = false) {
  eigs_result r;
  std::vector<double> re(nev), im(nev), res(nev);
  std::vector<nlk_vec*> vh;
  if (want_vectors) { for (int i = 0; i < 2 * nev; ++i) r.eigvecs.emplace_back(A.ctx()); for (auto& v : r.eigvecs) vh.push_back(v.handle()); }
  detail::eigs_log log{nullptr, tol > 0 ? tol : 3.1622776601683794e-08};
  if (!outdir.empty()) {
    log.f = std::fopen((outdir + "/eigs_output.txt").c_str(), "w");
    if (!log.f) throw error("cannot open " + outdir + "/eigs_output.txt");
    std::fprintf(log.f, "  iter            Re                 Im              modulus            residual     conv\n");
  }
  std::int32_t niter = 0, info = 0;
  const int rc = nlk_eigs(A.handle(), nev, kdim, tol, adjoint ? 1 : 0, x0 ? x0->handle() : nullptr, re.data(), im.data(), res.data(),
                          want_vectors ? vh.data() : nullptr, &niter, detail::eigs_cb, &log, &info);
  if (log.f) std::fclose(log.f);
  check(rc, "linear_stability_analysis_fixed_point");
  for (int i = 0; i < nev; ++i) {
    r.mu.emplace_back(re[i], im[i]);
    r.eigvals.push_back(std::log(r.mu.back()) / A.tau);
  }
  r.residuals = res; r.niter = niter; r.info = info;
  return r;
}

struct svds_result {
  std::vector<double> sigma, residuals;
  std::vector<nek_dvector> U, V;
  int niter = 0, info = 0;
};
// src/neklab_analysis.f90:107-156 (LightKrylov svds: Golub-Kahan with the direct and adjoint propagators)
inline svds_result transient_growth_analysis_fixed_point(exptA_linop& A, int nsv, int kdim, double tol = 0.0, const nek_dvector* x0 = nullptr) {
  svds_result r;
  r.sigma.resize(nsv); r.residuals.resize(nsv);
  std::vector<nlk_vec*> uh, vh;
  for (int i = 0; i < nsv; ++i) { r.U.emplace_back(A.ctx()); r.V.emplace_back(A.ctx()); }
  for (int i = 0; i < nsv; ++i) { uh.push_back(r.U[i].handle()); vh.push_back(r.V[i].handle()); }
  std::int32_t niter = 0, info = 0;
  check(nlk_svds(A.handle(), nsv, kdim, tol, x0 ? x0->handle() : nullptr, r.sigma.data(), r.residuals.data(), uh.data(), vh.data(), &niter, &info),
        "transient_growth_analysis_fixed_point");
  r.niter = niter; r.info = info;
  return r;
}

struct newton_result {
  std::vector<double> residual_history;
  int niter = 0, info = 0;
};
// src/neklab_analysis.f90:158-212: Newton-GMRES on F_tau(X) - X with the reference's tolerance schedulers
// (tol_mode 1 = nek_constant_tol, 2 = nek_dynamic_tol; src/systems/neklab_systems.f90:229-335).  X is updated in place.
inline newton_result newton_fixed_point_iteration(context& c, double tau, nek_dvector& X, double tol, int tol_mode = 2, int maxiter = 20, int gmres_kdim = 100) {
  newton_result r;
  r.residual_history.assign(maxiter + 1, 0.0);
  std::int32_t niter = 0, info = 0;
  check(nlk_newton_fixed_point(c.handle(), tau, X.handle(), tol, tol_mode, maxiter, gmres_kdim, r.residual_history.data(), &niter, &info),
        "newton_fixed_point_iteration");
  r.residual_history.resize(niter + 1);
  r.niter = niter; r.info = info;
  return r;
}

}  // namespace neklab
