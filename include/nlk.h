/* nlk.h -- C-ABI of the B200-native exptA hot path (neklab drop-in boundary).
 *
 * Every entry point is what a Fortran ISO_C_BINDING shim of neklab would bind for this
 * path.  For each group the reference interface it replaces is cited
 * as file:line under /root/reference.  All functions return 0 on success, non-zero on
 * error (message via nlk_last_error()); the reference aborts (`type_error`/`nek_stop_error`,
 * src/neklab_nek_setup.f90:406-417) where we return a status.
 *
 * Layout conventions: host arrays are double, element-major, x (r) fastest, exactly Nek's
 * (lx1,ly1,lz1,lelv) / (lx2,ly2,lz2,lelv) column-major storage.  Vectors/operators are opaque
 * handles owning device (HBM) buffers.  One host thread per nlk_ctx; not re-entrant (same as
 * the reference's COMMON-block state).  There is NO CPU fallback: any entry point that needs
 * the device fails with an error if no CUDA device is usable.
 */
#ifndef NLK_H
#define NLK_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nlk_mesh nlk_mesh;   /* host-side geometry + numbering (Nek setup: geom1/geom2/set_vert) */
typedef struct nlk_ctx nlk_ctx;     /* device-resident solver state (Nek COMMON blocks)                  */
typedef struct nlk_vec nlk_vec;     /* nek_dvector          src/vectors/neklab_vectors.f90:26-50         */
typedef struct nlk_op nlk_op;       /* exptA_linop          src/linops/neklab_linops.f90:35-44           */

const char* nlk_last_error(void);
int nlk_version(void);

/* ------------------------------------------------------------------ mesh (host only; no GPU needed)
 * Replaces what Nek5000 derives from SIZE + .re2/.ma2 at start-up (un-vendored);
 * the in-tree consumers are bm1/vmult/v?mask/glo_num users in src/vectors/real_vectors.f90:100-113,208-233. */
typedef struct {
  int32_t ndim;            /* ldim                                   SIZE:12 */
  int32_t lx1;             /* GLL points per direction               SIZE:13 */
  int32_t lxd;             /* dealiasing GL points                   SIZE:14 */
  int64_t nelg;            /* global element count                   SIZE:16 */
  int64_t nel;             /* local element count (== nelg if gllnid==NULL) */
  const double* xm1;       /* [nel][lx1^ndim] local GLL coordinates */
  const double* ym1;
  const double* zm1;       /* NULL in 2-D */
  const int64_t* vertex;   /* [nelg][2^ndim] .ma2 global vertex ids, lexicographic corners, ALL elements */
  const char* cbc_v;       /* [nelg][2*ndim][3] velocity boundary codes of ALL elements (preprocessor face order): Dirichlet
                              flags of shared vertices/edges must be known on every rank that touches them */
  const char* cbc_t;       /* [nelg][2*ndim][3] temperature codes, or NULL */
  const int32_t* gllnid;   /* [nelg] owning rank of every element (Nek gllnid), or NULL = single rank */
  int32_t rank, nranks;
} nlk_mesh_desc;

typedef struct {
  int32_t ndim, lx1, lx2, lxd;
  int64_t nel, nelg;
  int64_t np1, np2;        /* points per element on mesh 1 / mesh 2 */
  int64_t nglob_local;     /* distinct global nodes touching this rank */
  int64_t nshared_local;   /* distinct global nodes with >1 local copies */
  int64_t nvert;           /* global vertex count */
  int32_t has_outflow;     /* pressure operator non-singular */
  int32_t nneigh;          /* neighbour ranks in the gather-scatter exchange */
  double volvm1, volvm2;   /* local sums of bm1 / bm2 */
} nlk_mesh_info_t;

int nlk_partition(const int64_t* pid, int64_t nelg, int32_t nranks, int32_t* gllnid); /* Nek power-of-two rule on .ma2 leaf ids */
int nlk_mesh_create(const nlk_mesh_desc* d, nlk_mesh** out);
int nlk_mesh_destroy(nlk_mesh* m);
int nlk_mesh_info(const nlk_mesh* m, nlk_mesh_info_t* out);
int nlk_mesh_glo_num(const nlk_mesh* m, int64_t* glo /* [nel][np1], 1-based */);
/* name in: bm1 jac binvm1 vmult vmask0..2 tmask bm2 g11 g12 g13 g22 g23 g33 (mesh-1 sized, bm2 mesh-2 sized) */
int nlk_mesh_field(const nlk_mesh* m, const char* name, double* out);
/* gather-scatter exchange plan with neighbour rank index k (0..nneigh-1): rank id, count, global node ids (sorted) */
int nlk_mesh_neighbor(const nlk_mesh* m, int32_t k, int32_t* rank, int64_t* count, int64_t* gids /* may be NULL */);
/* 1-D operators: name in z1 w1 z2 w2 zd wd D I12 D12 I1d Dd ; returns count written */
int nlk_mesh_basis(const nlk_mesh* m, const char* name, double* out, int64_t cap);

/* ------------------------------------------------------------------ small dense host kernels (LAPACK-free)
 * replaces LightKrylov's `eig`/`schur` calls through stdlib_linalg (src/neklab_otd.f90:229,248). */
int nlk_dense_eig(int32_t n, const double* A /* row-major n*n */, double* wr, double* wi,
                  double* VR /* n*n row-major; complex pairs stored LAPACK dgeev style in columns */);

/* ------------------------------------------------------------------ solver context (device)
 * replaces Nek's param()/COMMON state poked by setup_nek (src/neklab_nek_setup.f90:39-247). */
typedef struct {
  double viscosity;        /* param(2): h1 = 1/Re                      .par [VELOCITY] viscosity */
  double density;          /* param(1)                                 .par [VELOCITY] density   */
  int32_t torder;          /* |param(27)|, 1..3                        .par timeStepper          */
  double vtol;             /* param(22) -> TOLHDF   src/neklab_nek_setup.f90:228 */
  double ptol;             /* param(21) -> TOLPDF   src/neklab_nek_setup.f90:227 */
  int32_t ifheat;          /* temperature coupled (exptA_temp_linop)   */
  double conductivity, rhocp, ttol;
  double buoyancy[3];      /* f_c += buoyancy[c]*T'  (examples/rayBen/baseflow/rayBen.usr:75-105) */
  double filter_weight;    /* param(103)            .par filterWeight */
  double filter_cutoff;    /* filterCutoffRatio */
  int32_t cg_maxit;        /* Helmholtz CG cap */
  int32_t gmres_maxit;     /* Nek: 100 */
  int32_t lgmres;          /* SIZE lgmres = 30 */
  int32_t precond;         /* 0 = mass-scaled identity, 1 = Schwarz + coarse (semg_xxt analogue; dense inverse up to 5000
                              vertices, sparse CSR + Jacobi-PCG above), 2 = Schwarz only, 3 = coarse only, 4 = Schwarz + sparse coarse */
  int32_t pr_proj;         /* residualProj: size of the pressure projection space (0 = off, Nek mxprev=20) */
  double cfl_limit;        /* 0.5 for the linear solver (src/linops/exponential_propagator.f90:12) */
  int32_t rst_mode;        /* 0 (default) = restart-field arithmetic exactly as written in src/vectors/real_vectors.f90:186-200
                              (rst slots receive alpha * the CURRENT fields of the other vector);
                              1 = consistent combination of the rst fields; 2 = matvec ignores input rst fields.
                              1/2 are NOT the reference behaviour; they exist to document its effect */
  int32_t coarse_iters;    /* Jacobi-PCG iterations of the sparse coarse solve used above 5000 vertices (crs_solve analogue) */
  int32_t step_variant;    /* 0 (default). Bit field of documented deviations in the start-up / pressure bookkeeping of the step,
                              used ONLY by the golden-value sweep of DESIGN.md 1.1 (tests/test_gpu_golden_sweep.py):
                              1 = prlagp is never updated (stays 0), 2 = dp is added to prp instead of to the extrapolated p*,
                              4 = the pressure of vec_in is ignored (prp = 0 at the start of every matvec),
                              8 = first-order pressure extrapolation on every step (p* = p^n) */
} nlk_params;

int nlk_params_default(nlk_params* p);
int nlk_ctx_create(const nlk_mesh* m, const nlk_params* p, int32_t device, nlk_ctx** out);
int nlk_ctx_destroy(nlk_ctx* c);
int nlk_ctx_set_tol(nlk_ctx* c, double vtol, double ptol);           /* setup_nek vtol/ptol args */
int nlk_ctx_set_dt(nlk_ctx* c, double dt);                           /* Nek `dt` (param(12)) used when the base flow is zero: setup_nek disables recompute_dt (:79-83) */
/* multi-rank: attach an NCCL communicator built from a broadcast unique id (128 bytes) */
int nlk_comm_unique_id(char id[128]);
int nlk_ctx_comm_init(nlk_ctx* c, const char id[128], int32_t rank, int32_t nranks);
int nlk_ctx_sync(nlk_ctx* c);                                        /* stream synchronize */
void* nlk_ctx_stream(nlk_ctx* c);                                    /* cudaStream_t */

/* ------------------------------------------------------------------ nek_dvector
 * src/vectors/neklab_vectors.f90:64-113 (TBP interfaces), src/vectors/real_vectors.f90 (bodies). */
int nlk_vec_create(nlk_ctx* c, nlk_vec** out);
int nlk_vec_destroy(nlk_vec* v);
int nlk_vec_copy(nlk_vec* dst, const nlk_vec* src);                  /* Fortran assignment / allocate(source=) */
int nlk_vec_zero(nlk_vec* v);                                        /* nek_dzero  real_vectors.f90:37-50   */
int nlk_vec_rand(nlk_vec* v, int32_t ifnorm, uint64_t seed);         /* nek_drand  :52-123 (seeded, C0, BC-satisfying) */
int nlk_vec_scal(nlk_vec* v, double alpha);                          /* nek_dscal  :125-160 */
int nlk_vec_axpby(double alpha, const nlk_vec* x, double beta, nlk_vec* self); /* nek_daxpby :162-206: self = alpha*x + beta*self (+rst quirk) */
int nlk_vec_dot(const nlk_vec* self, const nlk_vec* x, double* out); /* nek_ddot   :208-233 (bm1-weighted, no pressure) */
int nlk_vec_norm(const nlk_vec* self, double* out);
int nlk_vec_size(const nlk_vec* v, int64_t* n);                      /* nek_dsize  :235-247 */
int nlk_vec_save_rst(nlk_vec* self, const nlk_vec* state, int32_t irst); /* dsave_rst :249-291 */
int nlk_vec_get_rst(const nlk_vec* self, nlk_vec* out, int32_t irst);    /* dget_rst  :293-333 */
int nlk_vec_nrst(const nlk_vec* v, int32_t* nrst);                   /* dhas_rst_fields :335-338 */
int nlk_vec_clear_rst(nlk_vec* v);                                   /* dclear_rst_fields :340-346 */
/* nek2vec / vec2nek (src/neklab_utils.f90:84-134): host <-> device marshalling; NULL pointers are skipped */
int nlk_vec_upload(nlk_vec* v, const double* vx, const double* vy, const double* vz, const double* pr, const double* theta);
int nlk_vec_download(const nlk_vec* v, double* vx, double* vy, double* vz, double* pr, double* theta);

/* nek_zvector (src/vectors/neklab_vectors.f90:219-237; src/vectors/complex_vectors.f90:72-110): a complex vector is the
 * pair (re, im) of nek_dvector handles; the three non-trivial TBPs are provided fused (no scratch copy of the vector,
 * unlike nek_zscal/nek_zaxpby which deep-copy a temporary). */
int nlk_zvec_scal(nlk_vec* re, nlk_vec* im, double alpha_re, double alpha_im);                       /* nek_zscal  :72-80  */
int nlk_zvec_axpby(double alpha_re, double alpha_im, const nlk_vec* xre, const nlk_vec* xim,
                   double beta_re, double beta_im, nlk_vec* sre, nlk_vec* sim);                        /* nek_zaxpby :82-98  */
int nlk_zvec_dot(const nlk_vec* sre, const nlk_vec* sim, const nlk_vec* xre, const nlk_vec* xim,
                 double* out_re, double* out_im);                                                      /* nek_zdot   :100-110: conj(self).x */

/* block kernels replacing LightKrylov innerprod / linear_combination / double_gram_schmidt_step */
int nlk_basis_innerprod(nlk_vec* const* X, int32_t k, const nlk_vec* y, double* h);
int nlk_basis_axpy(nlk_vec* y, nlk_vec* const* X, int32_t k, const double* c);   /* y += X c (current fields; rst per quirk) */
int nlk_basis_dgs(nlk_vec* y, nlk_vec* const* X, int32_t k, double* h, double* norm_out);

/* ------------------------------------------------------------------ exptA_linop
 * src/linops/neklab_linops.f90:35-75 ; src/linops/exponential_propagator.f90 (init :4-13, matvec :15-60,
 * rmatvec :62-107, compute_rst :109-127, get_rst :129-142); exptA_temp_linop: exponential_propagator_temp.f90. */
typedef struct {
  int32_t nsteps;
  double dt;
  int64_t cg_iters, gmres_iters, steps, matvecs;
  double ms_total;         /* device time of the last matvec (CUDA events) */
  int64_t launches;        /* kernel launches issued by the last matvec */
} nlk_stats;

int nlk_exptA_create(nlk_ctx* c, double tau, const nlk_vec* baseflow, nlk_op** out);
int nlk_exptA_destroy(nlk_op* op);
int nlk_exptA_init(nlk_op* op);
int nlk_exptA_set_tau(nlk_op* op, double tau);                        /* apply_exptA :224-266 */
int nlk_exptA_matvec(nlk_op* op, const nlk_vec* in, nlk_vec* out);
int nlk_exptA_rmatvec(nlk_op* op, const nlk_vec* in, nlk_vec* out);
int nlk_exptA_stats(const nlk_op* op, nlk_stats* out);
/* exptA_proj_linop (src/linops/neklab_linops.f90:130-152; exponential_propagator_proj.f90): exptA with a projection of the
 * velocity onto the streamwise wavenumber alpha, u <- cos(alpha x) <2 u cos(alpha x)> + sin(alpha x) <2 u sin(alpha x)>
 * (`proj_alpha` :135-173; <.> = Nek `planar_avg`, the bm1-weighted average over the points that share their transverse
 * coordinates), applied to the initial condition and to the state at tau in matvec AND rmatvec (:46-47, :64-65).
 * idir = 1, 2, 3 is the averaged direction (the reference's gtpp_gs_setup idir); idir = 0 switches the projection off.
 * Single rank only.  nlk_exptA_apply_projection applies `proj_alpha` to a vector (self%proj on the solver state). */
int nlk_exptA_set_projection(nlk_op* op, double alpha, int32_t idir);
int nlk_exptA_apply_projection(nlk_op* op, nlk_vec* v);
/* bench hook: start from `in` (history reset as at the top of exptA_matvec), run nwarm untimed + nsteps timed perturbation
 * time steps (body of the loop at exponential_propagator.f90:39-46) and return the device time of the timed ones */
int nlk_exptA_time_steps(nlk_op* op, const nlk_vec* in, int32_t nwarm, int32_t nsteps, double* ms_timed);
/* nek2vec / vec2nek on the DEVICE-resident solver state (src/neklab_utils.f90:84-134): copy the perturbation fields
 * vxp,vyp,vzp,prp,tp of the context into / out of a vector (intent(out) semantics: nrst of the vector is reset) */
int nlk_nek2vec(nlk_ctx* c, nlk_vec* out);
int nlk_vec2nek(nlk_ctx* c, const nlk_vec* in);
int nlk_exptA_set_baseflow(nlk_op* op, const nlk_vec* baseflow);      /* nek_jacobian%X = X (src/systems/neklab_systems.f90) */
/* nek_system%response (src/systems/fixed_point.f90:4-40): out = F_tau(in) - in with the NONLINEAR stepper
 * (Nek `fluid`/`plan3`: makef/advab, cresvif, ophinv, incomprn), dt from the CFL of `in` at cfl_limit (0.4 in the reference) */
int nlk_nonlinear_map(nlk_ctx* c, double tau, double cfl_limit, const nlk_vec* in, nlk_vec* out);
/* newton_fixed_point_iteration (src/neklab_analysis.f90:158-212) with gmres on nek_jacobian = exptA - I
 * (src/systems/fixed_point.f90:42-96) and the tolerance schedulers nek_constant_tol / nek_dynamic_tol
 * (src/systems/neklab_systems.f90:229-335; tol_mode 1 / 2).  X is updated in place; rnorm_hist gets niter+1 values. */
int nlk_newton_fixed_point(nlk_ctx* c, double tau, nlk_vec* X, double tol, int32_t tol_mode, int32_t maxiter, int32_t gmres_kdim,
                           double* rnorm_hist, int32_t* niter, int32_t* info);
/* neklab forcing registry (src/neklab_nek_forcing.f90): module arrays neklab_ffx/y/z(lv, lpert+1) added to the body force
 * inside `userf` through `neklab_forcing(ffx,ffy,ffz,ix,iy,iz,ieg,ipert)` (:96-114).  ipert = 0 is the slot of the nonlinear
 * solver, ipert = 1 the slot of the perturbation (lpert = 1); any other value is an error (the reference calls nek_end, :44-47).
 * Host arrays [nel][lx1^ndim]; NULL components are treated as zero (set) or skipped (get). */
int nlk_set_neklab_forcing(nlk_ctx* c, const double* fx, const double* fy, const double* fz, int32_t ipert);   /* :57-76  */
int nlk_get_neklab_forcing(nlk_ctx* c, double* fx, double* fy, double* fz, int32_t ipert);                     /* :36-55  */
int nlk_zero_neklab_forcing(nlk_ctx* c);                                                                       /* :25-34  */
int nlk_zero_neklab_forcing_ipert(nlk_ctx* c, int32_t ipert);                                                  /* :78-94  */
int nlk_ctx_set_forcing(nlk_ctx* c, const double* fx, const double* fy, const double* fz);                     /* = set_neklab_forcing(..., ipert = 1) */

/* ------------------------------------------------------------------ analysis entry points (device-resident bases)
 * src/neklab_analysis.f90:38-105 (eigs), :107-156 (svds), :158-212 (newton/gmres). */
typedef void (*nlk_eigs_cb)(int32_t iter, int32_t k, const double* lam_re, const double* lam_im,
                            const double* resid, void* user);
int nlk_eigs(nlk_op* op, int32_t nev, int32_t kdim, double tol, int32_t transpose, const nlk_vec* x0,
             double* lam_re, double* lam_im, double* resid, nlk_vec** eigvecs /* 2*nev (re,im) or NULL */,
             int32_t* niter, nlk_eigs_cb cb, void* user, int32_t* info);
int nlk_svds(nlk_op* op, int32_t nsv, int32_t kdim, double tol, const nlk_vec* x0, double* sigma, double* resid,
             nlk_vec** U, nlk_vec** V, int32_t* niter, int32_t* info);
int nlk_gmres(nlk_op* op, int32_t minus_identity, const nlk_vec* b, nlk_vec* x, int32_t kdim, double atol, double rtol,
              int32_t maxiter, int32_t transpose, int32_t* info);

/* resolvent_linop%matvec / %rmatvec  (/root/reference/src/linops/resolvent.f90:17-75) on a nek_zvector (re, im):
 *   tau = 2 pi / |omega| (1 if omega == 0);  b = forced response from rest over tau (evaluate_rhs, :80-112);
 *   re  = (I - exptA(tau))^-1 b by GMRES(kdim 64, atol 1e-12, rtol 1e-6, transpose = adjoint) (:114-134);
 *   im  = forced integration over tau/4 from re (evaluate_imaginary_part, :136-166).
 * The harmonic forcing Re(exp(+-i omega t) f) goes through the perturbation slot of the forcing registry and the
 * registry is zeroed afterwards, as in the reference.  op supplies the base flow; its tau is restored on return.
 * rtol <= 0 selects the reference's 1e-6.  info: 0 = GMRES converged. */
int nlk_resolvent_matvec(nlk_op* op, double omega, const nlk_vec* f_re, const nlk_vec* f_im, nlk_vec* out_re, nlk_vec* out_im,
                         int32_t adjoint, double rtol, int32_t* info);
/* one forced integration of the above (evaluate_rhs when x0 == NULL, evaluate_imaginary_part otherwise) */
int nlk_resolvent_integrate(nlk_op* op, double tau, double omega, const nlk_vec* f_re, const nlk_vec* f_im, int32_t adjoint,
                            const nlk_vec* x0, nlk_vec* out);

/* nek_upo_jacobian%matvec / %rmatvec  (/root/reference/src/systems/periodic_orbit.f90:46-113 / :115-181) on
 * nek_ext_dvector = (nek_dvector fields, period T) (/root/reference/src/vectors/real_extended_vectors.f90):
 *   base flow := X, dt from its CFL at 0.4 with endtime T_X, tolerances atol*0.1 (direct) / atol*0.5 (adjoint), atol = vtol;
 *   base flow and perturbation advanced TOGETHER (setup_linear_solver(solve_baseflow = .true.)) with the rst protocol of exptA;
 *   out = u'(T_X) - in + T_in * f'(X(T)),  T_out = <in, f'(X(0))>,  f' = (F_dt(X) - X)/dt (compute_fdot,
 *   /root/reference/src/systems/neklab_systems.f90:202-223).  On return vtol = ptol = atol as in the reference (:106-107).
 * The nonlinear map of nek_upo_system (:4-44) is nlk_nonlinear_map with tau = T (the period component of its result is 0). */
int nlk_upo_jacobian(nlk_ctx* c, const nlk_vec* X, double T_X, const nlk_vec* in, double T_in, nlk_vec* out, double* T_out,
                     int32_t transpose);

/* ------------------------------------------------------------------ kernel-level test/bench hooks
 * (the K-numbered kernels of SURVEY §2.3; host arrays in, host arrays out unless noted) */
int nlk_test_axhelm(nlk_ctx* c, const double* u, double h1, double h2, double* w);          /* K1, local (no dssum) */
int nlk_test_dssum(nlk_ctx* c, double* u);                                                   /* K2 */
int nlk_test_opdiv(nlk_ctx* c, const double* ux, const double* uy, const double* uz, double* p);   /* K5 */
int nlk_test_opgradt(nlk_ctx* c, const double* p, double* wx, double* wy, double* wz);       /* K5 */
int nlk_test_cdabdtp(nlk_ctx* c, const double* p, double* ep);                               /* K6 */
int nlk_test_convect(nlk_ctx* c, const double* u, const double* cx, const double* cy, const double* cz, double* out); /* K3 */
int nlk_test_convect_adj(nlk_ctx* c, const double* const* U, const double* const* cf, double* const* out);            /* K4 */
int nlk_test_helmholtz(nlk_ctx* c, const double* f, double h1, double h2, int32_t comp, double tol, double* x, int32_t* iters); /* K8 */
int nlk_test_pressure(nlk_ctx* c, const double* rhs, double tol, double* x, int32_t* iters); /* K9-K11 */
int nlk_test_precond(nlk_ctx* c, const double* r, double* z);                                /* K10/K11 */
int nlk_test_cfl(nlk_ctx* c, const double* ux, const double* uy, const double* uz, double dt, double* cfl); /* K14 */
/* time a device-resident kernel nrep times on the ctx stream with CUDA events; returns mean ms per launch.
 * which: 0 axhelm, 1 dssum, 2 cdabdtp, 3 convect(all comps), 4 precond, 5 vec dot, 6 sparse coarse solve, 7 Schwarz branch of the preconditioner,
 * 8 the FUSED Helmholtz apply the Jacobi-PCG launches (p = hd r + beta p on load, p.Ap on the way out), 9 the PCG update+reduce kernel,
 * 10 opgradt alone, 11 opdiv (fused load scaling) alone, 12 dssum of three fields */
int nlk_bench_kernel(nlk_ctx* c, int32_t which, int32_t nrep, double* ms_per_launch, double* algo_bytes);

#ifdef __cplusplus
}
#endif
#endif
