// Solver context: device-resident equivalent of Nek5000's COMMON-block state for one perturbation (lpert = 1),
// plus NCCL plumbing.  See SURVEY.md §8b "Global state".
#pragma once
#include "nlk_device.cuh"
#include "../../include/nlk.h"

struct nlk_mesh { nlk::HostMesh hm; };

namespace nlk {

// ---- NCCL through dlopen (libnccl.so.2: the copy torch already loaded, else the system one)
struct NcclId { char internal[128]; };
struct Nccl {
  void* lib = nullptr;
  void* comm = nullptr;
  void* comm2 = nullptr;                 // ncclCommSplit copy used by the side stream (coarse branch of the preconditioner)
  int rank = 0, nranks = 1;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId /*ncclUniqueId by value*/, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*CommSplit)(void*, int, int, void**, void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
int nccl_load(Nccl& n);

struct DevNeighbor {
  int rank; int cnt;
  int32_t* rep;          // [cnt] representative local index
  int32_t* cp_off;       // [cnt+1]
  int32_t* cp_idx;       // local copies of every shared node
  double* sendbuf;       // [3*cnt]
  double* recvbuf;
};

}  // namespace nlk

// Phase timing of the time step (diagnostics; enabled by the environment variable NLK_PHASES=1, printed to stderr when the
// context is destroyed): CUDA-event pairs recorded on the library stream around each phase and resolved once per step.
namespace nlk {
enum Phase { PH_MAKEF, PH_VRES, PH_HELM, PH_PRHS, PH_PRES, PH_PRECOND, PH_EAPPLY, PH_ORTH, PH_CORR, PH_HEAT, PH_FILTER, PH_COUNT };
struct PhaseTimer {
  bool on = false;
  struct Rec { int id; cudaEvent_t a, b; };
  std::vector<Rec> recs; std::vector<cudaEvent_t> pool;
  double ms[PH_COUNT] = {0}; long calls[PH_COUNT] = {0}; long steps = 0;
  cudaEvent_t get() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
  int begin(int id, cudaStream_t st) { if (!on) return -1; Rec r{id, get(), get()}; cudaEventRecord(r.a, st); recs.push_back(r); return (int)recs.size() - 1; }
  void end(int idx, cudaStream_t st) { if (idx >= 0) cudaEventRecord(recs[idx].b, st); }
  void resolve(cudaStream_t st) {
    if (!on) return;
    cudaStreamSynchronize(st);
    for (auto& r : recs) { float t = 0; cudaEventElapsedTime(&t, r.a, r.b); ms[r.id] += t; ++calls[r.id]; pool.push_back(r.a); pool.push_back(r.b); }
    recs.clear(); ++steps;
  }
};
struct PhaseScope {
  PhaseTimer& t; cudaStream_t st; int idx;
  PhaseScope(PhaseTimer& t_, int id, cudaStream_t st_) : t(t_), st(st_), idx(t_.begin(id, st_)) {}
  ~PhaseScope() { t.end(idx, st); }
};
}  // namespace nlk

namespace nlk {
// second copy of the time-stepping state: when the base flow is advanced together with the perturbation (Nek `ifbase`,
// setup_linear_solver(solve_baseflow = .true.), src/neklab_nek_setup.f90:105-106) the nonlinear state and its BDF/EXT history
// live here and are swapped with the active set around the nonlinear step
struct StateBank {
  double* vp[3] = {nullptr, nullptr, nullptr}; double* prp = nullptr; double* tp = nullptr;
  double* vlag[2][3] = {{nullptr}}; double* exx1[3] = {nullptr}; double* exx2[3] = {nullptr};
  double* prlag = nullptr; double* proj_X = nullptr; double* proj_EX = nullptr; int nproj = 0;
};
}  // namespace nlk

struct nlk_ctx {
  nlk::PhaseTimer ph;
  nlk::StateBank* bank2 = nullptr;
  const nlk_mesh* mesh = nullptr;
  nlk::DevMesh dm{};
  nlk_params prm{};
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;            // side stream: the coarse-grid branch of the preconditioner runs beside the Schwarz branch
  cudaEvent_t ev_in = nullptr, ev_crs = nullptr;
  int device = 0;
  nlk::Nccl nccl;
  std::vector<nlk::DevNeighbor> neigh;
  std::vector<void*> allocs;             // everything cudaMalloc'ed, freed in destroy
  // scalars
  nlk::SolverScal* d_sc = nullptr;       // device
  nlk::SolverScal* h_sc = nullptr;       // pinned mirror
  double* d_red = nullptr;               // device scratch scalars [64 + lgmres]
  double* h_red = nullptr;               // pinned mirror (UVA: device kernels can write it directly)
  unsigned int* h_seq = nullptr; unsigned int seq = 0;   // publish sequence number of the zero-copy scalar read-back
  nlk::Reducer red{};
  // coordinates (rand field) and misc
  double* xyz[3] = {nullptr, nullptr, nullptr};
  int64_t* d_lglel = nullptr;
  double* ones2 = nullptr;               // N2 ones
  double* filterF = nullptr;             // n x n
  // base flow and perturbation state (Nek vx..t / vxp..tp)
  double* U[3] = {nullptr, nullptr, nullptr};
  double* T = nullptr;
  double* vp[3] = {nullptr, nullptr, nullptr};
  double* prp = nullptr;
  double* tp = nullptr;
  double* vlag[2][3] = {{nullptr}};
  double* exx1[3] = {nullptr}, *exx2[3] = {nullptr};
  double* prlag = nullptr;
  double* tlag[2] = {nullptr, nullptr};
  double* vgradt1 = nullptr, *vgradt2 = nullptr;
  double* bf[3] = {nullptr}, *bq = nullptr;
  double* forcing[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // neklab_ffx/y/z(lv, lpert+1): slot 0 nonlinear solver, slot 1 perturbation
  bool has_forcing[2] = {false, false};
  // work
  double* wk[8] = {nullptr};             // N1 each
  double* cg_x = nullptr, *cg_r = nullptr, *cg_p = nullptr, *cg_w = nullptr;
  double* cgm_x[3] = {nullptr}, *cgm_p[3] = {nullptr}, *cgm_w[3] = {nullptr};   // persistent multi-component PCG
  int* d_cg_iters = nullptr; unsigned long long* d_cg_total = nullptr; unsigned long long cg_total_seen = 0; bool use_cgp = false;
  double* pw[6] = {nullptr};             // N2 each
  double* gm_V = nullptr, *gm_Z = nullptr;   // (lgmres+1) x N2, lgmres x N2
  double* sw_w = nullptr, *sw_z = nullptr, *sw_t = nullptr;   // Schwarz work (N1)
  double* crs_part = nullptr, *crs_r = nullptr, *crs_y = nullptr;
  bool have_coarse = false, have_schwarz = false;
  // sparse coarse operator (large vertex counts): CSR of R0 E R0^T + Jacobi-PCG work vectors, all device-resident
  bool coarse_sparse = false; int crs_iters = 0; int64_t crs_nnz = 0;
  int32_t* crs_rowptr = nullptr; int32_t* crs_col = nullptr; double* crs_val = nullptr; double* crs_dinv = nullptr;
  double* crs_p = nullptr, *crs_q = nullptr, *crs_z = nullptr, *crs_rr = nullptr;
  double* cg_hd[4] = {nullptr, nullptr, nullptr, nullptr}, *cg_wa[4] = {nullptr, nullptr, nullptr, nullptr}, *cg_wb[4] = {nullptr, nullptr, nullptr, nullptr};
  double cg_key_h1[4] = {0, 0, 0, 0}, cg_key_h2[4] = {0, 0, 0, 0};           // (h1, h2) the weights of mask slot k were built for (lazily allocated)
  void* swf = nullptr;                      // fused Schwarz branch: pull tables + compact face buffers (nlk_schwarz.cu), released by swf_release
  double* cg_pap_partial = nullptr; unsigned int* cg_pap_counter = nullptr;   // block partials of p.Ap reduced inside the Helmholtz kernel
  double* crs_partial = nullptr;         // per-block partial sums of the coarse PCG (own scratch: the solve runs on the side stream)
  cudaGraphExec_t crs_graph = nullptr; const double* crs_graph_in = nullptr; double* crs_graph_out = nullptr;   // the fixed-count PCG as one graph launch
  // pressure projection (residualProj)
  double* proj_X = nullptr, *proj_EX = nullptr, *proj_w = nullptr, *proj_xbar = nullptr; int nproj = 0;
  // time stepping
  double dt = 0; int nsteps = 0; bool adjoint = false; bool nonlinear = false;
  // statistics
  long cg_iters = 0, gmres_iters = 0, steps = 0;
};

struct nlk_vec {
  nlk_ctx* c = nullptr;
  double* v[3] = {nullptr, nullptr, nullptr};
  double* pr = nullptr;
  double* theta = nullptr;
  int nrst = 0;
  double* rv[2][3] = {{nullptr}};        // rst slots allocated lazily
  double* rpr[2] = {nullptr, nullptr};
  double* rth[2] = {nullptr, nullptr};
};

struct nlk_op {
  nlk_ctx* c = nullptr;
  double tau = 1.0;
  nlk_vec* baseflow = nullptr;           // owned copy
  nlk_stats stats{};
  // exptA_proj_linop: planar-average projection onto one streamwise wavenumber
  int proj_dir = 0; double proj_alpha = 0; int64_t proj_ngroups = 0;
  int32_t *proj_off = nullptr, *proj_idx = nullptr, *proj_gid = nullptr; double *proj_cv = nullptr, *proj_sv = nullptr, *proj_coef = nullptr;
};

namespace nlk {
// internal API shared between translation units
int ctx_gs(nlk_ctx* c, Ptr3 f, int nf);                                   // full dssum (local + neighbour exchange)
int ctx_allreduce(nlk_ctx* c, double* d_ptr, int count, bool maxop);
int ctx_read_scalars(nlk_ctx* c, int count);                              // d_red[0..count) -> h_red, synchronises
int vec_alloc_rst(nlk_vec* v, int slot);
void vec_release_rst(nlk_vec* v);                   // free the rst slots (nrst = 0)
int exptA_apply(nlk_op* op, const nlk_vec* in, nlk_vec* out, bool transpose);
int exptA_project(nlk_op* op, double* const v[3]);
int step_setup(nlk_ctx* c, double tau, bool transpose);
int step_setup_cfl(nlk_ctx* c, double tau, double cfl_limit, CPtr3 u);
int step_advance(nlk_ctx* c, int istep);
int helmholtz_solve(nlk_ctx* c, double* rhs_local, double h1, double h2, const double* mask, double tol, double* x, int* iters);
int cg_weights(nlk_ctx* c, const double* mask, double h1, double h2, int* slot_out);
// fused Schwarz branch of the pressure preconditioner (nlk_schwarz.cu)
int swf_setup(nlk_ctx* c);
void swf_release(nlk_ctx* c);
int swf_apply(nlk_ctx* c, const double* r, const double* in_mul, const double* yc, double* z, cudaEvent_t wait_before_b = nullptr, int phase = 0);   // phase 1: kernel A (+ exchanges) only, 2: kernel B only
int helmholtz_solve_multi(nlk_ctx* c, int nf, double* const* rhs_local, double h1, double h2, const double* const* masks, double tol, double* const* sol);
int sync_cg_counter(nlk_ctx* c);
int pressure_solve(nlk_ctx* c, const double* rhs, double tol, double* x, int* iters);
int apply_E(nlk_ctx* c, const double* p, double* ep, const double* out_mul);
int apply_precond(nlk_ctx* c, const double* r, double* z, const double* in_mul);
int pressure_solve_projected(nlk_ctx* c, double* rhs, double tol, double* x, int* iters);
int coarse_setup_sparse(nlk_ctx* c);
int coarse_solve_sparse(nlk_ctx* c, const double* rc, double* yc);
int ortho(nlk_ctx* c, double* p);
int reset_history_pub(nlk_ctx* c);
int bank2_ensure(nlk_ctx* c);                       // allocate the second state set (lazily, first coupled run)
void bank_swap(nlk_ctx* c);                          // active state <-> bank2
int coupled_advance(nlk_ctx* c, int istep);          // one nek_advance with ifbase: perturbation step on U(t^{n-1}), then the nonlinear step of U
void make_filter_matrix(const Basis& b, double w, double cutoff, std::vector<double>& F);
void make_fdm_1d(const Basis& b, double lm, double ll, double lr, int bcl, int bcr, double* S, double* lam, int* nact);
double mesh_diag_local(const HostMesh& hm, int64_t e, int p);
std::vector<double> mat_transpose(const std::vector<double>& M, int r, int cdim);
template <class T> int dev_alloc(nlk_ctx* c, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T));
  if (e != cudaSuccess) { set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); return 1; }
  if (zero) cudaMemsetAsync(q, 0, count * sizeof(T), c->st);
  c->allocs.push_back(q); *p = (T*)q; return 0;
}
template <class T> int dev_upload(nlk_ctx* c, T** p, const std::vector<T>& h) {
  if (dev_alloc(c, p, h.size(), false)) return 1;
  if (!h.empty()) NLK_CUDA(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->st));
  NLK_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}
}  // namespace nlk
