! neklab_b200.f90 -- Fortran 2008 shim of the libnlk C-ABI (include/nlk.h) carrying neklab's own type and procedure names.
!
! What a neklab maintainer adds to bind the B200 path: this ONE file replaces src/vectors/real_vectors.f90,
! src/linops/exponential_propagator.f90 (+ _temp) and the driver bodies of src/neklab_analysis.f90 for the exptA hot path.
! `nek_dvector` keeps the deferred procedures of LightKrylov's `abstract_vector_rdp`
! (zero, rand, scal, axpby, dot, get_size; src/vectors/neklab_vectors.f90:64-93) and neklab's extras
! (save_rst, get_rst, has_rst_fields, clear_rst_fields; :95-113); the static arrays vx, vy, vz, pr, theta (+ rst copies)
! become one opaque handle on device-resident buffers.  `exptA_linop` keeps init / matvec / rmatvec
! (src/linops/neklab_linops.f90:35-75) and the `tau` + `baseflow` components.
!
! NOT COMPILED IN THIS REPOSITORY: the image has no Fortran compiler (SURVEY.md section 0).  The same entry points are
! exercised through ctypes (neklab_b200/api.py) and through the compiled C++ mirror include/neklab.hpp; every interface
! below is a mechanical transcription of a prototype in include/nlk.h (same order).  Link: -lnlk -lcudart.
module nlk_c
   use iso_c_binding
   implicit none
   public

   type, bind(C) :: nlk_mesh_desc
      integer(c_int32_t) :: ndim, lx1, lxd
      integer(c_int64_t) :: nelg, nel
      type(c_ptr) :: xm1, ym1, zm1, vertex, cbc_v, cbc_t, gllnid
      integer(c_int32_t) :: rank, nranks
   end type
   type, bind(C) :: nlk_params
      real(c_double) :: viscosity, density
      integer(c_int32_t) :: torder
      real(c_double) :: vtol, ptol
      integer(c_int32_t) :: ifheat
      real(c_double) :: conductivity, rhocp, ttol, buoyancy(3), filter_weight, filter_cutoff
      integer(c_int32_t) :: cg_maxit, gmres_maxit, lgmres, precond, pr_proj
      real(c_double) :: cfl_limit
      integer(c_int32_t) :: rst_mode, coarse_iters, step_variant
   end type
   type, bind(C) :: nlk_stats
      integer(c_int32_t) :: nsteps
      real(c_double) :: dt
      integer(c_int64_t) :: cg_iters, gmres_iters, steps, matvecs
      real(c_double) :: ms_total
      integer(c_int64_t) :: launches
   end type

   interface
      function nlk_last_error() bind(C, name="nlk_last_error") result(msg)
         import; type(c_ptr) :: msg
      end function
      ! ---- mesh / context
      integer(c_int) function nlk_partition(pid, nelg, nranks, gllnid) bind(C, name="nlk_partition")
         import; integer(c_int64_t), intent(in) :: pid(*); integer(c_int64_t), value :: nelg; integer(c_int32_t), value :: nranks
         integer(c_int32_t), intent(out) :: gllnid(*)
      end function
      integer(c_int) function nlk_mesh_create(desc, mesh) bind(C, name="nlk_mesh_create")
         import; type(nlk_mesh_desc), intent(in) :: desc; type(c_ptr), intent(out) :: mesh
      end function
      integer(c_int) function nlk_mesh_destroy(mesh) bind(C, name="nlk_mesh_destroy")
         import; type(c_ptr), value :: mesh
      end function
      integer(c_int) function nlk_params_default(p) bind(C, name="nlk_params_default")
         import; type(nlk_params), intent(out) :: p
      end function
      integer(c_int) function nlk_ctx_create(mesh, p, device, ctx) bind(C, name="nlk_ctx_create")
         import; type(c_ptr), value :: mesh; type(nlk_params), intent(in) :: p; integer(c_int32_t), value :: device; type(c_ptr), intent(out) :: ctx
      end function
      integer(c_int) function nlk_ctx_destroy(ctx) bind(C, name="nlk_ctx_destroy")
         import; type(c_ptr), value :: ctx
      end function
      integer(c_int) function nlk_ctx_set_tol(ctx, vtol, ptol) bind(C, name="nlk_ctx_set_tol")
         import; type(c_ptr), value :: ctx; real(c_double), value :: vtol, ptol
      end function
      integer(c_int) function nlk_ctx_set_dt(ctx, dt) bind(C, name="nlk_ctx_set_dt")
         import; type(c_ptr), value :: ctx; real(c_double), value :: dt
      end function
      integer(c_int) function nlk_comm_unique_id(id) bind(C, name="nlk_comm_unique_id")
         import; character(kind=c_char), intent(out) :: id(128)
      end function
      integer(c_int) function nlk_ctx_comm_init(ctx, id, rank, nranks) bind(C, name="nlk_ctx_comm_init")
         import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: id(128); integer(c_int32_t), value :: rank, nranks
      end function
      integer(c_int) function nlk_ctx_sync(ctx) bind(C, name="nlk_ctx_sync")
         import; type(c_ptr), value :: ctx
      end function
      ! ---- nek_dvector
      integer(c_int) function nlk_vec_create(ctx, v) bind(C, name="nlk_vec_create")
         import; type(c_ptr), value :: ctx; type(c_ptr), intent(out) :: v
      end function
      integer(c_int) function nlk_vec_destroy(v) bind(C, name="nlk_vec_destroy")
         import; type(c_ptr), value :: v
      end function
      integer(c_int) function nlk_vec_copy(dst, src) bind(C, name="nlk_vec_copy")
         import; type(c_ptr), value :: dst, src
      end function
      integer(c_int) function nlk_vec_zero(v) bind(C, name="nlk_vec_zero")
         import; type(c_ptr), value :: v
      end function
      integer(c_int) function nlk_vec_rand(v, ifnorm, seed) bind(C, name="nlk_vec_rand")
         import; type(c_ptr), value :: v; integer(c_int32_t), value :: ifnorm; integer(c_int64_t), value :: seed
      end function
      integer(c_int) function nlk_vec_scal(v, alpha) bind(C, name="nlk_vec_scal")
         import; type(c_ptr), value :: v; real(c_double), value :: alpha
      end function
      integer(c_int) function nlk_vec_axpby(alpha, x, beta, self) bind(C, name="nlk_vec_axpby")
         import; real(c_double), value :: alpha, beta; type(c_ptr), value :: x, self
      end function
      integer(c_int) function nlk_vec_dot(self, x, res) bind(C, name="nlk_vec_dot")
         import; type(c_ptr), value :: self, x; real(c_double), intent(out) :: res
      end function
      integer(c_int) function nlk_vec_norm(self, res) bind(C, name="nlk_vec_norm")
         import; type(c_ptr), value :: self; real(c_double), intent(out) :: res
      end function
      integer(c_int) function nlk_vec_size(v, n) bind(C, name="nlk_vec_size")
         import; type(c_ptr), value :: v; integer(c_int64_t), intent(out) :: n
      end function
      integer(c_int) function nlk_vec_save_rst(self, state, irst) bind(C, name="nlk_vec_save_rst")
         import; type(c_ptr), value :: self, state; integer(c_int32_t), value :: irst
      end function
      integer(c_int) function nlk_vec_get_rst(self, out, irst) bind(C, name="nlk_vec_get_rst")
         import; type(c_ptr), value :: self, out; integer(c_int32_t), value :: irst
      end function
      integer(c_int) function nlk_vec_nrst(v, nrst) bind(C, name="nlk_vec_nrst")
         import; type(c_ptr), value :: v; integer(c_int32_t), intent(out) :: nrst
      end function
      integer(c_int) function nlk_vec_clear_rst(v) bind(C, name="nlk_vec_clear_rst")
         import; type(c_ptr), value :: v
      end function
      integer(c_int) function nlk_vec_upload(v, vx, vy, vz, pr, theta) bind(C, name="nlk_vec_upload")
         import; type(c_ptr), value :: v, vx, vy, vz, pr, theta          ! c_loc(array) or c_null_ptr
      end function
      integer(c_int) function nlk_vec_download(v, vx, vy, vz, pr, theta) bind(C, name="nlk_vec_download")
         import; type(c_ptr), value :: v, vx, vy, vz, pr, theta
      end function
      integer(c_int) function nlk_nek2vec(ctx, v) bind(C, name="nlk_nek2vec")
         import; type(c_ptr), value :: ctx, v
      end function
      integer(c_int) function nlk_vec2nek(ctx, v) bind(C, name="nlk_vec2nek")
         import; type(c_ptr), value :: ctx, v
      end function
      ! ---- nek_zvector
      integer(c_int) function nlk_zvec_scal(re, im, ar, ai) bind(C, name="nlk_zvec_scal")
         import; type(c_ptr), value :: re, im; real(c_double), value :: ar, ai
      end function
      integer(c_int) function nlk_zvec_axpby(ar, ai, xre, xim, br, bi, sre, sim) bind(C, name="nlk_zvec_axpby")
         import; real(c_double), value :: ar, ai, br, bi; type(c_ptr), value :: xre, xim, sre, sim
      end function
      integer(c_int) function nlk_zvec_dot(sre, sim, xre, xim, ore, oim) bind(C, name="nlk_zvec_dot")
         import; type(c_ptr), value :: sre, sim, xre, xim; real(c_double), intent(out) :: ore, oim
      end function
      ! ---- block Gram-Schmidt (LightKrylov innerprod / linear_combination / double_gram_schmidt_step)
      integer(c_int) function nlk_basis_innerprod(X, k, y, h) bind(C, name="nlk_basis_innerprod")
         import; type(c_ptr), intent(in) :: X(*); integer(c_int32_t), value :: k; type(c_ptr), value :: y; real(c_double), intent(out) :: h(*)
      end function
      integer(c_int) function nlk_basis_axpy(y, X, k, c) bind(C, name="nlk_basis_axpy")
         import; type(c_ptr), value :: y; type(c_ptr), intent(in) :: X(*); integer(c_int32_t), value :: k; real(c_double), intent(in) :: c(*)
      end function
      integer(c_int) function nlk_basis_dgs(y, X, k, h, nrm) bind(C, name="nlk_basis_dgs")
         import; type(c_ptr), value :: y; type(c_ptr), intent(in) :: X(*); integer(c_int32_t), value :: k; real(c_double), intent(out) :: h(*), nrm
      end function
      ! ---- exptA_linop and the analysis drivers
      integer(c_int) function nlk_exptA_create(ctx, tau, baseflow, op) bind(C, name="nlk_exptA_create")
         import; type(c_ptr), value :: ctx, baseflow; real(c_double), value :: tau; type(c_ptr), intent(out) :: op
      end function
      integer(c_int) function nlk_exptA_destroy(op) bind(C, name="nlk_exptA_destroy")
         import; type(c_ptr), value :: op
      end function
      integer(c_int) function nlk_exptA_init(op) bind(C, name="nlk_exptA_init")
         import; type(c_ptr), value :: op
      end function
      integer(c_int) function nlk_exptA_set_tau(op, tau) bind(C, name="nlk_exptA_set_tau")
         import; type(c_ptr), value :: op; real(c_double), value :: tau
      end function
      integer(c_int) function nlk_exptA_set_baseflow(op, baseflow) bind(C, name="nlk_exptA_set_baseflow")
         import; type(c_ptr), value :: op, baseflow
      end function
      integer(c_int) function nlk_exptA_matvec(op, vin, vout) bind(C, name="nlk_exptA_matvec")
         import; type(c_ptr), value :: op, vin, vout
      end function
      integer(c_int) function nlk_exptA_rmatvec(op, vin, vout) bind(C, name="nlk_exptA_rmatvec")
         import; type(c_ptr), value :: op, vin, vout
      end function
      integer(c_int) function nlk_exptA_stats(op, st) bind(C, name="nlk_exptA_stats")
         import; type(c_ptr), value :: op; type(nlk_stats), intent(out) :: st
      end function
      integer(c_int) function nlk_nonlinear_map(ctx, tau, cfl_limit, vin, vout) bind(C, name="nlk_nonlinear_map")
         import; type(c_ptr), value :: ctx, vin, vout; real(c_double), value :: tau, cfl_limit
      end function
      integer(c_int) function nlk_newton_fixed_point(ctx, tau, X, tol, tol_mode, maxiter, gmres_kdim, rnorm_hist, niter, info) &
            bind(C, name="nlk_newton_fixed_point")
         import; type(c_ptr), value :: ctx, X; real(c_double), value :: tau, tol; integer(c_int32_t), value :: tol_mode, maxiter, gmres_kdim
         real(c_double), intent(out) :: rnorm_hist(*); integer(c_int32_t), intent(out) :: niter, info
      end function
      integer(c_int) function nlk_set_neklab_forcing(ctx, fx, fy, fz, ipert) bind(C, name="nlk_set_neklab_forcing")
         import; type(c_ptr), value :: ctx, fx, fy, fz; integer(c_int32_t), value :: ipert
      end function
      integer(c_int) function nlk_get_neklab_forcing(ctx, fx, fy, fz, ipert) bind(C, name="nlk_get_neklab_forcing")
         import; type(c_ptr), value :: ctx, fx, fy, fz; integer(c_int32_t), value :: ipert
      end function
      integer(c_int) function nlk_zero_neklab_forcing(ctx) bind(C, name="nlk_zero_neklab_forcing")
         import; type(c_ptr), value :: ctx
      end function
      integer(c_int) function nlk_zero_neklab_forcing_ipert(ctx, ipert) bind(C, name="nlk_zero_neklab_forcing_ipert")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: ipert
      end function
      integer(c_int) function nlk_eigs(op, nev, kdim, tol, transpose, x0, lam_re, lam_im, resid, eigvecs, niter, cb, user, info) &
            bind(C, name="nlk_eigs")
         import; type(c_ptr), value :: op, x0, eigvecs, user; integer(c_int32_t), value :: nev, kdim, transpose; real(c_double), value :: tol
         real(c_double), intent(out) :: lam_re(*), lam_im(*), resid(*); integer(c_int32_t), intent(out) :: niter, info; type(c_funptr), value :: cb
      end function
      integer(c_int) function nlk_svds(op, nsv, kdim, tol, x0, sigma, resid, U, V, niter, info) bind(C, name="nlk_svds")
         import; type(c_ptr), value :: op, x0, U, V; integer(c_int32_t), value :: nsv, kdim; real(c_double), value :: tol
         real(c_double), intent(out) :: sigma(*), resid(*); integer(c_int32_t), intent(out) :: niter, info
      end function
      integer(c_int) function nlk_gmres(op, minus_identity, b, x, kdim, atol, rtol, maxiter, transpose, info) bind(C, name="nlk_gmres")
         import; type(c_ptr), value :: op, b, x; integer(c_int32_t), value :: minus_identity, kdim, maxiter, transpose
         real(c_double), value :: atol, rtol; integer(c_int32_t), intent(out) :: info
      end function
      ! resolvent_linop (src/linops/resolvent.f90) and nek_upo_jacobian (src/systems/periodic_orbit.f90)
      integer(c_int) function nlk_resolvent_matvec(op, omega, f_re, f_im, out_re, out_im, adjoint, rtol, info) bind(C, name="nlk_resolvent_matvec")
         import; type(c_ptr), value :: op, f_re, f_im, out_re, out_im; real(c_double), value :: omega, rtol
         integer(c_int32_t), value :: adjoint; integer(c_int32_t), intent(out) :: info
      end function
      integer(c_int) function nlk_upo_jacobian(ctx, X, T_X, vin, T_in, vout, T_out, transpose) bind(C, name="nlk_upo_jacobian")
         import; type(c_ptr), value :: ctx, X, vin, vout; real(c_double), value :: T_X, T_in; real(c_double), intent(out) :: T_out
         integer(c_int32_t), value :: transpose
      end function
   end interface
end module nlk_c

!-------------------------------------------------------------------------------------------------------------------------
module neklab_b200
   use iso_c_binding
   use nlk_c
   use LightKrylov, only: dp, abstract_vector_rdp, abstract_linop_rdp
   use LightKrylov_Logger, only: stop_error, type_error
   implicit none
   private
   character(len=*), parameter :: this_module = 'neklab_b200'

   type(c_ptr), save, public :: nlk_ctx = c_null_ptr         ! one context per process, like Nek's COMMON blocks (not re-entrant)

   ! --> nek_dvector (src/vectors/neklab_vectors.f90:26-50): same TBP list, handle instead of static arrays
   type, extends(abstract_vector_rdp), public :: nek_dvector
      type(c_ptr) :: h = c_null_ptr
   contains
      private
      procedure, pass(self), public :: zero => nek_dzero
      procedure, pass(self), public :: rand => nek_drand
      procedure, pass(self), public :: scal => nek_dscal
      procedure, pass(self), public :: axpby => nek_daxpby
      procedure, pass(self), public :: dot => nek_ddot
      procedure, pass(self), public :: get_size => nek_dsize
      procedure, pass(self), public :: save_rst => dsave_rst
      procedure, pass(self), public :: get_rst => dget_rst
      procedure, pass(self), public :: has_rst_fields => dhas_rst_fields
      procedure, pass(self), public :: clear_rst_fields => dclear_rst_fields
      procedure, pass(self) :: nek_dcopy
      generic, public :: assignment(=) => nek_dcopy        ! value semantics: allocate(X(k), source=x) / X(i) = y deep-copy on the device
      final :: nek_dfinal
   end type nek_dvector

   ! --> exptA_linop (src/linops/neklab_linops.f90:35-44)
   type, extends(abstract_linop_rdp), public :: exptA_linop
      real(dp) :: tau = 1.0_dp
      type(nek_dvector) :: baseflow
      type(c_ptr) :: h = c_null_ptr
   contains
      private
      procedure, pass(self), public :: init => init_exptA
      procedure, pass(self), public :: matvec => exptA_matvec
      procedure, pass(self), public :: rmatvec => exptA_rmatvec
      final :: exptA_final
   end type exptA_linop

   public :: nek2vec, vec2nek, linear_stability_analysis_fixed_point, transient_growth_analysis_fixed_point
   public :: newton_fixed_point_iteration, set_neklab_forcing, get_neklab_forcing, zero_neklab_forcing, zero_neklab_forcing_ipert
   public :: resolvent_apply, upo_jacobian_apply

contains

   subroutine check(rc, procedure)
      integer(c_int), intent(in) :: rc
      character(len=*), intent(in) :: procedure
      character(kind=c_char), pointer :: cmsg(:)
      character(len=512) :: msg
      integer :: i
      if (rc == 0) return
      call c_f_pointer(nlk_last_error(), cmsg, [512])
      msg = ''
      do i = 1, 512
         if (cmsg(i) == c_null_char) exit
         msg(i:i) = cmsg(i)
      end do
      call stop_error(trim(msg), this_module, procedure)     ! the reference's nek_stop_error path (src/neklab_nek_setup.f90:406-417)
   end subroutine check

   subroutine ensure(self)
      class(nek_dvector), intent(inout) :: self
      if (.not. c_associated(self%h)) call check(nlk_vec_create(nlk_ctx, self%h), 'nek_dvector')
   end subroutine ensure

   ! ---- abstract_vector_rdp procedures (bodies of src/vectors/real_vectors.f90, one C call each)
   subroutine nek_dzero(self)                                     ! real_vectors.f90:37-50
      class(nek_dvector), intent(inout) :: self
      call ensure(self); call check(nlk_vec_zero(self%h), 'nek_dzero')
   end subroutine
   subroutine nek_drand(self, ifnorm)                             ! :52-123
      class(nek_dvector), intent(inout) :: self
      logical, optional, intent(in) :: ifnorm
      integer(c_int32_t) :: n
      n = 0; if (present(ifnorm)) n = merge(1, 0, ifnorm)
      call ensure(self); call check(nlk_vec_rand(self%h, n, 12345_c_int64_t), 'nek_drand')
   end subroutine
   subroutine nek_dscal(self, alpha)                              ! :125-160
      class(nek_dvector), intent(inout) :: self
      real(dp), intent(in) :: alpha
      call ensure(self); call check(nlk_vec_scal(self%h, alpha), 'nek_dscal')
   end subroutine
   subroutine nek_daxpby(alpha, vec, beta, self)                  ! :162-206 (incl. the rst arithmetic of :186-200)
      class(nek_dvector), intent(inout) :: self
      real(dp), intent(in) :: alpha, beta
      class(abstract_vector_rdp), intent(in) :: vec
      select type (vec)
      type is (nek_dvector)
         call ensure(self); call check(nlk_vec_axpby(alpha, vec%h, beta, self%h), 'nek_daxpby')
      class default
         call type_error('vec', 'nek_dvector', 'IN', this_module, 'nek_daxpby')
      end select
   end subroutine
   real(dp) function nek_ddot(self, vec) result(alpha)            ! :208-233 (bm1-weighted, pressure excluded)
      class(nek_dvector), intent(in) :: self
      class(abstract_vector_rdp), intent(in) :: vec
      alpha = 0.0_dp
      select type (vec)
      type is (nek_dvector)
         call check(nlk_vec_dot(self%h, vec%h, alpha), 'nek_ddot')
      class default
         call type_error('vec', 'nek_dvector', 'IN', this_module, 'nek_ddot')
      end select
   end function
   integer function nek_dsize(self) result(n)                     ! :235-247
      class(nek_dvector), intent(in) :: self
      integer(c_int64_t) :: n8
      call check(nlk_vec_size(self%h, n8), 'nek_dsize'); n = int(n8)
   end function
   subroutine dsave_rst(self, vec_rst, irst)                      ! :249-291
      class(nek_dvector), intent(inout) :: self
      class(abstract_vector_rdp), intent(in) :: vec_rst
      integer, intent(in) :: irst
      select type (vec_rst)
      type is (nek_dvector)
         call ensure(self); call check(nlk_vec_save_rst(self%h, vec_rst%h, int(irst, c_int32_t)), 'dsave_rst')
      class default
         call type_error('vec_rst', 'nek_dvector', 'IN', this_module, 'dsave_rst')
      end select
   end subroutine
   subroutine dget_rst(self, vec_rst, irst)                       ! :293-333
      class(nek_dvector), intent(in) :: self
      class(abstract_vector_rdp), intent(inout) :: vec_rst
      integer, intent(in) :: irst
      select type (vec_rst)
      type is (nek_dvector)
         call ensure(vec_rst); call check(nlk_vec_get_rst(self%h, vec_rst%h, int(irst, c_int32_t)), 'dget_rst')
      class default
         call type_error('vec_rst', 'nek_dvector', 'OUT', this_module, 'dget_rst')
      end select
   end subroutine
   logical function dhas_rst_fields(self) result(has_rst_fields)  ! :335-338
      class(nek_dvector), intent(in) :: self
      integer(c_int32_t) :: nrst
      call check(nlk_vec_nrst(self%h, nrst), 'dhas_rst_fields'); has_rst_fields = nrst > 0
   end function
   subroutine dclear_rst_fields(self)                             ! :340-346
      class(nek_dvector), intent(inout) :: self
      call check(nlk_vec_clear_rst(self%h), 'dclear_rst_fields')
   end subroutine
   subroutine nek_dcopy(self, from)
      class(nek_dvector), intent(inout) :: self
      type(nek_dvector), intent(in) :: from
      if (.not. c_associated(from%h)) return
      call ensure(self); call check(nlk_vec_copy(self%h, from%h), 'nek_dcopy')
   end subroutine
   subroutine nek_dfinal(self)
      type(nek_dvector), intent(inout) :: self
      integer(c_int) :: rc
      if (c_associated(self%h)) rc = nlk_vec_destroy(self%h)
      self%h = c_null_ptr
   end subroutine

   ! ---- nek2vec / vec2nek (src/neklab_utils.f90:84-134): Nek COMMON arrays <-> device vector
   subroutine nek2vec(vec, vx_, vy_, vz_, pr_, t_)
      type(nek_dvector), intent(out) :: vec
      real(dp), target, intent(in) :: vx_(*), vy_(*), vz_(*), pr_(*), t_(*)
      call ensure(vec)
      call check(nlk_vec_upload(vec%h, c_loc(vx_), c_loc(vy_), c_loc(vz_), c_loc(pr_), c_loc(t_)), 'nek2vec')
   end subroutine
   subroutine vec2nek(vx_, vy_, vz_, pr_, t_, vec)
      type(nek_dvector), intent(in) :: vec
      real(dp), target, intent(inout) :: vx_(*), vy_(*), vz_(*), pr_(*), t_(*)
      call check(nlk_vec_download(vec%h, c_loc(vx_), c_loc(vy_), c_loc(vz_), c_loc(pr_), c_loc(t_)), 'vec2nek')
   end subroutine

   ! ---- exptA_linop (src/linops/exponential_propagator.f90)
   subroutine init_exptA(self)                                    ! :4-13
      class(exptA_linop), intent(inout) :: self
      if (.not. c_associated(self%h)) call check(nlk_exptA_create(nlk_ctx, self%tau, self%baseflow%h, self%h), 'init_exptA')
      call check(nlk_exptA_set_tau(self%h, self%tau), 'init_exptA')
      call check(nlk_exptA_init(self%h), 'init_exptA')
   end subroutine
   subroutine exptA_matvec(self, vec_in, vec_out)                 ! :15-60 (time loop, get_rst, compute_rst all on the device)
      class(exptA_linop), intent(inout) :: self
      class(abstract_vector_rdp), intent(in) :: vec_in
      class(abstract_vector_rdp), intent(out) :: vec_out
      select type (vec_in)
      type is (nek_dvector)
         select type (vec_out)
         type is (nek_dvector)
            call ensure(vec_out)
            call check(nlk_exptA_set_tau(self%h, self%tau), 'exptA_matvec')        ! apply_exptA sets A%tau before every call (:224-266)
            call check(nlk_exptA_matvec(self%h, vec_in%h, vec_out%h), 'exptA_matvec')
         class default
            call type_error('vec_out', 'nek_dvector', 'OUT', this_module, 'exptA_matvec')
         end select
      class default
         call type_error('vec_in', 'nek_dvector', 'IN', this_module, 'exptA_matvec')
      end select
   end subroutine
   subroutine exptA_rmatvec(self, vec_in, vec_out)                ! :62-107
      class(exptA_linop), intent(inout) :: self
      class(abstract_vector_rdp), intent(in) :: vec_in
      class(abstract_vector_rdp), intent(out) :: vec_out
      select type (vec_in)
      type is (nek_dvector)
         select type (vec_out)
         type is (nek_dvector)
            call ensure(vec_out)
            call check(nlk_exptA_set_tau(self%h, self%tau), 'exptA_rmatvec')
            call check(nlk_exptA_rmatvec(self%h, vec_in%h, vec_out%h), 'exptA_rmatvec')
         class default
            call type_error('vec_out', 'nek_dvector', 'OUT', this_module, 'exptA_rmatvec')
         end select
      class default
         call type_error('vec_in', 'nek_dvector', 'IN', this_module, 'exptA_rmatvec')
      end select
   end subroutine
   subroutine exptA_final(self)
      type(exptA_linop), intent(inout) :: self
      integer(c_int) :: rc
      if (c_associated(self%h)) rc = nlk_exptA_destroy(self%h)
      self%h = c_null_ptr
   end subroutine

   ! ---- analysis drivers with device-resident Krylov bases (src/neklab_analysis.f90)
   subroutine linear_stability_analysis_fixed_point(exptA, kdim, nev, adjoint, eigvals, residuals, eigvecs)   ! :38-105
      type(exptA_linop), intent(inout) :: exptA
      integer, intent(in) :: kdim, nev
      logical, optional, intent(in) :: adjoint
      complex(dp), intent(out) :: eigvals(nev)
      real(dp), intent(out) :: residuals(nev)
      type(nek_dvector), optional, intent(inout) :: eigvecs(2*nev)       ! (re, im) pairs
      real(c_double) :: lre(nev), lim(nev)
      type(c_ptr), target :: hv(2*nev)
      integer(c_int32_t) :: niter, info, tr
      integer :: i
      tr = 0; if (present(adjoint)) tr = merge(1, 0, adjoint)
      if (present(eigvecs)) then
         do i = 1, 2*nev
            call ensure(eigvecs(i)); hv(i) = eigvecs(i)%h
         end do
         call check(nlk_eigs(exptA%h, int(nev, c_int32_t), int(kdim, c_int32_t), 0.0_c_double, tr, c_null_ptr, lre, lim, residuals, &
                             c_loc(hv), niter, c_null_funptr, c_null_ptr, info), 'linear_stability_analysis_fixed_point')
      else
         call check(nlk_eigs(exptA%h, int(nev, c_int32_t), int(kdim, c_int32_t), 0.0_c_double, tr, c_null_ptr, lre, lim, residuals, &
                             c_null_ptr, niter, c_null_funptr, c_null_ptr, info), 'linear_stability_analysis_fixed_point')
      end if
      eigvals = log(cmplx(lre, lim, kind=dp))/exptA%tau                  ! :84
   end subroutine
   subroutine transient_growth_analysis_fixed_point(exptA, nsv, kdim, sigma, residuals)                       ! :107-156
      type(exptA_linop), intent(inout) :: exptA
      integer, intent(in) :: nsv, kdim
      real(dp), intent(out) :: sigma(nsv), residuals(nsv)
      integer(c_int32_t) :: niter, info
      call check(nlk_svds(exptA%h, int(nsv, c_int32_t), int(kdim, c_int32_t), 0.0_c_double, c_null_ptr, sigma, residuals, c_null_ptr, c_null_ptr, &
                          niter, info), 'transient_growth_analysis_fixed_point')
   end subroutine
   subroutine newton_fixed_point_iteration(bf, tau, tol, tol_mode, info)                                      ! :158-212
      type(nek_dvector), intent(inout) :: bf
      real(dp), intent(in) :: tau, tol
      integer, optional, intent(in) :: tol_mode                          ! 1 = nek_constant_tol, 2 = nek_dynamic_tol
      integer, intent(out) :: info
      real(c_double) :: hist(64)
      integer(c_int32_t) :: niter, i4, mode
      mode = 1; if (present(tol_mode)) mode = int(tol_mode, c_int32_t)
      call check(nlk_newton_fixed_point(nlk_ctx, tau, bf%h, tol, mode, 40_c_int32_t, 30_c_int32_t, hist, niter, i4), 'newton_fixed_point_iteration')
      info = int(i4)
   end subroutine

   ! ---- resolvent_linop%matvec / %rmatvec (src/linops/resolvent.f90:17-75): the nek_zvector is passed as its (re, im) nek_dvector pair
   subroutine resolvent_apply(A, omega, f_re, f_im, out_re, out_im, adjoint, info)
      type(exptA_linop), intent(inout) :: A                              ! supplies the base flow (self%baseflow of resolvent_linop)
      real(dp), intent(in) :: omega
      type(nek_dvector), intent(in) :: f_re, f_im
      type(nek_dvector), intent(inout) :: out_re, out_im
      logical, intent(in) :: adjoint
      integer, intent(out) :: info
      integer(c_int32_t) :: i4, adj
      call ensure(out_re); call ensure(out_im)
      adj = 0; if (adjoint) adj = 1
      call check(nlk_resolvent_matvec(A%h, omega, f_re%h, f_im%h, out_re%h, out_im%h, adj, 0.0_c_double, i4), 'resolvent_matvec')
      info = int(i4)
   end subroutine

   ! ---- nek_upo_jacobian%matvec / %rmatvec (src/systems/periodic_orbit.f90:46-181): nek_ext_dvector = (nek_dvector, T)
   subroutine upo_jacobian_apply(X, T_X, vec_in, T_in, vec_out, T_out, transpose)
      type(nek_dvector), intent(in) :: X, vec_in
      real(dp), intent(in) :: T_X, T_in
      type(nek_dvector), intent(inout) :: vec_out
      real(dp), intent(out) :: T_out
      logical, intent(in) :: transpose
      integer(c_int32_t) :: tr
      real(c_double) :: t
      call ensure(vec_out)
      tr = 0; if (transpose) tr = 1
      call check(nlk_upo_jacobian(nlk_ctx, X%h, T_X, vec_in%h, T_in, vec_out%h, t, tr), 'jac_direct_map')
      T_out = t
   end subroutine

   ! ---- forcing registry (src/neklab_nek_forcing.f90)
   subroutine set_neklab_forcing(fx, fy, fz, ipert)
      real(dp), target, intent(in) :: fx(*), fy(*), fz(*)
      integer, intent(in) :: ipert
      call check(nlk_set_neklab_forcing(nlk_ctx, c_loc(fx), c_loc(fy), c_loc(fz), int(ipert, c_int32_t)), 'set_neklab_forcing')
   end subroutine
   subroutine get_neklab_forcing(fx, fy, fz, ipert)
      real(dp), target, intent(out) :: fx(*), fy(*), fz(*)
      integer, intent(in) :: ipert
      call check(nlk_get_neklab_forcing(nlk_ctx, c_loc(fx), c_loc(fy), c_loc(fz), int(ipert, c_int32_t)), 'get_neklab_forcing')
   end subroutine
   subroutine zero_neklab_forcing()
      call check(nlk_zero_neklab_forcing(nlk_ctx), 'zero_neklab_forcing')
   end subroutine
   subroutine zero_neklab_forcing_ipert(ipert)
      integer, intent(in) :: ipert
      call check(nlk_zero_neklab_forcing_ipert(nlk_ctx, int(ipert, c_int32_t)), 'zero_neklab_forcing_ipert')
   end subroutine
end module neklab_b200
