! ndsm_b200_bindings.f90 -- ISO_C_BINDING interface module for the C ABI of ndsmf.so (include/ndsm_b200.h).
!
! This is how a Fortran host (the reference's drivers ndsm_vector_potential.f90 / ndsm_poisson.f90, or any
! caller written like them) reaches the CUDA layer: every interface below binds one extern "C" entry point,
! argument for argument.  Kinds mirror ndsm_python_wrapper.f90:56-88: C ints for the shape / option
! vectors, C doubles for everything real, nsize by VALUE as INTEGER(C_SIZE_T).
!
! COMPILE-UNTESTED: the build image of this repository has no Fortran compiler (no gfortran / flang / nvfortran,
! neither here nor on the GPU box), so this file has only been checked by eye against the header.  With a
! compiler at hand:   gfortran -c include/ndsm_b200_bindings.f90   (see include/Makefile.fortran).
!
! Section numbers follow include/ndsm_b200.h.
MODULE NDSM_B200_BINDINGS
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PUBLIC

  ! return codes (ndsm_b200.h); 0 and 1 are the reference's (ndsm_poisson.f90:44-45)
  INTEGER(C_INT), PARAMETER :: NDSM_B200_OK = 0, NDSM_B200_ERR_NOT_CONVERGED = 1, NDSM_B200_ERR_SHAPE = 2, &
                               NDSM_B200_ERR_CUDA = 3, NDSM_B200_ERR_STENCIL = 4, NDSM_B200_ERR_ARG = 5, &
                               NDSM_B200_ERR_INTERNAL = 6

  INTERFACE
    ! ------------------------------------------------------------------------------------------------------
    ! 1. Reference surface (ndsm_python_wrapper.f90:56-234)
    ! ------------------------------------------------------------------------------------------------------
    ! drop-in for NDSM_PYTHON_WRAPPER's ndsm_vector_solve (:56): nshape4 = (nx,ny,nz,3), A/B of shape (nx,ny,nz,3)
    FUNCTION ndsm_vector_solve(nsize,nshape4,ioptc,ropt,x,y,z,A,B) BIND(C,NAME="ndsm_vector_solve") RESULT(ierr)
      IMPORT :: C_INT, C_SIZE_T, C_DOUBLE
      INTEGER(C_SIZE_T), VALUE      :: nsize
      INTEGER(C_INT), INTENT(IN)    :: nshape4(4)
      INTEGER(C_INT), INTENT(INOUT) :: ioptc(0:15)
      REAL(C_DOUBLE), INTENT(INOUT) :: ropt(0:15)
      REAL(C_DOUBLE), INTENT(IN)    :: x(*), y(*), z(*)
      REAL(C_DOUBLE), INTENT(INOUT) :: A(*), B(*)
      INTEGER(C_INT)                :: ierr
    END FUNCTION
    ! option-vector index getters (:164-234); names as exported, including the reference's "iopt_iopt" typo
    FUNCTION get_iopt_len() BIND(C,NAME="get_iopt_len") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_ierr() BIND(C,NAME="get_iopt_ierr") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_ms() BIND(C,NAME="get_iopt_ms") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_ncycles() BIND(C,NAME="get_iopt_ncycles") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_debug() BIND(C,NAME="get_iopt_debug") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_dumax() BIND(C,NAME="get_iopt_dumax") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_iopt_nmaxex() BIND(C,NAME="get_iopt_iopt_nmaxex") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_true() BIND(C,NAME="get_iopt_true") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_iopt_false() BIND(C,NAME="get_iopt_false") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_ropt_tim() BIND(C,NAME="get_ropt_tim") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_ropt_vtol() BIND(C,NAME="get_ropt_vtol") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION get_ropt_ctol() BIND(C,NAME="get_ropt_ctol") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION

    ! ------------------------------------------------------------------------------------------------------
    ! 2. Device-resident entry, multi-GPU entries, scalar Poisson backend
    ! ------------------------------------------------------------------------------------------------------
    ! device-resident variant of ndsm_vector_solve: dA, dB are CUDA device addresses (CUDA Fortran DEVICE arrays via
    ! C_DEVLOC, or OpenACC host_data use_device) of dense (nx,ny,nz,3) arrays
    FUNCTION ndsm_b200_vector_solve_device(nshape4,ioptc,ropt,x,y,z,dA,dB) &
             BIND(C,NAME="ndsm_b200_vector_solve_device") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_PTR
      INTEGER(C_INT), INTENT(IN)    :: nshape4(4)
      INTEGER(C_INT), INTENT(INOUT) :: ioptc(0:15)
      REAL(C_DOUBLE), INTENT(INOUT) :: ropt(0:15)
      REAL(C_DOUBLE), INTENT(IN)    :: x(*), y(*), z(*)
      TYPE(C_PTR), VALUE            :: dA, dB
      INTEGER(C_INT)                :: ierr
    END FUNCTION
    ! bootstrap of the multi-GPU decomposition: one process (MPI rank) per GPU; id128 = 128 bytes from rank 0
    FUNCTION ndsm_b200_dist_unique_id(out128) BIND(C,NAME="ndsm_b200_dist_unique_id") RESULT(ierr)
      IMPORT :: C_INT, C_CHAR
      CHARACTER(KIND=C_CHAR), INTENT(OUT) :: out128(128)
      INTEGER(C_INT)                      :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_dist_init(rank,world,id128) BIND(C,NAME="ndsm_b200_dist_init") RESULT(ierr)
      IMPORT :: C_INT, C_CHAR
      INTEGER(C_INT), VALUE              :: rank, world
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: id128(128)
      INTEGER(C_INT)                     :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_dist_finalize() BIND(C,NAME="ndsm_b200_dist_finalize") RESULT(ierr)
      IMPORT :: C_INT
      INTEGER(C_INT) :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_dist_world() BIND(C,NAME="ndsm_b200_dist_world") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    FUNCTION ndsm_b200_dist_rank() BIND(C,NAME="ndsm_b200_dist_rank") RESULT(v)
      IMPORT :: C_INT
      INTEGER(C_INT) :: v
    END FUNCTION
    ! planes [k0,k1) (0-based, k1 exclusive) of rank `rank`
    FUNCTION ndsm_b200_slab_range(nz,world,rank,k0,k1) BIND(C,NAME="ndsm_b200_slab_range") RESULT(ierr)
      IMPORT :: C_INT
      INTEGER(C_INT), VALUE       :: nz, world, rank
      INTEGER(C_INT), INTENT(OUT) :: k0, k1
      INTEGER(C_INT)              :: ierr
    END FUNCTION
    ! this rank's z-slab of A and B; faces6 = six C pointers (host or device addresses, see the flags) to the dense
    ! boundary-normal faces in the order x0,x1,y0,y1,z0,z1 (ndsm_vector_potential.f90:225-246)
    FUNCTION ndsm_b200_vector_solve_rank(nshape4,ioptc,ropt,x,y,z,faces6,faces_on_device,A_slab,B_slab,out_on_device) &
             BIND(C,NAME="ndsm_b200_vector_solve_rank") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_PTR
      INTEGER(C_INT), INTENT(IN)    :: nshape4(4)
      INTEGER(C_INT), INTENT(INOUT) :: ioptc(0:15)
      REAL(C_DOUBLE), INTENT(INOUT) :: ropt(0:15)
      REAL(C_DOUBLE), INTENT(IN)    :: x(*), y(*), z(*)
      TYPE(C_PTR), INTENT(IN)       :: faces6(6)
      INTEGER(C_INT), VALUE         :: faces_on_device, out_on_device
      TYPE(C_PTR), VALUE            :: A_slab, B_slab
      INTEGER(C_INT)                :: ierr
    END FUNCTION
    ! solve_poisson_bvp (ndsm_poisson.f90:63) for a 2D/3D scalar problem; copt like bvp%copt(1:2*ndim),
    ! NUL-terminated ("NDDNDD"//C_NULL_CHAR); rhs may be C_NULL_PTR-equivalent only through the C_PTR variant below
    FUNCTION ndsm_b200_poisson_solve(ndim,nshape,copt,ms,ncycles_max,nmaxex,du_max,vc_tol,ex_tol, &
                                     x,y,z,u,rhs,du_last,ncycles) BIND(C,NAME="ndsm_b200_poisson_solve") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_CHAR
      INTEGER(C_INT), VALUE         :: ndim, ms, ncycles_max, nmaxex, du_max
      INTEGER(C_INT), INTENT(IN)    :: nshape(*)
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: copt(*)
      REAL(C_DOUBLE), VALUE         :: vc_tol, ex_tol
      REAL(C_DOUBLE), INTENT(IN)    :: x(*), y(*), z(*), rhs(*)
      REAL(C_DOUBLE), INTENT(INOUT) :: u(*)
      REAL(C_DOUBLE), INTENT(OUT)   :: du_last
      INTEGER(C_INT), INTENT(OUT)   :: ncycles
      INTEGER(C_INT)                :: ierr
    END FUNCTION
    ! the same 3D solve on z-slabs, one process per GPU: u_slab / rhs_slab are device addresses of planes [k0,k1)
    FUNCTION ndsm_b200_poisson_solve_rank(nshape3,copt,ms,ncycles_max,nmaxex,du_max,vc_tol,ex_tol,x,y,z, &
                                          u_slab,rhs_slab,du_last,ncycles) &
             BIND(C,NAME="ndsm_b200_poisson_solve_rank") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_CHAR, C_PTR
      INTEGER(C_INT), INTENT(IN)    :: nshape3(3)
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: copt(*)
      INTEGER(C_INT), VALUE         :: ms, ncycles_max, nmaxex, du_max
      REAL(C_DOUBLE), VALUE         :: vc_tol, ex_tol
      REAL(C_DOUBLE), INTENT(IN)    :: x(*), y(*), z(*)
      TYPE(C_PTR), VALUE            :: u_slab, rhs_slab
      REAL(C_DOUBLE), INTENT(OUT)   :: du_last
      INTEGER(C_INT), INTENT(OUT)   :: ncycles
      INTEGER(C_INT)                :: ierr
    END FUNCTION

    ! ------------------------------------------------------------------------------------------------------
    ! 3. MG_HANDLE operator seam (ndsm_multigrid_core.f90:86-136,165,278,341): the handle is an opaque C pointer;
    !    levels are 0-based (0 = finest); which: 0 = u, 1 = rhs, 2 = residual scratch
    ! ------------------------------------------------------------------------------------------------------
    FUNCTION ndsm_b200_new_mg_handle(ndim,nshape,ngrids,x,y,z,du_max,nmax_exact) &
             BIND(C,NAME="ndsm_b200_new_mg_handle") RESULT(h)      ! new_mg_handle (:165)
      IMPORT :: C_INT, C_DOUBLE, C_PTR
      INTEGER(C_INT), VALUE      :: ndim, ngrids, du_max, nmax_exact
      INTEGER(C_INT), INTENT(IN) :: nshape(*)
      REAL(C_DOUBLE), INTENT(IN) :: x(*), y(*), z(*)
      TYPE(C_PTR)                :: h
    END FUNCTION
    SUBROUTINE ndsm_b200_delete_mg_handle(h) BIND(C,NAME="ndsm_b200_delete_mg_handle")   ! delete_mg_handle (:278)
      IMPORT :: C_PTR
      TYPE(C_PTR), VALUE :: h
    END SUBROUTINE
    FUNCTION ndsm_b200_mg_set_options(h,ms,ex_tol,copt) BIND(C,NAME="ndsm_b200_mg_set_options") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_CHAR, C_PTR
      TYPE(C_PTR), VALUE                 :: h
      INTEGER(C_INT), VALUE              :: ms
      REAL(C_DOUBLE), VALUE              :: ex_tol
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: copt(*)
      INTEGER(C_INT)                     :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_ngrids(h) BIND(C,NAME="ndsm_b200_mg_ngrids") RESULT(n)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT)     :: n
    END FUNCTION
    FUNCTION ndsm_b200_mg_level_shape(h,level,shape3) BIND(C,NAME="ndsm_b200_mg_level_shape") RESULT(ierr)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: h
      INTEGER(C_INT), VALUE       :: level
      INTEGER(C_INT), INTENT(OUT) :: shape3(3)
      INTEGER(C_INT)              :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_put(h,which,level,dense) BIND(C,NAME="ndsm_b200_mg_put") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_PTR
      TYPE(C_PTR), VALUE         :: h
      INTEGER(C_INT), VALUE      :: which, level
      REAL(C_DOUBLE), INTENT(IN) :: dense(*)
      INTEGER(C_INT)             :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_get(h,which,level,dense) BIND(C,NAME="ndsm_b200_mg_get") RESULT(ierr)
      IMPORT :: C_INT, C_DOUBLE, C_PTR
      TYPE(C_PTR), VALUE          :: h
      INTEGER(C_INT), VALUE       :: which, level
      REAL(C_DOUBLE), INTENT(OUT) :: dense(*)
      INTEGER(C_INT)              :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_relax(h,level,nsweeps) BIND(C,NAME="ndsm_b200_mg_relax") RESULT(ierr)       ! MG_RELAX (:115)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE    :: h
      INTEGER(C_INT), VALUE :: level, nsweeps
      INTEGER(C_INT)        :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_residual(h,level) BIND(C,NAME="ndsm_b200_mg_residual") RESULT(ierr)          ! MG_RESIDUAL (:126)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE    :: h
      INTEGER(C_INT), VALUE :: level
      INTEGER(C_INT)        :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_restrict(h,level) BIND(C,NAME="ndsm_b200_mg_restrict") RESULT(ierr)          ! mg_restrict (:1010)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE    :: h
      INTEGER(C_INT), VALUE :: level
      INTEGER(C_INT)        :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_residual_restrict(h,level,fused) BIND(C,NAME="ndsm_b200_mg_residual_restrict") RESULT(ierr)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: h
      INTEGER(C_INT), VALUE       :: level
      INTEGER(C_INT), INTENT(OUT) :: fused
      INTEGER(C_INT)              :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_interp_add(h,level) BIND(C,NAME="ndsm_b200_mg_interp_add") RESULT(ierr)      ! mg_interp + add_correction (:865,:692)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE    :: h
      INTEGER(C_INT), VALUE :: level
      INTEGER(C_INT)        :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_solve_exact(h,level,iters) BIND(C,NAME="ndsm_b200_mg_solve_exact") RESULT(ierr)  ! solve_exact (:728)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: h
      INTEGER(C_INT), VALUE       :: level
      INTEGER(C_INT), INTENT(OUT) :: iters
      INTEGER(C_INT)              :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_v_cycle(h) BIND(C,NAME="ndsm_b200_mg_v_cycle") RESULT(ierr)                  ! v_cycle (:341)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE :: h
      INTEGER(C_INT)     :: ierr
    END FUNCTION
    FUNCTION ndsm_b200_mg_solve(h,vc_tol,nmax,u_dense,rhs_dense,du_last,ncycles) &
             BIND(C,NAME="ndsm_b200_mg_solve") RESULT(ierr)                                            ! solve_poisson_bvp
      IMPORT :: C_INT, C_DOUBLE, C_PTR
      TYPE(C_PTR), VALUE            :: h
      REAL(C_DOUBLE), VALUE         :: vc_tol
      INTEGER(C_INT), VALUE         :: nmax
      REAL(C_DOUBLE), INTENT(INOUT) :: u_dense(*)
      REAL(C_DOUBLE), INTENT(IN)    :: rhs_dense(*)
      REAL(C_DOUBLE), INTENT(OUT)   :: du_last
      INTEGER(C_INT), INTENT(OUT)   :: ncycles
      INTEGER(C_INT)                :: ierr
    END FUNCTION

    ! ------------------------------------------------------------------------------------------------------
    ! 4. Introspection / workspace
    ! ------------------------------------------------------------------------------------------------------
    FUNCTION ndsm_b200_device_count() BIND(C,NAME="ndsm_b200_device_count") RESULT(n)
      IMPORT :: C_INT
      INTEGER(C_INT) :: n
    END FUNCTION
    SUBROUTINE ndsm_b200_release_workspace() BIND(C,NAME="ndsm_b200_release_workspace")
    END SUBROUTINE
    FUNCTION ndsm_b200_trace_ncycles(s) BIND(C,NAME="ndsm_b200_trace_ncycles") RESULT(n)   ! s = 0..5 chi faces, 6..8 Ax,Ay,Az
      IMPORT :: C_INT
      INTEGER(C_INT), VALUE :: s
      INTEGER(C_INT)        :: n
    END FUNCTION
    FUNCTION ndsm_b200_trace_du(s,c) BIND(C,NAME="ndsm_b200_trace_du") RESULT(du)
      IMPORT :: C_INT, C_DOUBLE
      INTEGER(C_INT), VALUE :: s, c
      REAL(C_DOUBLE)        :: du
    END FUNCTION
  END INTERFACE

CONTAINS

  ! The reference's compute_vector_potential call (ndsm_python_wrapper.f90:129) with the reference's own argument
  ! kinds: INTEGER(8) shape / option vectors as in ndsm_root.f90:61 (IT), converted to the C ints of the ABI exactly
  ! like the wrapper converts in the other direction (:113-127,151).
  SUBROUTINE compute_vector_potential_b200(nshape, iopt, ropt, x, y, z, Apot, B)
    INTEGER(C_INT64_T), INTENT(IN)    :: nshape(4)
    INTEGER(C_INT64_T), INTENT(INOUT) :: iopt(0:15)
    REAL(C_DOUBLE), INTENT(INOUT)     :: ropt(0:15)
    REAL(C_DOUBLE), INTENT(IN)        :: x(*), y(*), z(*)
    REAL(C_DOUBLE), INTENT(INOUT)     :: Apot(*), B(*)
    INTEGER(C_INT) :: nshape4(4), ioptc(0:15), ierr
    INTEGER(C_SIZE_T) :: nsize
    nshape4 = INT(nshape, C_INT)
    ioptc = INT(iopt, C_INT)
    nsize = INT(nshape(1)*nshape(2)*nshape(3)*nshape(4), C_SIZE_T)
    ierr = ndsm_vector_solve(nsize, nshape4, ioptc, ropt, x, y, z, Apot, B)
    iopt = INT(ioptc, C_INT64_T)
  END SUBROUTINE

END MODULE NDSM_B200_BINDINGS
