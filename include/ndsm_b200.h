/*
 * ndsm_b200.h -- C ABI of ndsmf.so, the B200 (sm_100a) drop-in for the one hot path of
 * sag2021/ndsm: the multigrid V-cycle vector-potential solve.
 *
 * Section 1 is EXACTLY the reference's ISO_C_BINDING surface
 * (fortran/ndsm_python_wrapper.f90:56-234), same names, argument meaning and return values,
 * so the reference's unmodified ndsm.py (ndsm.py:141-210) drives this library.
 * Sections 2-4 are additions (new symbols only; nothing in section 1 changes).
 *
 * All pointers are HOST pointers unless a parameter is documented as a device pointer.
 * Arrays are Fortran-ordered like the reference: shape (nx,ny,nz[,3]), x fastest
 * (numpy [3,nz,ny,nx], C-contiguous).  The library never exits the process: failures are
 * reported through the return value / ioptc[3] and one "ERROR(sub):msg:id" line on stderr
 * (format of ndsm_root.f90:476-488).  There is no CPU fallback: without a CUDA device every
 * compute entry point returns NDSM_B200_ERR_CUDA.
 */
#ifndef NDSM_B200_H
#define NDSM_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes / ioptc[IOPT_IERR] values.  0 and 1 are the reference's (ndsm_poisson.f90:44-45). */
#define NDSM_B200_OK 0
#define NDSM_B200_ERR_NOT_CONVERGED 1 /* V-cycle limit hit (ndsm_poisson.f90:147-150) or mesh < 2 pts (ndsm_vector_potential.f90:213-216) */
#define NDSM_B200_ERR_SHAPE 2         /* min(nshape) < 4: the reference indexes an empty hierarchy here (undefined behaviour) */
#define NDSM_B200_ERR_CUDA 3          /* no device / CUDA runtime failure / out of memory */
#define NDSM_B200_ERR_STENCIL 4       /* restriction stencil wider than the compiled capacity */
#define NDSM_B200_ERR_ARG 5           /* inconsistent arguments (nsize != nx*ny*nz*3, nshape4[3] != 3, NULL pointer, NDSM_DEVICE out of range) */
#define NDSM_B200_ERR_INTERNAL 6      /* a consistency check inside the library failed (slab plan, message matching); never a CUDA error */

/* ------------------------------------------------------------------------------------------
 * 1. Reference surface (replaces fortran/ndsm_python_wrapper.f90)
 * ------------------------------------------------------------------------------------------ */

/* ndsm_python_wrapper.f90:56-154.  nsize = 3*nx*ny*nz; nshape4 = {nx,ny,nz,3};
 * ioptc[16]/ropt[16]: option vectors indexed by the getters below (in: ms, ncycles, flxcrl,
 * debug, dumax, nmaxex, vtol, ctol; out: ioptc[3] = ierr, ropt[2] = wall seconds).
 * x,y,z: uniform mesh vectors.  A: initial guess in, vector potential out.  B: only the
 * face-normal boundary values are read; overwritten with curl A.  Returns ioptc[3]. */
int ndsm_vector_solve(size_t nsize, const int* nshape4, int* ioptc, double* ropt, const double* x,
                      const double* y, const double* z, double* A, double* B);

int get_iopt_len(void);         /* :164-168  -> 16 */
int get_iopt_ierr(void);        /* :170-174  -> 16 (sic: the reference returns IOPT_LEN) */
int get_iopt_ms(void);          /* :176-180  -> 0 */
int get_iopt_ncycles(void);     /* :182-186  -> 1 */
int get_iopt_debug(void);       /* :188-192  -> 5 */
int get_iopt_dumax(void);       /* :194-198  -> 6 */
int get_iopt_iopt_nmaxex(void); /* :200-204  -> 7 (sic: doubled "iopt") */
int get_iopt_true(void);        /* :206-210  -> 1 */
int get_iopt_false(void);       /* :212-216  -> 0 */
int get_ropt_tim(void);         /* :218-222  -> 2 */
int get_ropt_vtol(void);        /* :224-228  -> 0 */
int get_ropt_ctol(void);        /* :230-234  -> 1 */

/* ------------------------------------------------------------------------------------------
 * 2. Device-resident entry and the scalar Poisson backend
 * ------------------------------------------------------------------------------------------ */

/* Same contract as ndsm_vector_solve, but dA and dB are DEVICE pointers to dense (nx,ny,nz,3)
 * arrays on the current device (dA: initial guess in / A out; dB: boundary values in / curl A
 * out).  x,y,z,ioptc,ropt stay host pointers.  ropt[2] = seconds including device sync. */
int ndsm_b200_vector_solve_device(const int* nshape4, int* ioptc, double* ropt, const double* x,
                                  const double* y, const double* z, double* dA, double* dB);

/* Multi-GPU (one process per GPU, z-slab domain decomposition).  Levels with at least NDSM_SLAB_MIN_PLANES (default
 * 16) planes per rank and NDSM_SLAB_MIN_POINTS (default 8e6) points are partitioned with NDSM_HALO_PLANES (default 6)
 * halo planes per side: a colour pass also updates the halo planes it has valid inputs for, so ONE exchange of the
 * halo feeds up to six colour passes (communication-avoiding smoothing; same arithmetic on the same inputs, hence
 * the same bits as the single-GPU solve).  Coarser levels are replicated on every rank.
 * Transport: peer-memory stores over NVLink / NVSwitch (CUDA IPC mappings of a symmetric heap, flag hand-over, two
 * small kernels per exchange captured in the V-cycle's CUDA graph; csrc/peer.cu).  NCCL only carries the IPC handles;
 * NDSM_P2P=0 (set before dist_init), or peer mappings being unavailable, moves the data path to NCCL send/recv groups.
 * Every rank holds a z-slab of all three components and solves them concurrently on three streams; the six 2D chi
 * solves of the BC setup are distributed (face f on rank f mod world) and their At faces broadcast.
 * Bootstrap: rank 0 calls dist_unique_id, the 128 bytes are distributed by the caller (MPI, torch.distributed,
 * a file ...), every rank calls dist_init on its own device (NDSM_DEVICE or the current device).
 * vector_solve_rank: every rank passes all six boundary faces as dense arrays (face f of shape (n1,n2) with the
 * lower-numbered axis fastest: x-faces (ny,nz), y-faces (nx,nz), z-faces (nx,ny); ndsm_vector_potential.f90:225-246)
 * and receives planes [k0,k1) = slab_range(nz, world, rank) of A and B, laid out (nx,ny,k1-k0,3).  Faces and
 * outputs may be host or device pointers (flags).  The initial guess is zero (what ndsm.py always passes).
 * The V-cycle trace of a rank covers all three 3D solves, and the chi solves of the faces it owns. */
int ndsm_b200_dist_unique_id(void* out128);
int ndsm_b200_dist_init(int rank, int world, const void* id128);
int ndsm_b200_dist_finalize(void);
int ndsm_b200_dist_world(void);
const char* ndsm_b200_dist_transport(void); /* description of the active data path */
int ndsm_b200_dist_rank(void);
int ndsm_b200_slab_range(int nz, int world, int rank, int* k0, int* k1);
int ndsm_b200_vector_solve_rank(const int* nshape4, int* ioptc, double* ropt, const double* x, const double* y,
                                const double* z, const double* const* faces6, int faces_on_device, double* A_slab,
                                double* B_slab, int out_on_device);

/* solve_poisson_bvp for a caller-defined scalar problem (ndsm_poisson.f90:63-155 with
 * new_mg_handle, ndsm_multigrid_core.f90:165): ndim = 2 or 3; copt = 2*ndim chars 'N'/'D' in the
 * reference order [lo_1..lo_ndim, hi_1..hi_ndim]; u: initial guess in / solution out (Dirichlet
 * faces hold the boundary data); rhs may be NULL (zero).  Returns 0/1 like ierr; *ncycles =
 * V-cycles performed. */
int ndsm_b200_poisson_solve(int ndim, const int* nshape, const char* copt, int ms, int ncycles_max,
                            int nmaxex, int du_max, double vc_tol, double ex_tol, const double* x,
                            const double* y, const double* z, double* u, const double* rhs,
                            double* du_last, int* ncycles);
/* The same 3D solve on z-slabs, one process per GPU (after ndsm_b200_dist_init): rank r passes planes [k0,k1) =
 * ndsm_b200_slab_range(nz, world, r) of u (in: initial guess incl. Dirichlet data, out: solution) and of rhs
 * (NULL == 0) as dense (k1-k0, ny, nx) DEVICE arrays; nshape3 is the global shape.  Every rank returns the same
 * ierr / du_last / ncycles.  The finest level must be large enough to be partitioned (NDSM_B200_ERR_ARG otherwise);
 * pure-Neumann problems (all six copt == 'N') are not supported on partitioned levels yet.  Bit-identical to
 * ndsm_b200_poisson_solve with the max metric (BASELINE config 5: weak scaling 1025 x 1025 x (128 G + 1)). */
int ndsm_b200_poisson_solve_rank(const int* nshape3, const char* copt, int ms, int ncycles_max, int nmaxex, int du_max,
                                 double vc_tol, double ex_tol, const double* x, const double* y, const double* z,
                                 double* d_u_slab, const double* d_rhs_slab, double* du_last, int* ncycles);

/* ------------------------------------------------------------------------------------------
 * 3. MG_HANDLE operator seam (ndsm_multigrid_core.f90:86-136,165,278,341): the GPU smoother,
 *    residual and transfer operators exposed one call at a time for parity tests.  Levels are
 *    0-based (0 = finest).  "which": 0 = u, 1 = rhs, 2 = residual scratch r.
 * ------------------------------------------------------------------------------------------ */
typedef struct ndsm_b200_mg ndsm_b200_mg;

ndsm_b200_mg* ndsm_b200_new_mg_handle(int ndim, const int* nshape, int ngrids /* <0: reference rule */,
                                      const double* x, const double* y, const double* z, int du_max,
                                      int nmax_exact);                       /* new_mg_handle :165 */
void ndsm_b200_delete_mg_handle(ndsm_b200_mg* h);                            /* delete_mg_handle :278 */
int ndsm_b200_mg_set_options(ndsm_b200_mg* h, int ms, double ex_tol, const char* copt); /* bvp%ms, %ex_tol, %copt */
int ndsm_b200_mg_ngrids(const ndsm_b200_mg* h);
int ndsm_b200_mg_level_shape(const ndsm_b200_mg* h, int level, int* shape3);
int ndsm_b200_mg_level_mesh(const ndsm_b200_mg* h, int level, int dim, double* out);
int ndsm_b200_mg_put(ndsm_b200_mg* h, int which, int level, const double* dense);
int ndsm_b200_mg_get(ndsm_b200_mg* h, int which, int level, double* dense);
int ndsm_b200_mg_relax(ndsm_b200_mg* h, int level, int nsweeps);             /* MG_RELAX :115-121 */
int ndsm_b200_mg_residual(ndsm_b200_mg* h, int level);                        /* MG_RESIDUAL :126-132 -> r */
int ndsm_b200_mg_restrict(ndsm_b200_mg* h, int level);                        /* mg_restrict :1010: r(level) -> rhs(level+1), u(level+1)=0 */
int ndsm_b200_mg_residual_restrict(ndsm_b200_mg* h, int level, int* fused);    /* fine_to_coarse's transfer :539-558 in one step: rhs(level+1) = R (rhs - L u), u(level+1) = 0; *fused = 1 when r was never written (K2+K3 fused kernels) */
int ndsm_b200_mg_interp_add(ndsm_b200_mg* h, int level);                      /* mg_interp+add_correction :865,692: u(level-1) += P u(level) */
int ndsm_b200_mg_solve_exact(ndsm_b200_mg* h, int level, int* iters);         /* solve_exact :728 */
int ndsm_b200_mg_v_cycle(ndsm_b200_mg* h);                                    /* v_cycle :341 */
int ndsm_b200_mg_solve(ndsm_b200_mg* h, double vc_tol, int nmax, double* u_dense, const double* rhs_dense,
                       double* du_last, int* ncycles);                        /* solve_poisson_bvp */
int ndsm_b200_mg_update_u(ndsm_b200_mg* h, const double* u_old_dense, double* u_new_dense, double* du_max,
                          double* du_mean);                                   /* update_u :1077 */

/* Host-only planning (no CUDA calls): hierarchy shapes (new_mg_handle, ndsm_multigrid_core.f90:215-262),
 * finite-difference weights, HBM layout and the 1-D transfer tables derived from ndsm_interp.f90:120-146
 * (prolongation: lo, wl, wh per fine index) and :218-252,277-282 (restriction: first, count, c2 per coarse
 * index, w2).  Levels are 0-based; tables of level L connect L (fine) to L+1 (coarse). */
typedef struct ndsm_b200_plan ndsm_b200_plan;
#define NDSM_B200_RMAX 8
ndsm_b200_plan* ndsm_b200_plan_create(int ndim, const int* nshape, int ngrids, const double* x, const double* y,
                                      const double* z);
void ndsm_b200_plan_destroy(ndsm_b200_plan* p);
int ndsm_b200_plan_ngrids(const ndsm_b200_plan* p);
int ndsm_b200_plan_level(const ndsm_b200_plan* p, int level, int* shape3, long long* layout4 /* hp,mcnt,ps,cs */,
                         double* weights5 /* wx,wy,wz,w1,wc */);
int ndsm_b200_plan_mesh(const ndsm_b200_plan* p, int level, int dim, double* out);
int ndsm_b200_plan_interp(const ndsm_b200_plan* p, int level, int dim, int* lo, double* wl, double* wh);
int ndsm_b200_plan_restrict(const ndsm_b200_plan* p, int level, int dim, int* first, int* count, double* c2,
                            double* w2);
int ndsm_b200_ngrids_for(int nmin); /* FLOOR(LOG(nmin/2.0)/LOG(2.0)), ndsm_vector_potential.f90:341-342 */
/* z-slab partition of the hierarchy over `world` ranks: *ndist = number of partitioned levels; zs holds
 * (ndist+1) rows of world+1 plane boundaries (row ndist = producers of the first replicated level).
 * zs must have room for ngrids*(world+1) ints.  min_planes >= 0: every level whose slabs keep that many planes is
 * partitioned; min_planes < 0: the thresholds a solve on `world` ranks uses (16 planes per rank; 8e6 points per level,
 * 1e6 from 8 ranks on; NDSM_SLAB_MIN_PLANES / NDSM_SLAB_MIN_POINTS override). */
int ndsm_b200_plan_slab_partition(const ndsm_b200_plan* p, int world, int min_planes, int* ndist, int* zs);
/* Host-only replay of the offset allocator of the multi-GPU symmetric heap (csrc/sym_alloc.hpp): ops[i] > 0 allocates
 * that many bytes and stores the offset in out[i] (-1: no room); ops[i] < 0 frees the block of operation -ops[i]-1.
 * Every rank runs the same sequence and must arrive at the same offsets (CPU tests, world_size 2 over gloo). */
int ndsm_b200_plan_sym_heap(long long segment_bytes, const long long* ops, int nops, long long* out);

/* Stage hooks of the driver (host dense arrays) */
int ndsm_b200_bc_setup(const int* nshape4, const int* ioptc, const double* ropt, const double* x, const double* y,
                       const double* z, const double* B, double* phi6, double** chi6, double** At1_6,
                       double** At2_6); /* ndsm_vector_potential.f90:247-399 */
int ndsm_b200_flux_curl(const int* nshape4, int flxcrl, const double* x, const double* y, const double* z,
                        const double* phi6, double* A, double* B); /* :453-477: add_flux_balance_fields + curl */

/* ------------------------------------------------------------------------------------------
 * 4. Introspection
 * ------------------------------------------------------------------------------------------ */
int ndsm_b200_device_count(void);               /* 0 when no usable CUDA device */
unsigned long long ndsm_b200_launch_count(void); /* kernels launched by this library so far */
/* peer-memory transport: bytes this process has stored into other GPUs' memory over NVLink so far, and the number of
 * messages (one k_push launch with remote segments each); 0 on the NCCL data path */
unsigned long long ndsm_b200_peer_bytes_sent(void);
unsigned long long ndsm_b200_peer_messages_sent(void);
/* trace of the last ndsm_vector_solve / _device / poisson_solve / mg_solve call:
 * solves 0..5 = chi faces 1..6, 6..8 = Ax, Ay, Az (poisson/mg_solve: solve 0) */
int ndsm_b200_trace_nsolves(void);
int ndsm_b200_trace_ncycles(int solve);
double ndsm_b200_trace_du(int solve, int cycle);
int ndsm_b200_trace_nexact(int solve, int cycle);
/* milliseconds of the last vector solve: [0] total wall, [1] input staging + H2D, [2] BC setup,
 * [3] 3D solves, [4] flux+curl, [5] D2H, [6] device-resident total (CUDA events), [7] kernel launches */
int ndsm_b200_last_timing(double* out8);
/* finest-level points this process smoothed per component solve in the last call (its z-slab) */
unsigned long long ndsm_b200_last_slab_points(void);
/* number of z-partitioned multigrid levels of the 3D solves in the last call (0 = everything replicated) */
int ndsm_b200_last_partitioned_levels(void);
/* how the three component solves of the last call were scheduled: 0 = one after the other (one GPU), 1 = three
 * concurrent streams, 2 = one batched launch sequence (NDSM_BATCH_COMPONENTS) */
int ndsm_b200_last_components_mode(void);
/* how a NDSM_COMPONENT_GROUPS value is read (host only, no GPU needed): returns the number of groups (0: the value is
 * not a partition of {0,1,2} and is ignored) and, per component, the index of its group */
int ndsm_b200_parse_component_groups(const char* spec, int* group_of3);
/* CUDA-event timing of the finest-level 3D kernels (off by default).  cls: 0 = k_relax3d colour pass,
 * 1 = k_residual3d, 2 = restriction, 3 = prolongation, 4 = update_u reduction (2 launches), 5 = one halo exchange,
 * 6 = all work on levels >= 2 of one V-cycle, 7 = all work on level 1 of one V-cycle (two brackets per cycle). */
void ndsm_b200_profile_enable(int on);
int ndsm_b200_profile_get(int cls, unsigned long long* count, double* total_ms);
/* Device and pinned staging buffers are cached between calls (a 513^3 solve needs ~15 GB and cudaMalloc/cudaFree
 * of that costs more than the solve).  release_workspace returns the cache to CUDA; workspace_bytes reports it. */
void ndsm_b200_release_workspace(void);
unsigned long long ndsm_b200_workspace_bytes(void);
const char* ndsm_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NDSM_B200_H */
