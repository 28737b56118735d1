// kernels.cuh -- launch wrappers for the sm_100a FP64 kernels of the NDSM V-cycle path.
// Every wrapper enqueues on `st` and returns immediately; all pointers are device pointers
// that address local plane 0 of a colour-split level (see common.cuh) unless noted "dense".
#pragma once
#include "common.cuh"

namespace ndsm {

extern unsigned long long g_launches;  // number of kernels this library has launched
extern unsigned long long g_peer_bytes, g_peer_msgs;  // peer-memory transport: bytes / messages stored into other GPUs

// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline numbers).
// Classes are recorded only for level-0 launches (tagged by the MG driver through prof_scope()).
enum ProfClass { PROF_RELAX0 = 0, PROF_RESID0, PROF_RESTRICT0, PROF_INTERP0, PROF_DIFF0, PROF_EXCH, PROF_TAIL, PROF_LEVEL1, PROF_NCLASS };
void prof_enable(bool on);
bool prof_enabled();
void prof_begin(int cls, cudaStream_t st);  // record start event (no-op when disabled)
void prof_end(int cls, cudaStream_t st);    // record stop event
void prof_collect();                        // after a stream sync: fold finished event pairs into the totals
void prof_get(int cls, unsigned long long* count, double* total_ms);
void prof_reset();

struct Weights {  // finite-difference weights of one level
  double wx, wy, wz;  // 1/h^2
  double w1;          // 3D: 1/(2*(wx+wy+wz))  (ndsm_optimized.f90:92-93); 2D: 1/(2wx+2wy) (ndsm_poisson.f90:483-489)
  double wc;          // 3D: 2*(wx+wy+wz)      (ndsm_optimized.f90:384)
};

struct InterpTab {  // prolongation table of one dimension (ndsm_interp.f90:120-146), indexed by fine index
  const int* lo;
  const double* wl;
  const double* wh;
};
struct RestrictTab {  // restriction table of one dimension (ndsm_interp.f90:218-252,277-282), by coarse index
  const int* first;
  const int* count;
  const double* c2;  // [n_coarse][NDSM_RMAX] : |dq_c - |q_f - q_c||
  double w2;         // dq_f / dq_c^2
};

// K1: one colour pass of the 3D red/black Gauss-Seidel sweep (ndsm_optimized.f90:103-167).
// rhs may be nullptr (rhs == 0 on the finest level of the vector-potential solves).
// uread (optional): array whose OTHER colour is read instead of u's (first pass of a ping-pong V-cycle).
void relax3d_half(double* u, const double* rhs, const Grid& g, const Bounds& b, int colour, const Weights& w,
                  int ext, cudaStream_t st, const double* uread = nullptr);
// keeps the stream busy for that long (one thread spinning on %globaltimer)
void stream_delay(double microseconds, cudaStream_t st);
// dst := src on the points no colour pass updates (Dirichlet faces incl. edges), both colours
void copy_fixed_points(const double* src, double* dst, const Grid& g, const Bounds& b, cudaStream_t st);
// Batched variants: up to three independent problems on the same grid in ONE launch (the components Ax, Ay, Az on a
// multi-GPU slab).  Every member brings its own arrays, Dirichlet pattern and pass colour; rhs is either given for
// all members or for none.
#define NDSM_BATCH_MAX 3
struct RelaxBatch {
  int n;
  double* u[NDSM_BATCH_MAX];
  const double* rhs[NDSM_BATCH_MAX];
  Bounds b[NDSM_BATCH_MAX];
  int colour[NDSM_BATCH_MAX];
  int klo[NDSM_BATCH_MAX], khi[NDSM_BATCH_MAX];  // filled by the wrapper
};
void relax3d_half_batch(const RelaxBatch& bt, const Grid& g, const Weights& w, int ext, cudaStream_t st);
struct ResidualBatch {
  int n;
  const double* u[NDSM_BATCH_MAX];
  const double* rhs[NDSM_BATCH_MAX];
  double* r[NDSM_BATCH_MAX];
  Bounds b[NDSM_BATCH_MAX];
};
void residual3d_batch(const ResidualBatch& bt, const Grid& g, const Weights& w, cudaStream_t st);
struct TransferBatch {  // restriction: src = r (fine), dst = rhs (coarse); prolongation: src = u (coarse), dst = u (fine)
  int n;
  const double* src[NDSM_BATCH_MAX];
  double* dst[NDSM_BATCH_MAX];
};
void restrict_direct_batch(const TransferBatch& bt, const Grid& gf, const Grid& gc, const RestrictTab& tx,
                           const RestrictTab& ty, const RestrictTab& tz, cudaStream_t st);
void interp_add_zt_batch(const TransferBatch& bt, const Grid& gc, const Grid& gf, const InterpTab& tx,
                         const InterpTab& ty, const InterpTab& tz, cudaStream_t st);
// K2: residual r = rhs - L u on non-Dirichlet points, 0 elsewhere (ndsm_optimized.f90:346-447).
void residual3d(const double* u, const double* rhs, double* r, const Grid& g, const Bounds& b, const Weights& w,
                cudaStream_t st);
// 2D (chi) variants: generic N-D relax / residual specialised to ndim = 2 (ndsm_poisson.f90:280-358,451-618).
void relax2d_half(double* u, const double* rhs, const Grid& g, const Bounds& b, int colour, const Weights& w,
                  cudaStream_t st);
void residual2d(const double* u, const double* rhs, double* r, const Grid& g, const Bounds& b, const Weights& w,
                cudaStream_t st);
// u -= sum(u)/N over all points (pure-Neumann gauge; ndsm_poisson.f90:529-547, ndsm_optimized.f90:173-189).
// scratch: >= reduce_scratch_doubles() doubles.
void subtract_mean(double* u, const Grid& g, double* scratch, cudaStream_t st);
// the same on a z-partitioned level: out2[0] = sum over this slab's owned planes; after the pairs of all ranks
// have been gathered (rank order), the mean is subtracted from the owned and the halo planes
void slab_sum(const double* u, const Grid& g, double* scratch, double* out2, cudaStream_t st);
void subtract_gathered_mean(double* u, const Grid& g, const double* pairs, int world, int halo, cudaStream_t st);
// K3: rhs_c = R r_f (ndsm_multigrid_core.f90:1010-1065 + ndsm_interp.f90:186-292), exact reference summation order.
void restrict_level(const double* rf, const Grid& gf, double* rhsc, const Grid& gc, const RestrictTab& tx,
                    const RestrictTab& ty, const RestrictTab& tz, cudaStream_t st);
// K3 tiled: same arithmetic as restrict_level (bit-identical) through a shared-memory window; for stencils of
// at most 5 points in x and y.  restrict_tiled_fits computes the window from the host tables and says
// whether the level pair qualifies (the caller uses restrict_level otherwise).
bool restrict_tiled_fits(const int* first_x, const int* count_x, int ncx, const int* first_y, const int* count_y,
                         int ncy, int* hwp, int* fyw);
void restrict_tiled(const double* rf, const Grid& gf, double* rhsc, const Grid& gc, const RestrictTab& tx,
                    const RestrictTab& ty, const RestrictTab& tz, int hwp, int fyw, cudaStream_t st);
// K4: u_f += P u_c on every fine point (ndsm_multigrid_core.f90:865-921,692-712 + ndsm_interp.f90:85-158).
void interp_add(const double* uc, const Grid& gc, double* uf, const Grid& gf, const InterpTab& tx,
                const InterpTab& ty, const InterpTab& tz, cudaStream_t st);
// K3 separable (3D, default): same 1-D weights applied one dimension at a time (HBM-bound; rounding differs from
// the reference's triple product at the 1e-16 level)
// K3 separable, direct: per-thread window reads from the colour-split arrays, no shared memory (same bits as
// restrict_sep); usable when restrict_direct_fits() holds (windows <= 5 points, <= 3 z windows open).
bool restrict_direct_fits(const int* const first[3], const int* const count[3], const int nc[3]);
void restrict_direct(const double* rf, const Grid& gf, double* rhsc, const Grid& gc, const RestrictTab& tx,
                     const RestrictTab& ty, const RestrictTab& tz, cudaStream_t st);
// K2+K3 fused (levels where restrict_direct_fits holds): rhsc[planes gc.k0..) = R (rhs - L u) without ever writing r;
// rz: scratch of residual_restrict_scratch(gf, gc.nzl) doubles; u valid one plane beyond the z windows it covers
size_t residual_restrict_scratch(const Grid& gf, int ncz_local);
void residual_restrict(const double* u, const double* rhs, const Grid& gf, const Bounds& b, const Weights& w,
                       double* rz, double* rhsc, const Grid& gc, const RestrictTab& tx, const RestrictTab& ty,
                       const RestrictTab& tz, cudaStream_t st);
bool restrict_sep_fits(const int* first_x, const int* count_x, int ncx, const int* first_y, const int* count_y,
                       int ncy);
void restrict_sep(const double* rf, const Grid& gf, double* rhsc, const Grid& gc, const RestrictTab& tx,
                  const RestrictTab& ty, const RestrictTab& tz, cudaStream_t st);
// K4 tiled (3D): same arithmetic through a shared-memory window of coarse planes
bool interp_tiled_fits(const int* lo_x, int nfx, int ncx, const int* lo_y, int nfy, int ncy);
// K4 z-lerped tile (3D, default): z-lerps shared through a small shared-memory tile, 16-byte u_f accesses.
bool interp_zt_fits(const int* lo_x, int nfx, int ncx, const int* lo_y, int nfy, int ncy);
void interp_add_zt(const double* uc, const Grid& gc, double* uf, const Grid& gf, const InterpTab& tx,
                   const InterpTab& ty, const InterpTab& tz, cudaStream_t st);
void interp_add_tiled(const double* uc, const Grid& gc, double* uf, const Grid& gf, const InterpTab& tx,
                      const InterpTab& ty, const InterpTab& tz, cudaStream_t st);
// K5: coarsest-level relaxation solve to ex_tol inside ONE thread block (ndsm_multigrid_core.f90:728-800).
// info[0] = iterations done, info[1] = converged flag.  Returns false if the level does not fit in shared memory.
bool solve_exact_smem(int ndim, double* u, const double* rhs, const Grid& g, const Bounds& b, int first_colour,
                      const Weights& w, bool all_neumann, bool du_max, double ex_tol, int nmax, int* info,
                      cudaStream_t st);
// Small-level sub-V-cycle in one thread block (levels of at most SMALL_MAX_POINTS points, dense in shared memory)
#define SMALL_MAX_POINTS 4096
#define SMALL_MAX_LEVELS 6
struct SmallLevel {
  int nx, ny, nz;
  Weights w;
  Bounds b;
  InterpTab it[3];    // towards the next coarser level
  RestrictTab rt[3];
  int off_u, off_rhs;  // shared-memory offsets (doubles)
};
struct SmallArgs {
  int nlev;
  SmallLevel lv[SMALL_MAX_LEVELS];
  int first_colour, all_neumann, du_max, nmax_exact, ms;
  double ex_tol;
  int off_r, off_sav, smem_doubles;
};
void vcycle_small_prepare();
// a batch of independent solves on the same grids: member m uses args_dev[slot[m]] (device memory)
struct SmallBatch {
  int n;
  const double* rhs_in[NDSM_BATCH_MAX];
  double* u_out[NDSM_BATCH_MAX];
  int* info[NDSM_BATCH_MAX];
  int slot[NDSM_BATCH_MAX];
};
void vcycle_small_batch(const SmallBatch& bt, const Grid& g0, const SmallArgs& a0, const SmallArgs* args_dev,
                        cudaStream_t st);
void vcycle_small(int ndim, const double* rhs_in, double* u_out, const Grid& g0, const SmallArgs& a, int* info,
                  cudaStream_t st);
// K6: out[0] = max|a-b|, out[1] = sum|a-b| over owned planes; then a := b  (update_u: a=caller's u, b=V-cycled u;
// ndsm_multigrid_core.f90:1077-1122).  With copy=false it is du_metrics (:808-853) and leaves a untouched.
void diff_reduce(double* a, const double* b, const Grid& g, bool copy, double* scratch, double* out, cudaStream_t st);
// npairs (max,sum) pairs and info[2] -> mapped pinned host memory (device-accessible pointer), npairs <= 24
void publish_results(const double* pairs, int npairs, const int* info, double* host_mapped, cudaStream_t st);
// batch of ninfo <= 3 solves: npairs pairs (<= 60), then info[m][0..1] for every member
void publish_results_batch(const double* pairs, int npairs, const int* const* info, int ninfo, double* host_mapped,
                           cudaStream_t st);
// nsweeps pure-Neumann 2D sweeps with the per-sweep mean subtraction folded into the passes (opt-in path of the
// chi solves, see kernels.cu); scratch: relax2d_fused_mean_scratch(g) doubles, zero-initialised once
size_t relax2d_fused_mean_scratch(const Grid& g);
void relax2d_fused_mean(double* u, const double* rhs, const Grid& g, const Bounds& b, const Weights& w, int nsweeps,
                        double* scratch, cudaStream_t st);
size_t reduce_scratch_doubles();
void solve_exact_prepare();  // one-time function attributes (call before stream capture)

// layout conversion: dense (nx,ny,nzl) x-fastest <-> colour-split; split = dense - shift
void split_from_dense(const double* dense, double* split, const Grid& g, double shift, cudaStream_t st);
void dense_from_split(const double* split, double* dense, const Grid& g, cudaStream_t st);

// K7 pieces: boundary-condition setup (ndsm_vector_potential.f90:283-306,387-399,647-682,977-1031,1070-1106)
void extract_face(const double* Bc_dense, int nx, int ny, int nz, int dim, int layer, double* face, cudaStream_t st);
void trapz_face(const double* face, int n1, int n2, double dq1, double dq2, double* scratch, double* out,
                cudaStream_t st);
void compute_At(const double* chi_split, const Grid& g2, double fac, int face_id, double* At1, double* At2,
                cudaStream_t st);
void write_face(double* A_split, const Grid& g, int dim, int layer, const double* face, cudaStream_t st);

// K8: flux-balance fields + curl (ndsm_vector_potential.f90:759-872,880-950).  Dense arrays are addressed as
// "pointer to a first global plane + component stride" so that a z-slab can be written in place.
// A_dense[planes ka..kb) = unsplit(A_split) (+ flux-balance potential of component comp when add_flux)
void unsplit_A(const double* As, const Grid& g, int comp, const double* x, const double* y, const double* z,
               const double* phi /*6*/, const double* Lq /*3*/, bool add_flux, int ka, int kb, double* A_dense,
               cudaStream_t st);
// B[planes k0..k1) = curl A; A holds planes from ka on (needs k-1,k+1 or the one-sided stencil planes)
// comp < 0: all three components of B; comp = c: only B_c (reads the two components of A it depends on)
void curl_dense(const double* A, int ka, i64 csA, int nx, int ny, int nz, double dqx, double dqy, double dqz, int k0,
                int k1, double* B, i64 csB, cudaStream_t st, int comp = -1);
void add_flux_dense(double* A, i64 csA, double* B, i64 csB, int nx, int ny, int k0, int k1, const double* x,
                    const double* y, const double* z, const double* phi, const double* Lq, cudaStream_t st);

}  // namespace ndsm
