// mg.hpp -- device-resident multigrid hierarchy and V-cycle driver.
// Mirrors the reference's MG_HANDLE (ndsm_multigrid_core.f90:86-101), new_mg_handle (:165-270),
// v_cycle (:341-377) and solve_poisson_bvp (ndsm_poisson.f90:63-155); relax/residual are the
// CUDA operators of kernels.cu sitting behind the reference's MG_RELAX / MG_RESIDUAL seam.
//
// Multi-GPU: the finest levels are partitioned into z-slabs (one per rank) with halo planes; levels
// below a size threshold are replicated on every rank (all-gather of the restricted rhs once per
// V-cycle, no scatter on the way up).  A process holds one slab (NCCL mode, one process per GPU) or all
// slabs (virtual mode: every rank on one device, used to test the slab logic on a single GPU).
#pragma once
#include <array>
#include <map>
#include <memory>
#include <vector>
#include "kernels.cuh"

namespace ndsm {

// host-side 1-D transfer tables (computed with the reference's own formulas, in double)
void bracket_uniform(const double* q, int nq, double q0, int* lo, int* hi, int* ierr);  // ndsm_interp.f90:373-435
int ngrids_for(int nmin);  // FLOOR(LOG(nmin/2.0)/LOG(2.0)), ndsm_vector_potential.f90:341-342

// Host-only description of one level: shape, mesh, weights, HBM layout and the 1-D transfer
// tables to the next coarser level.  Built without touching CUDA (unit-testable on CPU).
struct HostLevel {
  int n[3];
  std::vector<double> mesh[3];
  Weights w;
  Grid g;
  // tables towards the next coarser level (empty on the coarsest level)
  std::vector<int> lo[3];
  std::vector<double> wl[3], wh[3];
  std::vector<int> first[3], count[3];
  std::vector<double> c2[3];  // [n_coarse][NDSM_RMAX]
  double w2[3];
};
// throws NdsmError(2) for shapes the reference cannot handle, NdsmError(4) if a stencil exceeds NDSM_RMAX
std::vector<HostLevel> build_hierarchy(int ndim, const int* shape, int ngrids, const double* const* mesh);

// z-slab partition of the hierarchy over `world` ranks (host-only).
struct SlabPlan {
  int world = 1;
  int ndist = 0;                     // levels [0, ndist) are partitioned; the rest are replicated
  int halo = 0;                      // halo planes on each side of a partitioned array
  std::vector<std::vector<int>> zs;  // zs[level][r] .. zs[level][r+1]: planes owned by rank r; the entry for
                                     // level ndist says who PRODUCES which planes of the first replicated level
};
#define NDSM_HALO 6
// a level is partitioned only while every rank keeps at least min_planes planes and the level has at least
// min_points points (smaller levels are replicated)
SlabPlan plan_slabs(const std::vector<HostLevel>& hl, int ndim, int world, int min_planes, long long min_points = 0);
// the thresholds MG uses for `world` ranks: 16 planes per rank; 8e6 points per level, 1e6 from 8 ranks on;
// NDSM_SLAB_MIN_PLANES / NDSM_SLAB_MIN_POINTS override
void slab_policy(int world, int* min_planes, long long* min_points);

// Communication between slabs.  Two-sided, matched in call order per (from,to) pair; all calls are
// enqueued on the stream and may be captured into a CUDA graph.
struct Comm {
  virtual ~Comm() {}
  virtual int world() const = 0;
  virtual int first_rank() const = 0;  // first global rank held by this process
  virtual int nlocal() const = 0;      // consecutive ranks held by this process
  virtual void begin(cudaStream_t st) = 0;
  virtual void send(int my_rank, int to_rank, const double* src, size_t n, cudaStream_t st) = 0;
  virtual void recv(int my_rank, int from_rank, double* dst, size_t n, cudaStream_t st) = 0;
  virtual void end(cudaStream_t st) = 0;
  // every rank contributes n doubles; result [n*world] on every process, ordered by rank
  virtual void gathern(int my_rank, const double* send, int n, double* recv_all, cudaStream_t st) = 0;
  void gather2(int my_rank, const double* send2, double* recv_all, cudaStream_t st) { gathern(my_rank, send2, 2, recv_all, st); }
  // in-place broadcast of buf[0..n) from root_rank into every rank's own copy of the same array; when all
  // local ranks share one copy (virtual mode) this is a no-op
  virtual void bcast(int root_rank, double* buf, size_t n, cudaStream_t st) = 0;
  // an independent communicator over the same ranks (collective: every rank calls it in the same order);
  // lets independent solves overlap their exchanges on different streams
  virtual std::unique_ptr<Comm> clone(cudaStream_t st) = 0;
  // sub-communicator of the ranks that pass the same colour (collective); ranks keep their relative order.
  // Not available for virtual ranks or the peer-memory transport (returns nullptr).
  virtual std::unique_ptr<Comm> split(int colour) = 0;
  // Memory that bcast() / gather2() may target.  The peer-memory transport writes straight into the other
  // ranks' copies of such a buffer, so it must sit at the same offset of a symmetric heap on every rank:
  // sym_alloc is COLLECTIVE (every rank calls it in the same order with the same size).  Other transports
  // hand out ordinary pool memory.
  virtual void* sym_alloc(size_t bytes);
  virtual void sym_free(void* p);
  // largest message (doubles) one send() of this communicator will carry to a neighbour (collective)
  virtual void reserve(size_t) {}
  // all ranks have enqueued everything that precedes this call (collective, stream ordered)
  virtual void barrier(cudaStream_t) {}
  // host bootstrap: every rank contributes `bytes` bytes, result ordered by rank (synchronises the stream)
  virtual void allgather_host(const void*, void*, size_t, cudaStream_t) { throw NdsmError(NDSM_ERR_INTERNAL); }
  // true once a device-side wait of this transport has timed out (results are then invalid)
  virtual bool failed() { return false; }
  virtual const char* transport() const = 0;
  // one-sided transport (peer-memory stores): independent solves may use clones of it concurrently
  virtual bool one_sided() const { return false; }
  // the communicator outlives a call (hierarchies built on it may be kept between calls)
  virtual bool persistent() const { return false; }
  // changes whenever something a captured graph may have baked in (the inbox) is reallocated
  virtual unsigned long long epoch() const { return 0; }
};
std::unique_ptr<Comm> make_virtual_comm(int world);  // all ranks in this process, on the current device
bool nccl_unique_id(void* out128);                    // ncclGetUniqueId (rank 0)
std::unique_ptr<Comm> make_nccl_comm(int rank, int world, const void* id128);  // one rank per process / GPU
// Peer-memory transport (peer.cu): one process per GPU, halo planes and replicated levels are written straight
// into the neighbours' HBM over NVLink (CUDA IPC mappings of a symmetric heap) and handed over with flags in
// peer memory; `boot` is only used to exchange the IPC handles.  Returns nullptr (after one stderr line) when
// peer mappings cannot be set up, in which case the caller stays on `boot`.
std::unique_ptr<Comm> make_peer_comm(Comm* boot, cudaStream_t st);
void peer_fabric_shutdown();  // unmap / free the symmetric heap (before the boot communicator goes away)

struct SolveTrace {
  std::vector<double> du;      // du after each V-cycle
  std::vector<int> nexact;     // coarsest-solve iterations in each V-cycle
  int ierr = 0;
};

struct Level {
  Grid g;               // local view: k0/nzl = owned planes, cs includes the halo planes
  Bounds b;
  Weights w;
  int H = 0;            // halo planes on each side (0 for replicated levels)
  bool dist = false;    // partitioned level
  double* u = nullptr;    // colour-split, local plane 0
  double* rhs = nullptr;  // colour-split (level 0: supplied per solve, may be null)
  // transfer tables between this level (fine) and the next (coarse); device pointers
  InterpTab it[3];
  RestrictTab rt[3];
  // tiled restriction towards the next level: shared-memory window, or fused == false (plain kernel)
  bool fused = false;
  int rr_hwp = 0, rr_fyw = 0;
  bool itiled = false;  // tiled prolongation from the next level
  bool rdirect = false;  // direct separable restriction to the next level (default when its windows are regular)
  bool icols = false;   // z-lerped-tile prolongation from the next level (default for large 3D levels)
  bool rsep = false;    // separable restriction towards the next level (default; not bit-identical)
};

struct Slab {
  int rank = 0;
  std::vector<Level> lv;
  double* r = nullptr;      // residual scratch base (sized for the largest level incl. halos)
  double* arena = nullptr;
  double* d_out = nullptr;  // [2] local max / sum of |du|
};

class MGBatch;

class MG {
  friend class MGBatch;
 public:
  // shape: (nx,ny,nz) with nz == 1 for ndim == 2.  ngrids < 0 => reference rule from min(shape).
  // comm == nullptr: one slab covering the whole grid.
  MG(int ndim, const int* shape, int ngrids, const double* const* mesh, cudaStream_t st, Comm* comm = nullptr);
  ~MG();
  MG(const MG&) = delete;
  MG& operator=(const MG&) = delete;

  void set_options(int ms, double ex_tol, const char* copt, bool du_max, int nmax_exact);
  int ngrids() const { return (int)mesh_.size(); }
  int ndim() const { return ndim_; }
  int nslabs() const { return (int)slabs_.size(); }
  const Level& level(int g, int s = 0) const { return slabs_[s].lv[g]; }
  Level& level(int g, int s = 0) { return slabs_[s].lv[g]; }
  const std::vector<double>& mesh(int g, int d) const { return mesh_[g][d]; }
  int slab_rank(int s) const { return slabs_[s].rank; }
  double* r_scratch(int g, int s = 0);  // local plane 0 of the residual scratch laid out like level g
  size_t level_doubles(int g, int s = 0) const { return (size_t)2 * slabs_[s].lv[g].g.cs; }
  double* level_base(double* plane0, int g, int s = 0) const {  // allocation start (first halo plane)
    const Level& L = slabs_[s].lv[g];
    return plane0 - (i64)L.H * L.g.ps;
  }
  const SlabPlan& plan() const { return plan_; }

  // operators on a level (enqueue only)
  void relax(int g);            // one full red+black sweep (+ pure-Neumann mean subtraction)
  void relax_sweeps(int g, int n);  // n sweeps; 2D pure-Neumann levels may fold the mean subtraction into the passes
  void residual(int g);         // r_scratch <- rhs - L u
  void restrict_to(int g);      // rhs[g+1] <- R r_scratch (fine level g); u[g+1] <- 0
  bool fused_restrict_ok(int g) const;
  void residual_restrict_to(int g);  // rhs[g+1] <- R (rhs - L u) without writing r (K2+K3 fused); u[g+1] <- 0
  void interp_add_from(int c);  // u[c-1] += P u[c]
  int solve_exact(int g);       // coarsest level (synchronises only in the fallback path)
  void v_cycle();               // from the finest grid; leaves coarsest info in d_info_
  // halo planes with the z-neighbours: which = 0 (u) or 2 (r scratch); arr overrides the array (per slab)
  void exchange(int g, int which, int colour_mask, int nplanes, const std::vector<double*>* arr = nullptr);

  // solve_poisson_bvp: u[s] (colour-split, level-0 layout of slab s incl. halos, pointer to local plane 0)
  // in/out; rhs[s] colour-split or nullptr (== 0)
  int solve(const std::vector<double*>& u, const std::vector<const double*>& rhs, double vc_tol, int nmax,
            double* du_last, SolveTrace* tr);
  int solve(double* u, const double* rhs, double vc_tol, int nmax, double* du_last, SolveTrace* tr);
  void set_level0_rhs(const double* rhs, int s = 0) { rhs0_[s] = rhs; }
  // the caller filled the halo planes of its level-0 rhs slabs (enables communication-avoiding smoothing there)
  void set_level0_rhs_halo_valid(bool ok) { rhs0_halo_ok_ = ok; }
  // the same solve as a state machine (one outstanding solve per MG instance)
  void solve_begin(const std::vector<double*>& u, const std::vector<const double*>& rhs, double vc_tol, int nmax,
                   SolveTrace* tr);
  void solve_begin(double* u, const double* rhs, double vc_tol, int nmax, SolveTrace* tr);
  void solve_enqueue();            // enqueue the next V-cycle (+ update_u) on the stream
  bool solve_poll();               // wait for it, read du; true when converged or nmax reached
  int solve_end(double* du_last);  // returns ierr
  bool solve_done() const { return ss_.done; }
  void enqueue_cycle();            // V-cycle + update_u + result copies (no sync)
  bool coarsest_in_smem(const double* rhs_coarsest) const;
  int small_from() const { return small_from_; }

  cudaStream_t stream() const { return st_; }
  bool du_max() const { return du_max_; }
  int nmax_exact() const { return nmax_exact_; }
  int last_nexact();  // synchronises

 private:
  int ndim_;
  std::vector<std::vector<std::vector<double>>> mesh_;  // [level][dim]
  std::vector<Slab> slabs_;
  SlabPlan plan_;
  Comm* comm_ = nullptr;
  cudaStream_t st_;
  int ms_ = -1;
  double ex_tol_ = -1;
  char copt_[8];
  bool du_max_ = true;
  int nmax_exact_ = 0;
  bool all_neumann_ = false;
  int first_colour_ = 0;
  std::vector<const double*> rhs0_;  // level-0 rhs per slab for the current solve (nullptr == 0)
  bool rhs0_halo_ok_ = false;        // halo planes of a caller-supplied level-0 rhs hold the neighbours' values
  // communication-avoiding smoothing on partitioned levels: valid_[g][c] = number of halo planes (each side)
  // that currently hold up-to-date values of colour c.  A colour pass may also update `e` halo planes if the
  // other colour is valid to depth e+1, so one 4-plane exchange feeds four passes (bit-identical values).
  std::vector<std::array<int, 2>> valid_;
  // static_ok_[g][c]: the Dirichlet points of colour c (never updated by a pass) are current in the halo planes
  std::vector<std::array<bool, 2>> static_ok_;
  std::vector<int> rneed_, ineed_;   // halo planes the restriction (of r, per fine level) / prolongation (of u, per
                                     // coarse level) read; the same on every rank
  void finish_restrict(int g);
  void need_halo(int g, int depth);  // make both colours of u[g] valid to at least `depth` planes
  void need_halo_colour(int g, int colour);
  int* tab_i_ = nullptr;             // packed transfer tables
  double* tab_d_ = nullptr;
  double* shared_ = nullptr;         // replicated levels, usav, reduction scratch, results
  double* usav_ = nullptr;           // coarsest u_sav for the fallback solve_exact
  double* scratch_ = nullptr;        // reduction scratch
  double* fm_scratch_ = nullptr;     // fused-mean 2D sweeps (NDSM_B200_FUSED_MEAN=1), nullptr when off
  double* d_all_ = nullptr;          // [2*world] gathered (max,sum) pairs
  double* d_allm_[2] = {nullptr, nullptr};  // gathered slab sums of the pure-Neumann mean, alternating
  int mean_parity_ = 0;
  int* d_info_ = nullptr;            // [2] coarsest-solve iterations / converged
  double* h_out_ = nullptr;          // pinned: [2*world] pairs + 2 ints
  double* h_out_dev_ = nullptr;      // device view of h_out_ (mapped), nullptr: copy nodes instead
  // levels >= small_from_ (each <= SMALL_MAX_POINTS points) run as one single-block kernel; 0 = disabled
  int small_from_ = 0;
  SmallArgs small_args_;
  void build_small_args();
  struct SolveState {
    std::vector<double*> u;
    double vc_tol = 0, du = 1.7976931348623157e308;
    int nmax = 0, it = 0;
    bool converged = false, done = false;
    bool use_graph = false;
    bool pingpong = false;  // V-cycles alternate between the hierarchy's level-0 array and the caller's (see enqueue_cycle)
    SolveTrace* tr = nullptr;
    std::vector<double*> zero_rhs;
  } ss_;
  // The V-cycle graph outlives the solve: a hierarchy that is kept across calls (vecpot.cu's solver cache) replays
  // it as long as every pointer and option baked into its kernel nodes is unchanged (the key).  Two slots: the
  // even and the odd cycle of a ping-pong solve address different arrays.
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    std::vector<unsigned long long> key;
    unsigned long long launches = 0, peer_bytes = 0, peer_msgs = 0;
  } gslot_[2];
  std::vector<unsigned long long> graph_key(int parity) const;
  void capture_cycle(int parity);
  void drop_graphs();
  double* u0_home_ = nullptr;       // the hierarchy's own level-0 array (single slab), where even cycles work
  const double* pp_read_ = nullptr;  // ping-pong: array the first colour pass of the cycle reads (previous iterate)
};

// Several independent solves on the same grids (the components Ax, Ay, Az) as ONE launch sequence: mg_batch.cu.
// members[0] is the leader: its stream and communicator carry everything; every member keeps its own hierarchy.
class MGBatch {
 public:
  explicit MGBatch(const std::vector<MG*>& members);
  ~MGBatch();
  MGBatch(const MGBatch&) = delete;
  MGBatch& operator=(const MGBatch&) = delete;
  // same grids / plan / options (after set_options), 3D, not pure Neumann, small-level kernel in use
  static bool compatible(const std::vector<MG*>& members);
  const std::vector<MG*>& members() const { return m_; }
  // u[i][s]: member i's iterate on slab s (level-0 layout, local plane 0), in/out; rhs == 0 for every member.
  // tr, du_last, ierr: per member (may be nullptr).
  void solve(const std::vector<std::vector<double*>>& u, double vc_tol, int nmax, SolveTrace* const* tr,
             double* du_last, int* ierr);
  // the same as a state machine (cf. MG::solve_begin ...): several groups interleaved by one host thread
  void solve_begin(const std::vector<std::vector<double*>>& u, double vc_tol, int nmax, SolveTrace* const* tr);
  void solve_enqueue();
  bool solve_poll();
  bool solve_done() const { return mask() == 0; }
  void solve_end(double* du_last, int* ierr);
  cudaStream_t stream() const { return st_; }
  // np halo planes of every member's level-0 array in one exchange
  void exchange_level0(const std::vector<std::vector<double*>>& u, int np);
  void drop_graphs();

 private:
  struct Item {
    MG* m;
    int mask;                         // colours
    const std::vector<double*>* arr;  // per slab, nullptr: the level's u (which == 0) or r scratch (which == 2)
  };
  struct Slot {
    cudaGraphExec_t exec = nullptr;
    std::vector<unsigned long long> key;
    unsigned long long launches = 0, peer_bytes = 0, peer_msgs = 0;
  };
  std::vector<MG*> m_;
  std::vector<char> on_;  // members still iterating
  std::vector<double> du_;
  std::vector<int> its_;
  std::vector<char> conv_;
  std::vector<SolveTrace*> tr_;
  std::vector<SmallArgs> small_host_;  // what d_small_ holds
  double vc_tol_ = 0;
  int nmax_ = 0;
  bool use_graph_ = false;
  MG* lead_ = nullptr;
  cudaStream_t st_ = nullptr;
  Comm* comm_ = nullptr;
  double* d_all_ = nullptr;      // gathered [rank][active member][max,sum]
  SmallArgs* d_small_ = nullptr;  // the members' small-level argument blocks
  double* h_out_ = nullptr;
  double* h_out_dev_ = nullptr;
  std::map<unsigned, Slot> graphs_;  // by mask of active members
  std::vector<MG*> active() const;
  unsigned mask() const;
  std::vector<unsigned long long> graph_key() const;
  void capture();
  void exchange(int g, int which, const std::vector<Item>& items, int np);
  void need_halo(int g, int depth);
  void relax(int g);
  void residual(int g);
  void restrict_to(int g);
  void interp_add_from(int c);
  void v_cycle();
  void enqueue_cycle();
};

}  // namespace ndsm
