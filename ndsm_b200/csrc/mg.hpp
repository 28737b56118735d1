// mg.hpp -- device-resident multigrid hierarchy and V-cycle driver.
// Mirrors the reference's MG_HANDLE (ndsm_multigrid_core.f90:86-101), new_mg_handle (:165-270),
// v_cycle (:341-377) and solve_poisson_bvp (ndsm_poisson.f90:63-155); relax/residual are the
// CUDA operators of kernels.cu sitting behind the reference's MG_RELAX / MG_RESIDUAL seam.
#pragma once
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

struct SolveTrace {
  std::vector<double> du;      // du after each V-cycle
  std::vector<int> nexact;     // coarsest-solve iterations in each V-cycle
  int ierr = 0;
};

struct Level {
  Grid g;
  Bounds b;
  Weights w;
  std::vector<double> mesh[3];
  double* u = nullptr;    // colour-split, local plane 0
  double* rhs = nullptr;  // colour-split (level 0: may alias caller data or be null)
  // transfer tables between this level (fine) and the next (coarse); device pointers
  InterpTab it[3];
  RestrictTab rt[3];
};

class MG {
 public:
  // shape: (nx,ny,nz) with nz == 1 for ndim == 2.  ngrids < 0 => reference rule from min(shape).
  MG(int ndim, const int* shape, int ngrids, const double* const* mesh, cudaStream_t st);
  ~MG();
  MG(const MG&) = delete;
  MG& operator=(const MG&) = delete;

  void set_options(int ms, double ex_tol, const char* copt, bool du_max, int nmax_exact);
  int ngrids() const { return (int)lv_.size(); }
  int ndim() const { return ndim_; }
  const Level& level(int g) const { return lv_[g]; }
  Level& level(int g) { return lv_[g]; }
  double* r_scratch() { return r_; }
  size_t level_doubles(int g) const { return (size_t)2 * lv_[g].g.cs; }

  // operators on a level (enqueue only)
  void relax(int g);          // one full red+black sweep (+ pure-Neumann mean subtraction)
  void residual(int g);       // r_scratch <- rhs - L u
  void restrict_to(int g);    // rhs[g+1] <- R r_scratch (fine level g)
  void interp_add_from(int c);  // u[c-1] += P u[c]
  int solve_exact(int g);     // returns iterations (synchronises only in the fallback path)
  void v_cycle();             // from the finest grid; leaves coarsest info in d_info_

  // solve_poisson_bvp: u (colour-split, level-0 layout) in/out; rhs colour-split or nullptr (== 0)
  int solve(double* u, const double* rhs, double vc_tol, int nmax, double* du_last, SolveTrace* tr);
  void set_level0_rhs(const double* rhs) { rhs0_ = rhs; }
  // the same solve as a state machine (one outstanding solve per MG instance)
  void solve_begin(double* u, const double* rhs, double vc_tol, int nmax, SolveTrace* tr);
  void solve_enqueue();            // enqueue the next V-cycle (+ update_u) on the stream
  bool solve_poll();               // wait for it, read du; true when converged or nmax reached
  int solve_end(double* du_last);  // returns ierr
  bool solve_done() const { return ss_.done; }
  void enqueue_cycle(double* u);  // V-cycle + update_u + result copies (no sync)
  bool coarsest_in_smem(const double* rhs_coarsest) const;

  cudaStream_t stream() const { return st_; }
  bool du_max() const { return du_max_; }
  int nmax_exact() const { return nmax_exact_; }
  int last_nexact();  // synchronises

 private:
  int ndim_;
  std::vector<Level> lv_;
  cudaStream_t st_;
  int ms_ = -1;
  double ex_tol_ = -1;
  char copt_[8];
  bool du_max_ = true;
  int nmax_exact_ = 0;
  bool all_neumann_ = false;
  int first_colour_ = 0;
  const double* rhs0_ = nullptr;  // level-0 rhs for the current solve (nullptr == 0)
  double* arena_ = nullptr;       // one allocation for all level arrays
  int* tab_i_ = nullptr;          // packed transfer tables
  double* tab_d_ = nullptr;
  double* r_ = nullptr;           // residual scratch (level-0 sized)
  double* usav_ = nullptr;        // coarsest u_sav for the fallback solve_exact
  double* scratch_ = nullptr;     // reduction scratch
  double* d_out_ = nullptr;       // [2] device result of reductions
  int* d_info_ = nullptr;         // [2] coarsest-solve iterations / converged
  double* h_out_ = nullptr;       // pinned [4]: du_max, du_sum, + info
  struct SolveState {
    double* u = nullptr;
    double vc_tol = 0, du = 1.7976931348623157e308;
    int nmax = 0, it = 0;
    bool converged = false, done = false;
    cudaGraphExec_t gexec = nullptr;
    unsigned long long graph_launches = 0;
    SolveTrace* tr = nullptr;
    double* zero_rhs = nullptr;
  } ss_;
};

}  // namespace ndsm
