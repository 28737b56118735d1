// vecpot.hpp -- device-resident vector-potential driver (compute_vector_potential,
// ndsm_vector_potential.f90:130-497) on top of the MG class.
#pragma once
#include <functional>
#include <vector>
#include "mg.hpp"

namespace ndsm {

enum {  // ndsm_vector_potential.f90:40-57
  IOPT_LEN = 16, IOPT_MS = 0, IOPT_NCYCLES = 1, IOPT_FACE1 = 2, IOPT_IERR = 3, IOPT_FLXCRL = 4, IOPT_DEBUG = 5,
  IOPT_DUMAX = 6, IOPT_NMAXEX = 7, IOPT_TRUE = 1, IOPT_FALSE = 0, ROPT_VTOL = 0, ROPT_CTOL = 1, ROPT_TIM = 2
};

struct Report {
  SolveTrace solves[9];  // chi faces 1..6, then Ax, Ay, Az
  double phi[6];
  // milliseconds
  double ms_total = 0, ms_in = 0, ms_bc = 0, ms_solve3d = 0, ms_post = 0, ms_out = 0, ms_device = 0;
  unsigned long long launches = 0;
  int ndist = 0;  // number of z-partitioned multigrid levels in the 3D solves
  int components_mode = 0;  // 0: one after the other, 1: three concurrent streams, 2: one batched launch sequence
  unsigned long long slab_points = 0;  // finest-level points this process smooths per component solve
};
extern Report g_report;

// NDSM_COMPONENT_GROUPS: "012" | "01,2" | "0,12" | "02,1" | "0,1,2" (every component once) -> sorted groups; anything
// else -> none.  Host only.
std::vector<std::vector<int>> parse_component_groups(const char* spec);

// Optional capture of BC-setup intermediates for parity tests (device -> host copies, dense faces)
struct BcCapture {
  double* chi[6] = {nullptr};
  double* At1[6] = {nullptr};
  double* At2[6] = {nullptr};
};

// Dense device arrays are described as "pointer to a first global z-plane + component stride".
struct DenseIn {   // initial guess A0: p == nullptr means zeros
  const double* p = nullptr;
  int kfirst = 0;       // global index of the first plane stored
  long long cstride = 0;
};
struct SlabOut {   // where one slab delivers A and B: planes [k0,k1), components cstride apart
  double* A = nullptr;
  double* B = nullptr;
  long long cstride = 0;
  int k0 = 0, k1 = 0;
};
// balanced z-range of rank `rank` (identical to the finest-level slab partition)
void output_range(int nz, int world, int rank, int* k0, int* k1);

// Optional hooks of the core (single-slab host entry).  `guess(c)` is called right before component c is solved and
// returns that component's initial guess (p == nullptr: zeros; kfirst/cstride describe the dense array p points
// into) -- the host scans / uploads A component by component, overlapped with the BC setup and the earlier solves.
// `component_ready(c)` is called as soon as component c of A is final in the output array (flux-balance field
// included) and `b_ready(c)` as soon as component c of B = curl A is (Bz once Ax and Ay are solved, Bx and By after
// Az), on stream st, so that their device-to-host copies overlap the remaining solves.
struct CoreHooks {
  std::function<DenseIn(int)> guess;
  std::function<void(int)> component_ready;
  std::function<void(int)> b_ready;
};

// drops the hierarchies, streams and graphs kept between calls (vecpot.cu, "Solver cache")
void solver_cache_clear();

// bn[f]: dense device faces (face f has shape (n1,n2) per ndsm_vector_potential.f90:225-246), all six on every
// rank.  comm == nullptr: single slab.  outs: one entry per slab held by this process.
// stop_after_bc: only run the BC setup (tests).  Returns iopt(IOPT_IERR).
int vector_solve_core(const int* nshape, const long long* iopt, const double* ropt, const double* x, const double* y,
                      const double* z, double* const* bn, const DenseIn& A0, Comm* comm, const std::vector<SlabOut>& outs,
                      cudaStream_t st, Report& rep, BcCapture* cap, bool stop_after_bc,
                      const CoreHooks* hooks = nullptr);

}  // namespace ndsm
