// mg.cu -- multigrid hierarchy, transfer tables and the V-cycle loop (host orchestration).
// See mg.hpp for the reference functions this mirrors.
#include "mg.hpp"
#include "pool.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ndsm {

extern bool g_debug;
void debug_msg(const char* sub, const char* msg);

// find_bracket_points_uniform (ndsm_interp.f90:373-435), 0-based results
void bracket_uniform(const double* q, int nq, double q0, int* lo, int* hi, int* ierr) {
  if (q0 <= q[0]) { *lo = 0; *hi = 1; *ierr = -1; return; }
  if (q0 >= q[nq - 1]) { *lo = nq - 2; *hi = nq - 1; *ierr = +1; return; }
  const double dq = q[1] - q[0];
  long long l = (long long)std::floor((q0 - q[0]) / dq) + 1;  // 1-based, :419
  if (l >= nq) { *lo = nq - 2; *hi = nq - 1; }                // :423-425
  else { *lo = (int)l - 1; *hi = (int)l; }
  *ierr = 0;
}

int ngrids_for(int nmin) { return (int)std::floor(std::log((double)nmin / 2.0) / std::log(2.0)); }

static inline i64 round_up(i64 a, i64 b) { return (a + b - 1) / b * b; }

static Grid make_grid(int nx, int ny, int nz) {
  Grid g;
  g.nx = nx; g.ny = ny; g.nz = nz;
  g.k0 = 0; g.nzl = nz;
  g.mcnt = (nx + 1) / 2;
  g.hp = (int)round_up(g.mcnt, 8);
  g.ps = round_up((i64)g.hp * ny, 32);
  g.cs = g.ps * nz;
  return g;
}

std::vector<HostLevel> build_hierarchy(int ndim, const int* shape, int ngrids, const double* const* mesh) {
  if (ndim != 2 && ndim != 3) throw NdsmError(2);
  int nmin = shape[0];
  for (int d = 1; d < ndim; ++d) nmin = shape[d] < nmin ? shape[d] : nmin;
  for (int d = 0; d < ndim; ++d)
    if (shape[d] < 2) throw NdsmError(2);
  if (ngrids < 0) ngrids = ngrids_for(nmin);
  if (ngrids < 1 || ngrids > 40) throw NdsmError(2);  // reference indexes an empty hierarchy here (UB)
  std::vector<HostLevel> lv(ngrids);

  // --- shapes (ndsm_multigrid_core.f90:215-217), meshes (:231-262), weights
  int sh[3] = {shape[0], shape[1], ndim == 3 ? shape[2] : 1};
  for (int g = 0; g < ngrids; ++g) {
    HostLevel& L = lv[g];
    if (g > 0)
      for (int d = 0; d < ndim; ++d) {
        int v = (int)std::floor(sh[d] * 0.5);
        sh[d] = v > 1 ? v : 1;
      }
    for (int d = 0; d < 3; ++d) L.n[d] = sh[d];
    L.g = make_grid(sh[0], sh[1], sh[2]);
    for (int d = 0; d < ndim; ++d) {
      if (sh[d] < 2) throw NdsmError(2);
      L.mesh[d].resize(sh[d]);
      if (g == 0) {
        for (int j = 0; j < sh[d]; ++j) L.mesh[d][j] = mesh[d][j];
      } else {
        double qmin = mesh[d][0], qmax = mesh[d][0];
        for (int j = 1; j < shape[d]; ++j) {
          qmin = mesh[d][j] < qmin ? mesh[d][j] : qmin;
          qmax = mesh[d][j] > qmax ? mesh[d][j] : qmax;
        }
        const double Lq = qmax - qmin;
        for (int j = 0; j < sh[d]; ++j) L.mesh[d][j] = ((double)j * Lq) / (double)(sh[d] - 1) + qmin;  // :258
      }
    }
    const double hx = L.mesh[0][1] - L.mesh[0][0], hy = L.mesh[1][1] - L.mesh[1][0];
    L.w.wx = 1.0 / (hx * hx);
    L.w.wy = 1.0 / (hy * hy);
    if (ndim == 3) {
      const double hz = L.mesh[2][1] - L.mesh[2][0];
      L.w.wz = 1.0 / (hz * hz);
      L.w.wc = 2 * ((L.w.wx + L.w.wy) + L.w.wz);  // ndsm_optimized.f90:384 / :92
      L.w.w1 = 1.0 / L.w.wc;
    } else {
      L.w.wz = 0.0;
      double w0 = 0.0;  // ndsm_poisson.f90:483-489
      w0 = w0 + 2.0 * L.w.wx;
      w0 = w0 + 2.0 * L.w.wy;
      L.w.wc = w0;
      L.w.w1 = 1.0 / w0;
    }
  }

  // --- 1-D transfer tables with the reference formulas
  for (int g = 0; g + 1 < ngrids; ++g) {
    HostLevel& F = lv[g];
    HostLevel& C = lv[g + 1];
    for (int d = 0; d < 3; ++d) {
      const int nf = F.n[d], nc = C.n[d];
      F.lo[d].assign(nf, 0); F.wl[d].assign(nf, 0.0); F.wh[d].assign(nf, 1.0);
      F.first[d].assign(nc, 0); F.count[d].assign(nc, 1);
      F.c2[d].assign((size_t)nc * NDSM_RMAX, 0.0);
      F.w2[d] = 1.0;
      if (d >= ndim) {  // degenerate z of a 2D face: identity
        F.c2[d][0] = 1.0;
        continue;
      }
      const double* qf = F.mesh[d].data();
      const double* qc = C.mesh[d].data();
      for (int i = 0; i < nf; ++i) {  // ndsm_interp.f90:120-146
        int l, h, e;
        bracket_uniform(qc, nc, qf[i], &l, &h, &e);
        const double ql = qc[l], qh = qc[h], dq = qh - ql;
        F.lo[d][i] = l;
        F.wl[d][i] = +(qf[i] - ql) / dq;
        F.wh[d][i] = -(qf[i] - qh) / dq;
      }
      const double dqc = qc[1] - qc[0], dqf = qf[1] - qf[0];
      F.w2[d] = dqf / (dqc * dqc);  // :228
      for (int c = 0; c < nc; ++c) {  // :234-252
        int l, h, e, a, b;
        bracket_uniform(qf, nf, qc[c] - dqc, &l, &h, &e);
        a = (e < 0) ? l : h;
        bracket_uniform(qf, nf, qc[c] + dqc, &l, &h, &e);
        b = (e > 0) ? h : l;
        F.first[d][c] = a;
        F.count[d][c] = b - a + 1;
        if (b - a + 1 > NDSM_RMAX || b - a + 1 < 1) throw NdsmError(4);
        for (int j = a; j <= b; ++j) {  // :277-280
          const double c1 = std::fabs(qf[j] - qc[c]);
          F.c2[d][(size_t)c * NDSM_RMAX + (j - a)] = std::fabs(dqc - c1);
        }
      }
    }
  }
  return lv;
}

MG::MG(int ndim, const int* shape, int ngrids, const double* const* mesh, cudaStream_t st) : ndim_(ndim), st_(st) {
  std::vector<HostLevel> hl = build_hierarchy(ndim, shape, ngrids, mesh);
  ngrids = (int)hl.size();
  lv_.resize(ngrids);
  std::memset(copt_, 'N', sizeof copt_);
  for (int g = 0; g < ngrids; ++g) {
    lv_[g].g = hl[g].g;
    lv_[g].w = hl[g].w;
    for (int d = 0; d < 3; ++d) lv_[g].mesh[d] = hl[g].mesh[d];
  }

  // --- one arena for the level arrays
  i64 total = 0;
  auto take = [&](i64 n) { i64 o = total; total += round_up(n, 32); return o; };
  std::vector<i64> off_u(ngrids), off_rhs(ngrids);
  for (int g = 0; g < ngrids; ++g) {
    off_u[g] = take(2 * lv_[g].g.cs);
    off_rhs[g] = (g > 0) ? take(2 * lv_[g].g.cs) : -1;
  }
  const i64 off_r = take(2 * lv_[0].g.cs);
  const i64 off_sav = take(2 * lv_[ngrids - 1].g.cs);
  const i64 off_scr = take((i64)reduce_scratch_doubles());
  const i64 off_out = take(32);
  arena_ = static_cast<double*>(pool_alloc((size_t)total * sizeof(double)));
  CUDA_CHECK(cudaMemsetAsync(arena_, 0, (size_t)total * sizeof(double), st_));
  for (int g = 0; g < ngrids; ++g) {
    lv_[g].u = arena_ + off_u[g];
    lv_[g].rhs = (g > 0) ? arena_ + off_rhs[g] : nullptr;
  }
  r_ = arena_ + off_r;
  usav_ = arena_ + off_sav;
  scratch_ = arena_ + off_scr;
  d_out_ = arena_ + off_out;
  d_info_ = reinterpret_cast<int*>(arena_ + off_out + 8);
  h_out_ = static_cast<double*>(pool_alloc_host(8 * sizeof(double)));

  // --- pack every transfer table into one int and one double buffer: two uploads per hierarchy
  std::vector<int> hi;
  std::vector<double> hd;
  struct Off { size_t lo, wl, wh, first, count, c2; };
  std::vector<Off> offs((size_t)ngrids * 3);
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  for (int g = 0; g + 1 < ngrids; ++g)
    for (int d = 0; d < 3; ++d) {
      Off& o = offs[(size_t)g * 3 + d];
      auto put_i = [&](const std::vector<int>& v) { size_t at = hi.size(); hi.insert(hi.end(), v.begin(), v.end()); hi.resize(pad(hi.size())); return at; };
      auto put_d = [&](const std::vector<double>& v) { size_t at = hd.size(); hd.insert(hd.end(), v.begin(), v.end()); hd.resize(pad(hd.size())); return at; };
      o.lo = put_i(hl[g].lo[d]); o.first = put_i(hl[g].first[d]); o.count = put_i(hl[g].count[d]);
      o.wl = put_d(hl[g].wl[d]); o.wh = put_d(hl[g].wh[d]); o.c2 = put_d(hl[g].c2[d]);
    }
  tab_i_ = static_cast<int*>(pool_alloc((hi.size() + 32) * sizeof(int)));
  tab_d_ = static_cast<double*>(pool_alloc((hd.size() + 32) * sizeof(double)));
  if (!hi.empty()) CUDA_CHECK(cudaMemcpyAsync(tab_i_, hi.data(), hi.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
  if (!hd.empty()) CUDA_CHECK(cudaMemcpyAsync(tab_d_, hd.data(), hd.size() * sizeof(double), cudaMemcpyHostToDevice, st_));
  for (int g = 0; g + 1 < ngrids; ++g)
    for (int d = 0; d < 3; ++d) {
      const Off& o = offs[(size_t)g * 3 + d];
      lv_[g].it[d] = InterpTab{tab_i_ + o.lo, tab_d_ + o.wl, tab_d_ + o.wh};
      lv_[g].rt[d] = RestrictTab{tab_i_ + o.first, tab_i_ + o.count, tab_d_ + o.c2, hl[g].w2[d]};
    }
  CUDA_CHECK(cudaStreamSynchronize(st_));  // host staging vectors go out of scope
  set_options(5, 1e-13, "NNNNNN", true, 10000);
}

MG::~MG() {
  pool_free(tab_i_);
  pool_free(tab_d_);
  pool_free(arena_);
  pool_free_host(h_out_);
}

void MG::set_options(int ms, double ex_tol, const char* copt, bool du_max, int nmax_exact) {
  ms_ = ms;
  ex_tol_ = ex_tol;
  du_max_ = du_max;
  nmax_exact_ = nmax_exact;
  for (int d = 0; d < 2 * ndim_; ++d) copt_[d] = copt[d];
  all_neumann_ = true;
  for (int d = 0; d < 2 * ndim_; ++d)
    if (copt_[d] != 'N') all_neumann_ = false;
  // colour of the first pass: ndsm_optimized.f90:106 (3D, depends on the x-lower BC); ndsm_poisson.f90:499-501 (2D)
  first_colour_ = (ndim_ == 3 && copt_[0] == 'D') ? 1 : 0;
  for (auto& L : lv_) {
    const int n[3] = {L.g.nx, L.g.ny, L.g.nz};
    for (int d = 0; d < 3; ++d) {
      L.b.lb[d] = 0;
      L.b.ub[d] = n[d] - 1;
      if (d < ndim_) {
        if (copt_[d] == 'D') L.b.lb[d] = 1;                 // bcs(d,1) = copt(d)
        if (copt_[ndim_ + d] == 'D') L.b.ub[d] = n[d] - 2;  // bcs(d,2) = copt(ndim+d)
      }
    }
  }
}

void MG::relax(int g) {
  Level& L = lv_[g];
  const double* rhs = (g == 0) ? rhs0_ : L.rhs;
  if (ndim_ == 3) {
    // each colour pass is timed separately when profiling (PROF_RELAX0 = one k_relax3d launch on level 0)
    if (g == 0) prof_begin(PROF_RELAX0, st_);
    relax3d_half(L.u, rhs, L.g, L.b, first_colour_, L.w, st_);
    if (g == 0) { prof_end(PROF_RELAX0, st_); prof_begin(PROF_RELAX0, st_); }
    relax3d_half(L.u, rhs, L.g, L.b, first_colour_ ^ 1, L.w, st_);
    if (g == 0) prof_end(PROF_RELAX0, st_);
  } else {
    relax2d_half(L.u, rhs, L.g, L.b, 0, L.w, st_);
    relax2d_half(L.u, rhs, L.g, L.b, 1, L.w, st_);
  }
  if (all_neumann_) subtract_mean(L.u, L.g, scratch_, st_);
}

void MG::residual(int g) {
  Level& L = lv_[g];
  const double* rhs = (g == 0) ? rhs0_ : L.rhs;
  const bool prof = (g == 0 && ndim_ == 3);
  if (prof) prof_begin(PROF_RESID0, st_);
  if (ndim_ == 3) residual3d(L.u, rhs, r_, L.g, L.b, L.w, st_);
  else residual2d(L.u, rhs, r_, L.g, L.b, L.w, st_);
  if (prof) prof_end(PROF_RESID0, st_);
}

void MG::restrict_to(int g) {
  Level& F = lv_[g];
  Level& C = lv_[g + 1];
  const bool prof = (g == 0 && ndim_ == 3);
  if (prof) prof_begin(PROF_RESTRICT0, st_);
  restrict_level(r_, F.g, C.rhs, C.g, F.rt[0], F.rt[1], F.rt[2], st_);
  if (prof) prof_end(PROF_RESTRICT0, st_);
  CUDA_CHECK(cudaMemsetAsync(C.u, 0, (size_t)2 * C.g.cs * sizeof(double), st_));  // ndsm_multigrid_core.f90:557-558
}

void MG::interp_add_from(int c) {
  Level& C = lv_[c];
  Level& F = lv_[c - 1];
  const bool prof = (c == 1 && ndim_ == 3);
  if (prof) prof_begin(PROF_INTERP0, st_);
  interp_add(C.u, C.g, F.u, F.g, F.it[0], F.it[1], F.it[2], st_);
  if (prof) prof_end(PROF_INTERP0, st_);
}

// solve_exact (ndsm_multigrid_core.f90:728-800)
int MG::solve_exact(int g) {
  Level& L = lv_[g];
  const double* rhs = (g == 0) ? rhs0_ : L.rhs;
  if (rhs && solve_exact_smem(ndim_, L.u, rhs, L.g, L.b, first_colour_, L.w, all_neumann_, du_max_, ex_tol_,
                              nmax_exact_, d_info_, st_))
    return -1;  // result in d_info_
  // fallback: level too large for one block's shared memory -> host-driven loop with the same semantics
  CUDA_CHECK(cudaMemsetAsync(usav_, 0, (size_t)2 * L.g.cs * sizeof(double), st_));
  double du = HUGE_VAL;
  int it = 0, converged = 0;
  const double N = (double)((i64)L.g.nx * L.g.ny * L.g.nz);
  for (int i = 0; i < nmax_exact_; ++i) {
    if (du <= ex_tol_) { converged = 1; break; }
    relax(g);
    diff_reduce(usav_, L.u, L.g, true, scratch_, d_out_, st_);
    CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, 2 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    du = du_max_ ? h_out_[0] : h_out_[1] / N;
    ++it;
  }
  int info[2] = {it, converged};
  CUDA_CHECK(cudaMemcpyAsync(d_info_, info, sizeof info, cudaMemcpyHostToDevice, st_));
  CUDA_CHECK(cudaStreamSynchronize(st_));
  return it;
}

void MG::v_cycle() {  // ndsm_multigrid_core.f90:341-377
  const int ng = (int)lv_.size();
  for (int g = 0; g < ng - 1; ++g) {  // fine_to_coarse :482-560
    for (int s = 0; s < ms_; ++s) relax(g);
    residual(g);
    restrict_to(g);
  }
  solve_exact(ng - 1);
  for (int c = ng - 1; c >= 1; --c) {  // coarse_to_fine :593-684
    for (int s = 0; s < ms_; ++s) relax(c);
    interp_add_from(c);
    for (int s = 0; s < ms_; ++s) relax(c - 1);
  }
}

int MG::last_nexact() {
  int info[2];
  CUDA_CHECK(cudaMemcpyAsync(info, d_info_, sizeof info, cudaMemcpyDeviceToHost, st_));
  CUDA_CHECK(cudaStreamSynchronize(st_));
  return info[0];
}

// true when solve_exact() on level g will take the single-block shared-memory path (no host sync)
bool MG::coarsest_in_smem(const double* rhs_coarsest) const {
  const Grid& g = lv_.back().g;
  return rhs_coarsest != nullptr && (size_t)g.nx * g.ny * g.nz * 3 * sizeof(double) <= 200 * 1024 && g.nzl == g.nz;
}

// one iteration of solve_poisson_bvp's loop body, enqueue only: V-cycle, update_u, results to pinned memory
void MG::enqueue_cycle(double* u) {
  Level& L0 = lv_[0];
  v_cycle();
  if (ndim_ == 3) prof_begin(PROF_DIFF0, st_);
  diff_reduce(u, L0.u, L0.g, true, scratch_, d_out_, st_);  // update_u :122
  if (ndim_ == 3) prof_end(PROF_DIFF0, st_);
  CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, 2 * sizeof(double), cudaMemcpyDeviceToHost, st_));
  CUDA_CHECK(cudaMemcpyAsync(h_out_ + 2, d_info_, 2 * sizeof(int), cudaMemcpyDeviceToHost, st_));
}

// solve_poisson_bvp (ndsm_poisson.f90:63-155), split into begin / enqueue / poll / end so that several
// independent solves (the six chi faces) can be interleaved on their own streams by one host thread.
void MG::solve_begin(double* u, const double* rhs, double vc_tol, int nmax, SolveTrace* tr) {
  Level& L0 = lv_[0];
  const size_t bytes0 = (size_t)2 * L0.g.cs * sizeof(double);
  ss_ = SolveState();
  ss_.u = u; ss_.vc_tol = vc_tol; ss_.nmax = nmax; ss_.tr = tr;
  if (!rhs && (ndim_ == 2 || lv_.size() == 1)) {  // kernels of those paths always read rhs
    ss_.zero_rhs = static_cast<double*>(pool_alloc(bytes0));
    CUDA_CHECK(cudaMemsetAsync(ss_.zero_rhs, 0, bytes0, st_));
    rhs = ss_.zero_rhs;
  }
  rhs0_ = rhs;
  CUDA_CHECK(cudaMemcpyAsync(L0.u, u, bytes0, cudaMemcpyDeviceToDevice, st_));  // :100

  // The loop body is a static launch sequence (the coarsest solve iterates inside one kernel), so it is
  // captured once into a CUDA graph and replayed every V-cycle: ~250-500 launches per cycle otherwise.
  static const bool graphs_on = !(getenv("NDSM_B200_GRAPH") && atoi(getenv("NDSM_B200_GRAPH")) == 0);
  const double* rhs_coarsest = (lv_.size() == 1) ? rhs0_ : lv_.back().rhs;
  if (graphs_on && !prof_enabled() && nmax > 1 && coarsest_in_smem(rhs_coarsest)) {
    solve_exact_prepare();
    const unsigned long long l0 = g_launches;
    cudaGraph_t graph = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_cycle(u);
    } catch (...) {
      cudaStreamEndCapture(st_, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(st_, &graph));
    ss_.graph_launches = g_launches - l0;
    g_launches = l0;  // nothing ran during capture
    cudaError_t e = cudaGraphInstantiate(&ss_.gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { ss_.gexec = nullptr; cudaGetLastError(); }
  }
  if (g_debug) debug_msg("solve_poisson_bvp", "Performing V cycles...");
  if (nmax <= 0) ss_.done = true;
}

void MG::solve_enqueue() {
  if (ss_.done) return;
  if (ss_.gexec) {
    CUDA_CHECK(cudaGraphLaunch(ss_.gexec, st_));
    g_launches += ss_.graph_launches;
  } else {
    enqueue_cycle(ss_.u);
  }
}

bool MG::solve_poll() {
  if (ss_.done) return true;
  const Level& L0 = lv_[0];
  const double N = (double)((i64)L0.g.nx * L0.g.ny * L0.g.nz);
  CUDA_CHECK(cudaStreamSynchronize(st_));
  prof_collect();
  ss_.du = du_max_ ? h_out_[0] : h_out_[1] / N;
  const int* info = reinterpret_cast<const int*>(h_out_ + 2);
  if (ss_.tr) { ss_.tr->du.push_back(ss_.du); ss_.tr->nexact.push_back(info[0]); }
  if (!info[1]) printf(" Warning: IOPT_NMAXEX exceeded. Coarse-mesh solution may not have converged\n");
  if (g_debug) {
    char s[64];
    snprintf(s, sizeof s, "Solution delta: %12.4E", ss_.du);
    debug_msg("solve_poisson_bvp", s);
  }
  ++ss_.it;
  if (ss_.du < ss_.vc_tol) { ss_.converged = true; ss_.done = true; }  // :136 strict <
  else if (ss_.it >= ss_.nmax) ss_.done = true;
  return ss_.done;
}

int MG::solve_end(double* du_last) {
  if (ss_.gexec) cudaGraphExecDestroy(ss_.gexec);
  ss_.gexec = nullptr;
  if (du_last) *du_last = ss_.du;
  int ierr = 0;
  if (!ss_.converged) {
    ierr = 1;
    printf(" Warning: IOPT_NCYCLES exceeded. V-cycle iteration may not have converged\n");
  }
  if (ss_.tr) ss_.tr->ierr = ierr;
  rhs0_ = nullptr;
  if (ss_.zero_rhs) { CUDA_CHECK(cudaStreamSynchronize(st_)); pool_free(ss_.zero_rhs); ss_.zero_rhs = nullptr; }
  return ierr;
}

int MG::solve(double* u, const double* rhs, double vc_tol, int nmax, double* du_last, SolveTrace* tr) {
  solve_begin(u, rhs, vc_tol, nmax, tr);
  while (!ss_.done) {
    solve_enqueue();
    solve_poll();
  }
  return solve_end(du_last);
}

}  // namespace ndsm
