// mg.cu -- multigrid hierarchy, transfer tables and the V-cycle loop (host orchestration).
// See mg.hpp for the reference functions this mirrors.
#include "mg.hpp"
#include "pool.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
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

// ---------------------------------------------------------------------------------------------
// z-slab partition (host only)
// ---------------------------------------------------------------------------------------------
SlabPlan plan_slabs(const std::vector<HostLevel>& hl, int ndim, int world, int min_planes, long long min_points) {
  SlabPlan p;
  p.world = world;
  if (world <= 1 || ndim != 3) return p;
  const int ng = (int)hl.size();
  // halo depth: a deeper halo trades redundant smoothing of halo planes for fewer exchanges (one exchange
  // buys about halo-1 colour passes); NDSM_HALO_PLANES overrides the default
  int halo = NDSM_HALO;
  if (const char* e = std::getenv("NDSM_HALO_PLANES")) halo = std::max(4, std::min(64, std::atoi(e)));
  if (min_planes < halo) min_planes = halo;
  std::vector<int> z(world + 1);
  for (int r = 0; r <= world; ++r) z[r] = (int)((i64)hl[0].n[2] * r / world);  // balanced finest slabs
  int g = 0;
  while (true) {
    bool ok = (g < ng - 1);  // the coarsest level is always replicated (solve_exact needs the whole grid)
    // small levels are cheaper to solve redundantly on every rank than to exchange halos for
    if ((long long)hl[g].n[0] * hl[g].n[1] * hl[g].n[2] < min_points) ok = false;
    for (int r = 0; r < world && ok; ++r) ok = (z[r + 1] - z[r] >= min_planes);
    if (!ok) break;
    p.zs.push_back(z);
    p.ndist = g + 1;
    // coarse plane kc belongs to the rank that owns its anchor fine plane (centre of its restriction stencil)
    const std::vector<int>& first = hl[g].first[2];
    const std::vector<int>& count = hl[g].count[2];
    const int nc = hl[g + 1].n[2];
    std::vector<int> zc(world + 1, 0);
    zc[world] = nc;
    for (int r = 1; r < world; ++r) {
      int c = 0;
      while (c < nc && first[c] + (count[c] - 1) / 2 < z[r]) ++c;
      zc[r] = c;
    }
    z = zc;
    ++g;
  }
  if (p.ndist == 0) return p;
  p.zs.push_back(z);  // producer partition of the first replicated level
  p.halo = halo;
  // every stencil must stay inside owned planes + halo
  for (int lv = 0; lv < p.ndist; ++lv) {
    const int nzf = hl[lv].n[2], ncz = hl[lv + 1].n[2];
    for (int r = 0; r < world; ++r) {
      // one plane less than the halo is allowed for the stencil (head-room kept from an earlier fused
      // residual+restriction, which read u one plane further out; the exchange depth actually used is rneed_)
      const int f0 = p.zs[lv][r] - (p.halo - 1), f1 = p.zs[lv][r + 1] + (p.halo - 1);
      for (int c = p.zs[lv + 1][r]; c < p.zs[lv + 1][r + 1]; ++c) {
        const int a = hl[lv].first[2][c], b = a + hl[lv].count[2][c];
        if (a < (f0 < 0 ? 0 : f0) || b > (f1 > nzf ? nzf : f1)) throw NdsmError(6);
      }
      if (lv + 1 < p.ndist) {
        const int c0 = p.zs[lv + 1][r] - p.halo, c1 = p.zs[lv + 1][r + 1] + p.halo;
        for (int k = p.zs[lv][r]; k < p.zs[lv][r + 1]; ++k) {
          const int lo = hl[lv].lo[2][k], hi = (lo + 1 < ncz) ? lo + 1 : ncz - 1;
          if (lo < c0 || hi >= c1) throw NdsmError(6);
        }
      }
    }
  }
  return p;
}

void* Comm::sym_alloc(size_t bytes) { return pool_alloc(bytes); }
void Comm::sym_free(void* p) { pool_free(p); }

// ---------------------------------------------------------------------------------------------
// virtual communication: every rank lives in this process on one device; messages are device copies
// ---------------------------------------------------------------------------------------------
namespace {
struct VirtualComm : Comm {
  struct Msg { int from, to; const double* src; double* dst; size_t n; bool used; };
  int w;
  std::vector<Msg> sends, recvs;
  explicit VirtualComm(int world) : w(world) {}
  int world() const override { return w; }
  int first_rank() const override { return 0; }
  int nlocal() const override { return w; }
  void begin(cudaStream_t) override { sends.clear(); recvs.clear(); }
  void send(int my_rank, int to_rank, const double* src, size_t n, cudaStream_t) override {
    sends.push_back(Msg{my_rank, to_rank, src, nullptr, n, false});
  }
  void recv(int my_rank, int from_rank, double* dst, size_t n, cudaStream_t) override {
    recvs.push_back(Msg{from_rank, my_rank, nullptr, dst, n, false});
  }
  void end(cudaStream_t st) override {
    for (auto& r : recvs) {
      bool found = false;
      for (auto& s : sends)
        if (!s.used && s.from == r.from && s.to == r.to) {
          if (s.n != r.n) throw NdsmError(6);
          CUDA_CHECK(cudaMemcpyAsync(r.dst, s.src, s.n * sizeof(double), cudaMemcpyDeviceToDevice, st));
          s.used = true;
          found = true;
          break;
        }
      if (!found) throw NdsmError(6);
    }
  }
  void gathern(int my_rank, const double* send, int n, double* recv_all, cudaStream_t st) override {
    CUDA_CHECK(cudaMemcpyAsync(recv_all + (size_t)n * my_rank, send, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  void bcast(int, double*, size_t, cudaStream_t) override {}
  std::unique_ptr<Comm> clone(cudaStream_t) override { return std::unique_ptr<Comm>(new VirtualComm(w)); }
  std::unique_ptr<Comm> split(int) override { return nullptr; }
  const char* transport() const override { return "virtual ranks on one device (device copies)"; }
};
}  // namespace
std::unique_ptr<Comm> make_virtual_comm(int world) { return std::unique_ptr<Comm>(new VirtualComm(world)); }

// which levels are worth partitioning (the defaults and their environment overrides)
void slab_policy(int world, int* min_planes, long long* min_points) {
  *min_planes = 16;
  // measured at 513^3: level 1 (256^3) costs 1.66 ms per V-cycle replicated and 1.52 ms partitioned over two
  // ranks (8 halo exchanges included); 128^3 and below are cheaper replicated than exchanged
  *min_points = 8000000;
  // ... for up to four ranks.  From eight ranks on a replicated 129^3 level is seven eighths redundant work: measured
  // at 513^3 on 8 GPUs, partitioning it as well takes the three solves from 64.6 to 62.7 ms (same results, bit for bit)
  if (world >= 8) *min_points = 1000000;
  if (const char* e = getenv("NDSM_SLAB_MIN_PLANES")) *min_planes = atoi(e);
  if (const char* e = getenv("NDSM_SLAB_MIN_POINTS")) *min_points = atoll(e);
}

// ---------------------------------------------------------------------------------------------
// MG
// ---------------------------------------------------------------------------------------------
static Grid slab_grid(const Grid& full, int k0, int nzl, int H) {
  Grid g = full;
  g.k0 = k0;
  g.nzl = nzl;
  g.cs = full.ps * (nzl + 2 * H);
  return g;
}

MG::MG(int ndim, const int* shape, int ngrids, const double* const* mesh, cudaStream_t st, Comm* comm)
    : ndim_(ndim), comm_(comm), st_(st) {
  std::vector<HostLevel> hl = build_hierarchy(ndim, shape, ngrids, mesh);
  ngrids = (int)hl.size();
  std::memset(copt_, 'N', sizeof copt_);
  mesh_.resize(ngrids);
  for (int g = 0; g < ngrids; ++g) {
    mesh_[g].resize(3);
    for (int d = 0; d < 3; ++d) mesh_[g][d] = hl[g].mesh[d];
  }
  const int world = comm_ ? comm_->world() : 1;
  int min_planes;
  long long min_points;
  slab_policy(world, &min_planes, &min_points);
  plan_ = plan_slabs(hl, ndim, world, min_planes, min_points);
  valid_.assign(ngrids, std::array<int, 2>{{0, 0}});
  static_ok_.assign(ngrids, std::array<bool, 2>{{false, false}});
  const int nd = plan_.ndist;
  const int H = plan_.halo;
  // halo planes the transfers really read (the same number on every rank, so that both sides of an exchange agree):
  // restriction reads r of the fine level over the z windows of the coarse planes a rank produces, prolongation
  // reads the two coarse planes that bracket each owned fine plane
  rneed_.assign(ngrids, H);
  ineed_.assign(ngrids, H);
  for (int g = 0; g < nd; ++g) {
    int rn = 0, in = 0;
    for (int r = 0; r < world; ++r) {
      const int f0 = plan_.zs[g][r], f1 = plan_.zs[g][r + 1];
      for (int c = plan_.zs[g + 1][r]; c < plan_.zs[g + 1][r + 1]; ++c) {
        const int a = hl[g].first[2][c], b = a + hl[g].count[2][c];
        rn = std::max(rn, std::max(f0 - a, b - f1));
      }
      if (g + 1 < nd) {
        const int c0 = plan_.zs[g + 1][r], c1 = plan_.zs[g + 1][r + 1], ncz = hl[g + 1].n[2];
        for (int k = f0; k < f1; ++k) {
          const int lo = hl[g].lo[2][k], hi = (lo + 1 < ncz) ? lo + 1 : ncz - 1;
          in = std::max(in, std::max(c0 - lo, hi + 1 - c1));
        }
      }
    }
    rneed_[g] = std::min(H, std::max(rn, 1));
    if (g + 1 < nd) ineed_[g + 1] = std::min(H, std::max(in, 1));
  }
  if (getenv("NDSM_HALO_EXACT_DEPTH") && atoi(getenv("NDSM_HALO_EXACT_DEPTH")) == 0) {  // full-depth transfers exchanges
    rneed_.assign(ngrids, H);
    ineed_.assign(ngrids, H);
  }
  const int nlocal = (nd > 0 && comm_) ? comm_->nlocal() : 1;
  slabs_.resize(nlocal);
  rhs0_.assign(nlocal, nullptr);

  // --- shared arena: replicated levels, coarsest u_sav, reduction scratch, results
  i64 total = 0;
  auto take = [&](i64 n) { i64 o = total; total += round_up(n, 32); return o; };
  std::vector<i64> off_u(ngrids, -1), off_rhs(ngrids, -1);
  for (int g = nd; g < ngrids; ++g) {
    off_u[g] = take(2 * hl[g].g.cs);
    if (g > 0) off_rhs[g] = take(2 * hl[g].g.cs);
  }
  const i64 off_sav = take(2 * hl[ngrids - 1].g.cs);
  const i64 off_scr = take((i64)reduce_scratch_doubles());
  // pure-Neumann 2D sweeps with the mean subtraction folded into the passes (measured: chi V-cycle loop 16.4 ->
  // 13.5 ms at 513^2); NDSM_B200_FUSED_MEAN=0 selects the separate reduction kernels (read per hierarchy)
  const bool fused_mean = ndim == 2 && !(getenv("NDSM_B200_FUSED_MEAN") && atoi(getenv("NDSM_B200_FUSED_MEAN")) == 0);
  const i64 off_fm = fused_mean ? take((i64)relax2d_fused_mean_scratch(hl[0].g)) : -1;
  const i64 off_all = take(2 * (i64)world + 8);
  const i64 off_allm = take(2 * (2 * (i64)world + 8));  // two alternating buffers of gathered slab sums (pure Neumann)
  const i64 off_info = take(8);
  // bcast() / gather2() write into this arena on the other ranks (replicated levels, gathered (max,sum) pairs):
  // it comes from the communicator's symmetric heap (collective; its size does not depend on the rank)
  shared_ = static_cast<double*>(comm_ ? comm_->sym_alloc((size_t)total * sizeof(double))
                                       : pool_alloc((size_t)total * sizeof(double)));
  if (comm_ && nd > 0) comm_->reserve((size_t)2 * H * (size_t)hl[0].g.ps);  // both colours of a full halo
  CUDA_CHECK(cudaMemsetAsync(shared_, 0, (size_t)total * sizeof(double), st_));
  usav_ = shared_ + off_sav;
  scratch_ = shared_ + off_scr;
  fm_scratch_ = (off_fm >= 0) ? shared_ + off_fm : nullptr;
  d_all_ = shared_ + off_all;
  d_allm_[0] = shared_ + off_allm;
  d_allm_[1] = shared_ + off_allm + 2 * (i64)world + 8;
  d_info_ = reinterpret_cast<int*>(shared_ + off_info);
  h_out_ = static_cast<double*>(pool_alloc_host((2 * (size_t)world + 8) * sizeof(double)));
  {  // pinned memory is mapped under unified addressing; kernels write the per-cycle results there directly
    void* dp = nullptr;
    if (world <= 24 && cudaHostGetDevicePointer(&dp, h_out_, 0) == cudaSuccess) h_out_dev_ = static_cast<double*>(dp);
    else cudaGetLastError();
    if (getenv("NDSM_B200_MAPPED_RESULTS") && atoi(getenv("NDSM_B200_MAPPED_RESULTS")) == 0) h_out_dev_ = nullptr;
  }

  // --- per-slab arenas: partitioned levels (with halos), residual scratch, local reduction result
  for (int s = 0; s < nlocal; ++s) {
    Slab& S = slabs_[s];
    S.rank = (comm_ ? comm_->first_rank() : 0) + s;
    S.lv.resize(ngrids);
    i64 tot = 0;
    auto tk = [&](i64 n) { i64 o = tot; tot += round_up(n, 32); return o; };
    std::vector<i64> ou(ngrids, -1), orh(ngrids, -1);
    i64 rmax = 0;
    for (int g = 0; g < ngrids; ++g) {
      Level& L = S.lv[g];
      L.w = hl[g].w;
      if (g < nd) {
        L.dist = true;
        L.H = H;
        L.g = slab_grid(hl[g].g, plan_.zs[g][S.rank], plan_.zs[g][S.rank + 1] - plan_.zs[g][S.rank], H);
        ou[g] = tk(2 * L.g.cs);
        if (g > 0) orh[g] = tk(2 * L.g.cs);
      } else {
        L.dist = false;
        L.H = 0;
        L.g = hl[g].g;
      }
      if (g < ngrids - 1 || ngrids == 1) rmax = std::max<i64>(rmax, 2 * L.g.cs);
    }
    const i64 orr = tk(rmax);
    const i64 oout = tk(8);
    S.arena = static_cast<double*>(pool_alloc((size_t)tot * sizeof(double)));
    CUDA_CHECK(cudaMemsetAsync(S.arena, 0, (size_t)tot * sizeof(double), st_));
    S.r = S.arena + orr;
    S.d_out = S.arena + oout;
    for (int g = 0; g < ngrids; ++g) {
      Level& L = S.lv[g];
      if (g < nd) {
        L.u = S.arena + ou[g] + (i64)H * L.g.ps;
        L.rhs = (g > 0) ? S.arena + orh[g] + (i64)H * L.g.ps : nullptr;
      } else {
        L.u = shared_ + off_u[g];
        L.rhs = (g > 0) ? shared_ + off_rhs[g] : nullptr;
      }
    }
  }

  u0_home_ = slabs_[0].lv[0].u;
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
  static const bool fuse_on = !(getenv("NDSM_B200_TILED") && atoi(getenv("NDSM_B200_TILED")) == 0);
  // read per hierarchy (not cached): tests toggle it between solves
  const bool exact_restrict = getenv("NDSM_B200_EXACT_RESTRICT") && atoi(getenv("NDSM_B200_EXACT_RESTRICT")) != 0;
  for (auto& S : slabs_)
    for (int g = 0; g + 1 < ngrids; ++g) {
      for (int d = 0; d < 3; ++d) {
        const Off& o = offs[(size_t)g * 3 + d];
        S.lv[g].it[d] = InterpTab{tab_i_ + o.lo, tab_d_ + o.wl, tab_d_ + o.wh};
        S.lv[g].rt[d] = RestrictTab{tab_i_ + o.first, tab_i_ + o.count, tab_d_ + o.c2, hl[g].w2[d]};
      }
      // NDSM_B200_INTERP=tiled|simple selects the older prolongation kernels (same bits)
      const char* ipk = getenv("NDSM_B200_INTERP");
      S.lv[g].icols = ndim == 3 && !ipk && hl[g].n[0] >= 64 && hl[g].n[2] >= 8 &&
                      interp_zt_fits(hl[g].lo[0].data(), hl[g].n[0], hl[g + 1].n[0], hl[g].lo[1].data(), hl[g].n[1],
                                     hl[g + 1].n[1]);
      if (ndim == 3 && fuse_on && !(ipk && !strcmp(ipk, "simple")) && hl[g].n[0] >= 64 && hl[g].n[2] >= 8)
        S.lv[g].itiled = interp_tiled_fits(hl[g].lo[0].data(), hl[g].n[0], hl[g + 1].n[0], hl[g].lo[1].data(),
                                           hl[g].n[1], hl[g + 1].n[1]);
      if (ndim == 3 && !exact_restrict && hl[g + 1].n[0] >= 16 && hl[g + 1].n[1] >= 8 && hl[g + 1].n[2] >= 8)
        S.lv[g].rsep = restrict_sep_fits(hl[g].first[0].data(), hl[g].count[0].data(), hl[g + 1].n[0],
                                         hl[g].first[1].data(), hl[g].count[1].data(), hl[g + 1].n[1]);
      if (S.lv[g].rsep && !(getenv("NDSM_B200_RESTRICT") && !strcmp(getenv("NDSM_B200_RESTRICT"), "tile"))) {
        const int* const fi[3] = {hl[g].first[0].data(), hl[g].first[1].data(), hl[g].first[2].data()};
        const int* const co[3] = {hl[g].count[0].data(), hl[g].count[1].data(), hl[g].count[2].data()};
        const int nc[3] = {hl[g + 1].n[0], hl[g + 1].n[1], hl[g + 1].n[2]};
        S.lv[g].rdirect = restrict_direct_fits(fi, co, nc);
      }
      if (ndim == 3 && fuse_on && hl[g + 1].n[0] >= 16 && hl[g + 1].n[1] >= 8)
        S.lv[g].fused = restrict_tiled_fits(hl[g].first[0].data(), hl[g].count[0].data(), hl[g + 1].n[0],
                                            hl[g].first[1].data(), hl[g].count[1].data(), hl[g + 1].n[1],
                                            &S.lv[g].rr_hwp, &S.lv[g].rr_fyw);
    }
  CUDA_CHECK(cudaStreamSynchronize(st_));  // host staging vectors go out of scope
  // --- which levels run inside the single-block small-level kernel
  static const bool small_on = !(getenv("NDSM_B200_SMALL") && atoi(getenv("NDSM_B200_SMALL")) == 0);
  small_from_ = 0;
  if (small_on && ngrids >= 2) {
    int ls = ngrids;
    while (ls - 1 >= 1 && ls - 1 >= nd && (i64)hl[ls - 1].n[0] * hl[ls - 1].n[1] * hl[ls - 1].n[2] <= SMALL_MAX_POINTS &&
           ngrids - (ls - 1) <= SMALL_MAX_LEVELS)
      --ls;
    if (ls < ngrids) small_from_ = ls;
  }
  set_options(5, 1e-13, "NNNNNN", true, 10000);
}

void MG::build_small_args() {
  if (small_from_ <= 0) return;
  SmallArgs& a = small_args_;
  const int ng = ngrids();
  a.nlev = ng - small_from_;
  int off = 0;
  for (int l = 0; l < a.nlev; ++l) {
    const Level& L = slabs_[0].lv[small_from_ + l];
    SmallLevel& S = a.lv[l];
    S.nx = L.g.nx; S.ny = L.g.ny; S.nz = L.g.nz;
    S.w = L.w;
    S.b = L.b;
    for (int d = 0; d < 3; ++d) { S.it[d] = L.it[d]; S.rt[d] = L.rt[d]; }
    const int N = S.nx * S.ny * S.nz;
    S.off_u = off; off += N;
    S.off_rhs = off; off += N;
  }
  a.off_r = off; off += a.lv[0].nx * a.lv[0].ny * a.lv[0].nz;
  a.off_sav = off; off += a.lv[a.nlev - 1].nx * a.lv[a.nlev - 1].ny * a.lv[a.nlev - 1].nz;
  a.smem_doubles = off;
  a.first_colour = first_colour_;
  a.all_neumann = all_neumann_ ? 1 : 0;
  a.du_max = du_max_ ? 1 : 0;
  a.nmax_exact = nmax_exact_;
  a.ms = ms_;
  a.ex_tol = ex_tol_;
  if ((size_t)off * sizeof(double) > 200 * 1024) small_from_ = 0;
}

MG::~MG() {
  drop_graphs();
  pool_free(tab_i_);
  pool_free(tab_d_);
  for (auto& S : slabs_) pool_free(S.arena);
  if (comm_) comm_->sym_free(shared_);
  else pool_free(shared_);
  pool_free_host(h_out_);
}

double* MG::r_scratch(int g, int s) {
  const Level& L = slabs_[s].lv[g];
  return slabs_[s].r + (i64)L.H * L.g.ps;
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
  for (auto& S : slabs_)
    for (auto& L : S.lv) {
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
  build_small_args();
}

// halo planes with the z-neighbours (one grouped exchange)
void MG::exchange(int g, int which, int colour_mask, int np, const std::vector<double*>* arr) {
  if (g >= plan_.ndist || !comm_) return;
  const int world = plan_.world;
  prof_begin(PROF_EXCH, st_);
  comm_->begin(st_);
  for (size_t s = 0; s < slabs_.size(); ++s) {
    Slab& S = slabs_[s];
    const Level& L = S.lv[g];
    double* P = arr ? (*arr)[s] : (which == 0 ? L.u : r_scratch(g, (int)s));
    const size_t n = (size_t)np * L.g.ps;
    for (int c = 0; c < 2; ++c) {
      if (!(colour_mask & (1 << c))) continue;
      double* pc = P + (i64)c * L.g.cs;
      if (S.rank + 1 < world) {
        comm_->send(S.rank, S.rank + 1, pc + (i64)(L.g.nzl - np) * L.g.ps, n, st_);
        comm_->recv(S.rank, S.rank + 1, pc + (i64)L.g.nzl * L.g.ps, n, st_);
      }
      if (S.rank > 0) {
        comm_->send(S.rank, S.rank - 1, pc, n, st_);
        comm_->recv(S.rank, S.rank - 1, pc - (i64)np * L.g.ps, n, st_);
      }
    }
  }
  comm_->end(st_);
  prof_end(PROF_EXCH, st_);
}

void MG::need_halo(int g, int depth) {
  if (g >= plan_.ndist || !comm_) return;
  if (valid_[g][0] >= depth && valid_[g][1] >= depth) return;
  exchange(g, 0, 3, plan_.halo);
  valid_[g][0] = valid_[g][1] = plan_.halo;
  static_ok_[g][0] = static_ok_[g][1] = true;
}

void MG::relax(int g) {
  const bool dist = g < plan_.ndist && comm_;
  const size_t ns = dist ? slabs_.size() : 1;
  if (ndim_ == 3) {
    for (int pass = 0; pass < 2; ++pass) {
      const int colour = first_colour_ ^ pass;
      int ext = 0;
      if (dist) {  // the pass needs the other colour one plane beyond the planes it updates
        need_halo_colour(g, 1 - colour);
        // extended passes read rhs in the halo planes: exchanged for g > 0, identically zero for the
        // vector-potential solves on g == 0; a caller-supplied level-0 rhs has no trusted halo
        const bool rhs_halo_ok = (g > 0) || rhs0_[0] == nullptr || rhs0_halo_ok_;
        ext = rhs_halo_ok ? valid_[g][1 - colour] - 1 : 0;
      }
      // each colour pass is timed separately when profiling (PROF_RELAX0 = one k_relax3d launch on level 0)
      if (g == 0 && ns == 1) prof_begin(PROF_RELAX0, st_);
      for (size_t s = 0; s < ns; ++s) {
        Level& L = slabs_[s].lv[g];
        relax3d_half(L.u, (g == 0) ? rhs0_[s] : L.rhs, L.g, L.b, colour, L.w, ext, st_, (g == 0) ? pp_read_ : nullptr);
      }
      if (g == 0) pp_read_ = nullptr;  // only the first pass of the cycle reads the previous iterate's array
      if (g == 0 && ns == 1) prof_end(PROF_RELAX0, st_);
      if (dist) valid_[g][colour] = ext;
    }
  } else {
    Level& L = slabs_[0].lv[g];
    const double* rhs = (g == 0) ? rhs0_[0] : L.rhs;
    relax2d_half(L.u, rhs, L.g, L.b, 0, L.w, st_);
    relax2d_half(L.u, rhs, L.g, L.b, 1, L.w, st_);
  }
  if (all_neumann_) {  // u -= sum(u)/N after every sweep (ndsm_optimized.f90:173-189, ndsm_poisson.f90:538-541)
    if (dist) {
      // partitioned level: slab sums over the owned planes, gathered in rank order, subtracted from owned and
      // halo planes alike (the neighbour subtracts the same value from the planes my halo mirrors).  Two
      // gathered buffers alternate: a peer may already deliver the next sweep's sum while this one is applied.
      double* pairs = d_allm_[mean_parity_];
      mean_parity_ ^= 1;
      for (auto& S : slabs_) slab_sum(S.lv[g].u, S.lv[g].g, scratch_, S.d_out + 2, st_);
      for (auto& S : slabs_) comm_->gather2(S.rank, S.d_out + 2, pairs, st_);
      for (auto& S : slabs_) subtract_gathered_mean(S.lv[g].u, S.lv[g].g, pairs, plan_.world, S.lv[g].H, st_);
    } else {
      subtract_mean(slabs_[0].lv[g].u, slabs_[0].lv[g].g, scratch_, st_);
    }
  }
}

void MG::relax_sweeps(int g, int n) {
  if (fm_scratch_ && ndim_ == 2 && all_neumann_ && n > 0) {
    Level& L = slabs_[0].lv[g];
    relax2d_fused_mean(L.u, (g == 0) ? rhs0_[0] : L.rhs, L.g, L.b, L.w, n, fm_scratch_, st_);
    return;
  }
  for (int s = 0; s < n; ++s) relax(g);
}

// A pass needs the OTHER colour at least one plane deep.  When it is not, only that colour has to be exchanged:
// the pass that follows recomputes its own colour in the halo planes (extended pass) from exactly these values,
// so the own colour's halo is overwritten before anything reads it -- except its Dirichlet points, which no pass
// updates.  Those are current in the halo once the colour has been exchanged at full depth since the level was
// last changed from outside the smoother (start of a solve, prolongation: the correction is added on Dirichlet
// faces too, ndsm_multigrid_core.f90:706-710); until then both colours travel.  Half the bytes otherwise.
void MG::need_halo_colour(int g, int colour) {
  if (valid_[g][colour] >= 1) return;
  const bool one_colour = !(getenv("NDSM_HALO_ONE_COLOUR") && atoi(getenv("NDSM_HALO_ONE_COLOUR")) == 0);
  if (!one_colour || !static_ok_[g][1 - colour]) {
    exchange(g, 0, 3, plan_.halo);
    valid_[g][0] = valid_[g][1] = plan_.halo;
    static_ok_[g][0] = static_ok_[g][1] = true;
    return;
  }
  exchange(g, 0, 1 << colour, plan_.halo);
  valid_[g][colour] = plan_.halo;
  static_ok_[g][colour] = true;
}

void MG::residual(int g) {
  need_halo(g, 1);
  const size_t ns = (g < plan_.ndist) ? slabs_.size() : 1;
  const bool prof = (g == 0 && ndim_ == 3 && ns == 1);
  if (prof) prof_begin(PROF_RESID0, st_);
  for (size_t s = 0; s < ns; ++s) {
    Level& L = slabs_[s].lv[g];
    const double* rhs = (g == 0) ? rhs0_[s] : L.rhs;
    if (ndim_ == 3) residual3d(L.u, rhs, r_scratch(g, (int)s), L.g, L.b, L.w, st_);
    else residual2d(L.u, rhs, r_scratch(g, (int)s), L.g, L.b, L.w, st_);
  }
  if (prof) prof_end(PROF_RESID0, st_);
}

void MG::restrict_to(int g) {
  const int c = g + 1;
  const bool fdist = g < plan_.ndist, cdist = c < plan_.ndist;
  const size_t ns = fdist ? slabs_.size() : 1;
  if (fdist) exchange(g, 2, 3, rneed_[g]);
  const bool prof = (g == 0 && ndim_ == 3 && ns == 1);
  if (prof) prof_begin(PROF_RESTRICT0, st_);
  for (size_t s = 0; s < ns; ++s) {
    Level& F = slabs_[s].lv[g];
    Level& C = slabs_[s].lv[c];
    Grid gv = C.g;
    double* out = C.rhs;
    if (fdist && !cdist) {  // partitioned -> replicated: this rank produces planes [zs[c][r], zs[c][r+1]) of the full array
      const int r = slabs_[s].rank;
      gv.k0 = plan_.zs[c][r];
      gv.nzl = plan_.zs[c][r + 1] - gv.k0;
      out = C.rhs + (i64)gv.k0 * C.g.ps;
    }
    if (gv.nzl <= 0) continue;
    if (F.rdirect)
      restrict_direct(r_scratch(g, (int)s), F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], st_);
    else if (F.rsep)
      restrict_sep(r_scratch(g, (int)s), F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], st_);
    else if (F.fused)
      restrict_tiled(r_scratch(g, (int)s), F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], F.rr_hwp, F.rr_fyw, st_);
    else
      restrict_level(r_scratch(g, (int)s), F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], st_);
  }
  if (prof) prof_end(PROF_RESTRICT0, st_);
  finish_restrict(g);
}

// fine_to_coarse without the residual array (K2+K3 fused, kernels.cu): usable where the direct restriction is, and
// where rhs is valid in the halo planes the z windows reach into
bool MG::fused_restrict_ok(int g) const {
  if (ndim_ != 3 || g + 1 >= ngrids()) return false;
  // opt-in: measured on B200 at 513^3 the fused pair takes 0.655 + 0.370 ms against 0.419 + 0.508 ms for
  // k_residual3d + k_restrict_direct -- k_residual_rz is instruction-bound (310 instructions per warp and plane, a
  // third of them register moves and selects of the 4-point bookkeeping), so saving 8 B/point does not pay yet
  if (!(getenv("NDSM_B200_FUSED_RESTRICT") && atoi(getenv("NDSM_B200_FUSED_RESTRICT")) != 0)) return false;
  if (!slabs_[0].lv[g].rdirect) return false;
  const bool fdist = g < plan_.ndist && comm_;
  if (fdist && g == 0 && rhs0_[0] != nullptr && !rhs0_halo_ok_) return false;
  return true;
}

void MG::residual_restrict_to(int g) {
  const int c = g + 1;
  const bool fdist = g < plan_.ndist && comm_, cdist = c < plan_.ndist;
  const size_t ns = (g < plan_.ndist) ? slabs_.size() : 1;
  // the residual is evaluated on every fine plane of the z windows (rneed_ planes into the halo), which reads u
  // one plane further out
  if (fdist) need_halo(g, rneed_[g] + 1);
  const bool prof = (g == 0 && ndim_ == 3 && ns == 1);
  if (prof) prof_begin(PROF_RESTRICT0, st_);
  for (size_t s = 0; s < ns; ++s) {
    Level& F = slabs_[s].lv[g];
    Level& C = slabs_[s].lv[c];
    Grid gv = C.g;
    double* out = C.rhs;
    if ((g < plan_.ndist) && !cdist) {  // partitioned -> replicated: this rank produces planes [zs[c][r], zs[c][r+1])
      const int r = slabs_[s].rank;
      gv.k0 = plan_.zs[c][r];
      gv.nzl = plan_.zs[c][r + 1] - gv.k0;
      out = C.rhs + (i64)gv.k0 * C.g.ps;
    }
    if (gv.nzl <= 0) continue;
    residual_restrict(F.u, (g == 0) ? rhs0_[s] : F.rhs, F.g, F.b, F.w, slabs_[s].r, out, gv, F.rt[0], F.rt[1], F.rt[2], st_);
  }
  if (prof) prof_end(PROF_RESTRICT0, st_);
  finish_restrict(g);
}

// all-gather of a replicated coarse rhs, halo planes of a partitioned one, u[c] = 0
void MG::finish_restrict(int g) {
  const int c = g + 1;
  const bool fdist = g < plan_.ndist, cdist = c < plan_.ndist;
  if (fdist && !cdist && comm_) {  // all-gather the replicated rhs
    Level& C = slabs_[0].lv[c];
    comm_->begin(st_);
    for (int q = 0; q < plan_.world; ++q) {
      const int k0 = plan_.zs[c][q], cnt = plan_.zs[c][q + 1] - k0;
      if (cnt <= 0) continue;
      for (int col = 0; col < 2; ++col)
        comm_->bcast(q, C.rhs + (i64)col * C.g.cs + (i64)k0 * C.g.ps, (size_t)cnt * C.g.ps, st_);
    }
    comm_->end(st_);
  }
  if (cdist) {  // extended colour passes read rhs in the halo planes
    std::vector<double*> rp;
    for (auto& S : slabs_) rp.push_back(S.lv[c].rhs);
    exchange(c, 0, 3, plan_.halo, &rp);
    valid_[c][0] = valid_[c][1] = plan_.halo;  // u[c] = 0 everywhere, halos included
    static_ok_[c][0] = static_ok_[c][1] = true;
  }
  const size_t nc = cdist ? slabs_.size() : 1;
  for (size_t s = 0; s < nc; ++s) {  // ndsm_multigrid_core.f90:557-558
    Level& C = slabs_[s].lv[c];
    CUDA_CHECK(cudaMemsetAsync(level_base(C.u, c, (int)s), 0, (size_t)2 * C.g.cs * sizeof(double), st_));
  }
}

void MG::interp_add_from(int c) {
  const int f = c - 1;
  const bool fdist = f < plan_.ndist, cdist = c < plan_.ndist;
  const size_t ns = fdist ? slabs_.size() : 1;
  if (cdist && (valid_[c][0] < ineed_[c] || valid_[c][1] < ineed_[c])) {
    exchange(c, 0, 3, ineed_[c]);
    valid_[c][0] = std::max(valid_[c][0], ineed_[c]);
    valid_[c][1] = std::max(valid_[c][1], ineed_[c]);
  }
  const bool prof = (c == 1 && ndim_ == 3 && ns == 1);
  if (prof) prof_begin(PROF_INTERP0, st_);
  for (size_t s = 0; s < ns; ++s) {
    Level& C = slabs_[s].lv[c];
    Level& F = slabs_[s].lv[f];
    if (F.icols) interp_add_zt(C.u, C.g, F.u, F.g, F.it[0], F.it[1], F.it[2], st_);
    else if (F.itiled) interp_add_tiled(C.u, C.g, F.u, F.g, F.it[0], F.it[1], F.it[2], st_);
    else interp_add(C.u, C.g, F.u, F.g, F.it[0], F.it[1], F.it[2], st_);
  }
  if (prof) prof_end(PROF_INTERP0, st_);
  if (fdist) {  // owned planes changed (Dirichlet faces included); halos are refreshed by the next consumer
    valid_[f][0] = valid_[f][1] = 0;
    static_ok_[f][0] = static_ok_[f][1] = false;
  }
}

// solve_exact (ndsm_multigrid_core.f90:728-800) -- always on a replicated level
int MG::solve_exact(int g) {
  Level& L = slabs_[0].lv[g];
  const double* rhs = (g == 0) ? rhs0_[0] : L.rhs;
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
    diff_reduce(usav_, L.u, L.g, true, scratch_, slabs_[0].d_out, st_);
    CUDA_CHECK(cudaMemcpyAsync(h_out_, slabs_[0].d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, st_));
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
  const int ng = ngrids();
  // profiling brackets (3D only): PROF_LEVEL1 = all work on level 1, PROF_TAIL = all work on levels >= 2
  const bool pr = (ndim_ == 3 && ng > 3 && (small_from_ == 0 || small_from_ > 2));
  const int gend = (small_from_ > 0) ? small_from_ : ng - 1;
  for (int g = 0; g < gend; ++g) {  // fine_to_coarse :482-560
    if (pr && g == 1) prof_begin(PROF_LEVEL1, st_);
    if (pr && g == 2) { prof_end(PROF_LEVEL1, st_); prof_begin(PROF_TAIL, st_); }
    relax_sweeps(g, ms_);
    if (fused_restrict_ok(g)) {
      residual_restrict_to(g);
    } else {
      residual(g);
      restrict_to(g);
    }
  }
  const int ls = small_from_;
  int cstart = ng - 1;
  if (ls > 0) {
    // levels >= ls: the whole sub-V-cycle (and the pre-smooth of level ls) in one thread block
    Level& L = slabs_[0].lv[ls];
    vcycle_small(ndim_, L.rhs, L.u, L.g, small_args_, d_info_, st_);
    cstart = ls;
  } else {
    solve_exact(ng - 1);
  }
  for (int c = cstart; c >= 1; --c) {  // coarse_to_fine :593-684
    if (!(ls > 0 && c == ls)) relax_sweeps(c, ms_);
    if (pr && c == 2) { prof_end(PROF_TAIL, st_); prof_begin(PROF_LEVEL1, st_); }
    if (pr && c == 1) prof_end(PROF_LEVEL1, st_);
    interp_add_from(c);
    relax_sweeps(c - 1, ms_);
  }
}

int MG::last_nexact() {
  int info[2];
  CUDA_CHECK(cudaMemcpyAsync(info, d_info_, sizeof info, cudaMemcpyDeviceToHost, st_));
  CUDA_CHECK(cudaStreamSynchronize(st_));
  return info[0];
}

// true when solve_exact() on the coarsest level will take the single-block shared-memory path (no host sync)
bool MG::coarsest_in_smem(const double* rhs_coarsest) const {
  const Grid& g = slabs_[0].lv.back().g;
  return rhs_coarsest != nullptr && (size_t)g.nx * g.ny * g.nz * 3 * sizeof(double) <= 200 * 1024 && g.nzl == g.nz;
}

// one iteration of solve_poisson_bvp's loop body, enqueue only: V-cycle, update_u, results to pinned memory
//
// Ping-pong (single slab, 3D): update_u (:1077-1122) copies the new iterate over the old one after measuring their
// difference -- 8 of its 24 B/point.  Instead, cycle n works in array B[n%2] (B0 = the hierarchy's level-0 array,
// B1 = the caller's) and reads the previous iterate from the other one: the first colour pass reads the other
// colour there and writes its own colour here, which touches exactly the data an in-place pass touches; the
// points no pass updates are carried over first (copy_fixed_points); update_u only measures the difference.
void MG::enqueue_cycle() {
  const bool pp = ss_.pingpong;
  double* cur = nullptr;
  if (pp) {
    Level& L0 = slabs_[0].lv[0];
    double* other = (ss_.it & 1) ? ss_.u[0] : u0_home_;
    cur = (ss_.it & 1) ? u0_home_ : ss_.u[0];
    L0.u = other;
    copy_fixed_points(cur, other, L0.g, L0.b, st_);
    pp_read_ = cur;
  }
  v_cycle();
  const bool dist = plan_.ndist > 0 && comm_;
  const bool prof = (ndim_ == 3 && slabs_.size() == 1);
  if (prof) prof_begin(PROF_DIFF0, st_);
  for (size_t s = 0; s < slabs_.size(); ++s) {
    Level& L0 = slabs_[s].lv[0];
    if (pp) diff_reduce(cur, L0.u, L0.g, false, scratch_, slabs_[s].d_out, st_);
    else diff_reduce(ss_.u[s], L0.u, L0.g, true, scratch_, slabs_[s].d_out, st_);  // update_u :122
  }
  if (prof) prof_end(PROF_DIFF0, st_);
  const int npairs = dist ? plan_.world : 1;
  if (dist)
    for (auto& S : slabs_) comm_->gather2(S.rank, S.d_out, d_all_, st_);
  const double* pairs = dist ? d_all_ : slabs_[0].d_out;
  if (h_out_dev_) {  // results straight into mapped pinned memory (no copy-engine node, see k_publish)
    publish_results(pairs, npairs, d_info_, h_out_dev_, st_);
  } else {
    CUDA_CHECK(cudaMemcpyAsync(h_out_, pairs, 2 * npairs * sizeof(double), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaMemcpyAsync(h_out_ + 2 * npairs, d_info_, 2 * sizeof(int), cudaMemcpyDeviceToHost, st_));
  }
}

// solve_poisson_bvp (ndsm_poisson.f90:63-155), split into begin / enqueue / poll / end so that several
// independent solves (the six chi faces) can be interleaved on their own streams by one host thread.
void MG::solve_begin(const std::vector<double*>& u, const std::vector<const double*>& rhs, double vc_tol, int nmax,
                     SolveTrace* tr) {
  if (u.size() != slabs_.size() || rhs.size() != slabs_.size()) throw NdsmError(6);
  ss_ = SolveState();
  ss_.u = u; ss_.vc_tol = vc_tol; ss_.nmax = nmax; ss_.tr = tr;
  if (slabs_.size() == 1) slabs_[0].lv[0].u = u0_home_;  // a ping-pong solve that was cut short may have left it redirected
  pp_read_ = nullptr;
  ss_.zero_rhs.assign(slabs_.size(), nullptr);
  for (size_t s = 0; s < slabs_.size(); ++s) {
    Level& L0 = slabs_[s].lv[0];
    const size_t bytes0 = (size_t)2 * L0.g.cs * sizeof(double);
    const double* r = rhs[s];
    if (!r && (ndim_ == 2 || ngrids() == 1)) {  // kernels of those paths always read rhs
      ss_.zero_rhs[s] = static_cast<double*>(pool_alloc(bytes0));
      CUDA_CHECK(cudaMemsetAsync(ss_.zero_rhs[s], 0, bytes0, st_));
      r = ss_.zero_rhs[s] + (i64)L0.H * L0.g.ps;
    }
    rhs0_[s] = r;
    CUDA_CHECK(cudaMemcpyAsync(level_base(L0.u, 0, (int)s), level_base(u[s], 0, (int)s), bytes0,
                               cudaMemcpyDeviceToDevice, st_));  // :100
  }
  for (auto& v : valid_) v = {{0, 0}};  // the caller's halo planes are not trusted; refreshed on first use
  for (auto& v : static_ok_) v = {{false, false}};
  // no rank may write into another rank's arena (replicated levels, gathered pairs) before that rank has
  // cleared / filled it: everything enqueued above is ordered before any peer's first message of this solve
  if (comm_ && plan_.ndist > 0) comm_->barrier(st_);

  // The loop body is a static launch sequence (the coarsest solve iterates inside one kernel), so it is
  // captured into a CUDA graph and replayed every V-cycle: ~250-500 launches per cycle otherwise.
  static const bool graphs_on = !(getenv("NDSM_B200_GRAPH") && atoi(getenv("NDSM_B200_GRAPH")) == 0);
  const double* rhs_coarsest = (ngrids() == 1) ? rhs0_[0] : slabs_[0].lv.back().rhs;
  ss_.use_graph = graphs_on && !prof_enabled() && nmax > 1 && (small_from_ > 0 || coarsest_in_smem(rhs_coarsest));
  const bool pp_env = !(getenv("NDSM_B200_PINGPONG") && atoi(getenv("NDSM_B200_PINGPONG")) == 0);
  ss_.pingpong = pp_env && ndim_ == 3 && slabs_.size() == 1 && !(plan_.ndist > 0 && comm_) && ngrids() >= 2 &&
                 ms_ >= 1 && nmax > 1;
  if (g_debug) debug_msg("solve_poisson_bvp", "Performing V cycles...");
  if (nmax <= 0) ss_.done = true;
}

void MG::solve_begin(double* u, const double* rhs, double vc_tol, int nmax, SolveTrace* tr) {
  solve_begin(std::vector<double*>{u}, std::vector<const double*>{rhs}, vc_tol, nmax, tr);
}

std::vector<unsigned long long> MG::graph_key(int parity) const {
  std::vector<unsigned long long> k;
  auto add = [&](const void* p) { k.push_back((unsigned long long)(size_t)p); };
  k.push_back((unsigned long long)parity | ((unsigned long long)ss_.pingpong << 1) | ((unsigned long long)ms_ << 8) |
              ((unsigned long long)nmax_exact_ << 24));
  unsigned long long bits;
  memcpy(&bits, &ex_tol_, sizeof bits);
  k.push_back(bits);
  unsigned long long c = 0;
  for (int d = 0; d < 2 * ndim_; ++d) c = c * 3 + (copt_[d] == 'D' ? 1 : 2);
  k.push_back(c | ((unsigned long long)du_max_ << 40) | ((unsigned long long)pdl_enabled() << 41));
  for (size_t s = 0; s < slabs_.size(); ++s) { add(ss_.u[s]); add(rhs0_[s]); add(slabs_[s].lv[0].u); }
  add(u0_home_);
  k.push_back(comm_ ? comm_->epoch() : 0ull);
  return k;
}

void MG::drop_graphs() {
  for (auto& g : gslot_) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    g.key.clear();
  }
}

void MG::capture_cycle(int parity) {
  GraphSlot& gs = gslot_[parity];
  if (gs.exec) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
  solve_exact_prepare();
  vcycle_small_prepare();
  // the halo state is whatever the previous cycle left when the graph is replayed; capture the pessimistic pattern
  for (auto& v : valid_) v = {{0, 0}};
  for (auto& v : static_ok_) v = {{false, false}};
  const unsigned long long l0 = g_launches, b0 = g_peer_bytes, m0 = g_peer_msgs;
  cudaGraph_t graph = nullptr;
  CUDA_CHECK(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
  try {
    enqueue_cycle();
  } catch (...) {
    cudaStreamEndCapture(st_, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  CUDA_CHECK(cudaStreamEndCapture(st_, &graph));
  gs.launches = g_launches - l0;
  gs.peer_bytes = g_peer_bytes - b0;
  gs.peer_msgs = g_peer_msgs - m0;
  g_launches = l0;  // nothing ran during capture
  g_peer_bytes = b0;
  g_peer_msgs = m0;
  cudaError_t e = cudaGraphInstantiate(&gs.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { gs.exec = nullptr; cudaGetLastError(); }
  gs.key = graph_key(parity);
}

void MG::solve_enqueue() {
  if (ss_.done) return;
  if (ss_.use_graph) {
    const int parity = ss_.pingpong ? (ss_.it & 1) : 0;
    GraphSlot& gs = gslot_[parity];
    if (ss_.pingpong) slabs_[0].lv[0].u = (ss_.it & 1) ? ss_.u[0] : u0_home_;  // the array this cycle works in (part of the key)
    if (!gs.exec || gs.key != graph_key(parity)) {
      static const bool say = getenv("NDSM_B200_TRACE") && atoi(getenv("NDSM_B200_TRACE")) >= 2;
      if (say) {
        const std::vector<unsigned long long> k = graph_key(parity);
        int diff = -1;
        for (size_t i = 0; i < k.size() && i < gs.key.size(); ++i)
          if (k[i] != gs.key[i]) { diff = (int)i; break; }
        fprintf(stderr, "TRACE graph capture: ndim %d parity %d had_exec %d first differing key word %d of %zu\n", ndim_, parity,
                gs.exec != nullptr, diff, k.size());
      }
      capture_cycle(parity);
    }
    if (gs.exec) {
      CUDA_CHECK(cudaGraphLaunch(gs.exec, st_));
      g_launches += gs.launches;
      g_peer_bytes += gs.peer_bytes;
      g_peer_msgs += gs.peer_msgs;
      return;
    }
    ss_.use_graph = false;  // instantiation failed: launch directly from here on
  }
  enqueue_cycle();
}

bool MG::solve_poll() {
  if (ss_.done) return true;
  const Grid& g0 = slabs_[0].lv[0].g;
  const double N = (double)((i64)g0.nx * g0.ny * g0.nz);
  CUDA_CHECK(cudaStreamSynchronize(st_));
  prof_collect();
  if (comm_ && comm_->failed()) {
    fprintf(stderr, "ERROR(solve_poisson_bvp):a peer did not answer within the time-out (NDSM_P2P_TIMEOUT_MS):NDSM_B200_ERR_INTERNAL\n");
    throw NdsmError(NDSM_ERR_INTERNAL);
  }
  const int npairs = (plan_.ndist > 0 && comm_) ? plan_.world : 1;
  double dmax = 0.0, dsum = 0.0;
  for (int q = 0; q < npairs; ++q) {  // fixed rank order: every rank takes the same decision
    dmax = h_out_[2 * q] > dmax ? h_out_[2 * q] : dmax;
    dsum += h_out_[2 * q + 1];
  }
  ss_.du = du_max_ ? dmax : dsum / N;
  const int* info = reinterpret_cast<const int*>(h_out_ + 2 * npairs);
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
  if (ss_.pingpong) {
    Level& L0 = slabs_[0].lv[0];
    L0.u = u0_home_;
    // after an odd number of cycles the last iterate sits in the hierarchy's array: hand it to the caller
    if (ss_.it & 1)
      CUDA_CHECK(cudaMemcpyAsync(ss_.u[0], u0_home_, (size_t)2 * L0.g.cs * sizeof(double), cudaMemcpyDeviceToDevice, st_));
    pp_read_ = nullptr;
  }
  if (du_last) *du_last = ss_.du;
  int ierr = 0;
  if (!ss_.converged) {
    ierr = 1;
    printf(" Warning: IOPT_NCYCLES exceeded. V-cycle iteration may not have converged\n");
  }
  if (ss_.tr) ss_.tr->ierr = ierr;
  bool freed = false;
  rhs0_halo_ok_ = false;
  for (size_t s = 0; s < slabs_.size(); ++s) {
    rhs0_[s] = nullptr;
    if (ss_.zero_rhs[s]) {
      if (!freed) { CUDA_CHECK(cudaStreamSynchronize(st_)); freed = true; }
      pool_free(ss_.zero_rhs[s]);
      ss_.zero_rhs[s] = nullptr;
    }
  }
  return ierr;
}

int MG::solve(const std::vector<double*>& u, const std::vector<const double*>& rhs, double vc_tol, int nmax,
              double* du_last, SolveTrace* tr) {
  solve_begin(u, rhs, vc_tol, nmax, tr);
  while (!ss_.done) {
    solve_enqueue();
    solve_poll();
  }
  return solve_end(du_last);
}

int MG::solve(double* u, const double* rhs, double vc_tol, int nmax, double* du_last, SolveTrace* tr) {
  return solve(std::vector<double*>{u}, std::vector<const double*>{rhs}, vc_tol, nmax, du_last, tr);
}

}  // namespace ndsm
