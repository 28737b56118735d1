// vecpot.cu -- compute_vector_potential on the GPU (ndsm_vector_potential.f90:130-497):
// face fluxes -> six 2D pure-Neumann chi solves -> At Dirichlet data -> three 3D Laplace
// solves -> flux-balance fields -> curl.  Everything between the face upload and the final
// dense A/B arrays stays resident in HBM.
#include "vecpot.hpp"
#include "pool.hpp"

#include <chrono>
#include <unistd.h>
#include <algorithm>
#include <cstring>
#include <string>
#include <cmath>
#include <memory>

namespace ndsm {

bool g_debug = false;
Report g_report;

void debug_msg(const char* sub, const char* msg) {  // ndsm_root.f90:493-503
  fprintf(stderr, "DEBUG(%s):%s\n", sub, msg);
}

static const int imap_cp[6] = {0, 0, 1, 1, 2, 2};                                 // :82
static const int imap_nc[6][2] = {{1, 2}, {1, 2}, {0, 2}, {0, 2}, {0, 1}, {0, 1}};  // :83

struct DevBuf {  // RAII device allocation
  double* p = nullptr;
  DevBuf() {}
  explicit DevBuf(size_t n) { alloc(n); }
  void alloc(size_t n) { p = static_cast<double*>(pool_alloc(n * sizeof(double))); }
  ~DevBuf() { if (p) pool_free(p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// NDSM_B200_TRACE=1: wall-clock checkpoints of the host orchestration on stderr (where does a stage's time go:
// hierarchy construction, graph capture/instantiation, the V-cycle loop)
struct HostTrace {
  bool on, verbose = false;  // NDSM_B200_TRACE=1: stages, =2: every chi V-cycle as well
  std::chrono::steady_clock::time_point t0, last;
  HostTrace() : on(getenv("NDSM_B200_TRACE") && atoi(getenv("NDSM_B200_TRACE")) != 0) {
    verbose = on && atoi(getenv("NDSM_B200_TRACE")) >= 2;
    t0 = last = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "TRACE[%d] %-28s +%8.3f ms  (total %8.3f ms)\n", (int)getpid(), what,
            std::chrono::duration<double, std::milli>(now - last).count(),
            std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
  }
};

struct EvTimer {
  cudaEvent_t a, b;
  cudaStream_t st;
  explicit EvTimer(cudaStream_t s) : st(s) { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EvTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, st); }
  double stop() {
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Solver cache.  A drop-in caller solves the same mesh again and again (a time series of magnetograms): the
// hierarchies (host tables, device arenas, symmetric-heap blocks, channels), their streams and their captured
// V-cycle graphs are kept between calls instead of being rebuilt -- 6 + 3 hierarchy constructions and up to 12
// graph captures per call otherwise (measured: 3.5 ms of a 16 ms BC setup, 2.5 ms of an 80 ms 8-GPU solve).
// Slots 0-5: chi faces, 6-8: the three components (6 alone when they are solved one after the other).
// A slot is rebuilt when anything it was built from differs: shape, mesh values, communicator, stream, device,
// or any NDSM_* environment variable.  NDSM_B200_CACHE=0, a workspace cap of 0 (the reference's per-call
// ownership) and ndsm_b200_release_workspace() drop it.
// ---------------------------------------------------------------------------------------------------------------
extern "C" char** environ;
namespace {
std::string env_signature() {
  std::vector<std::string> v;
  for (char** e = environ; e && *e; ++e)
    if (!strncmp(*e, "NDSM_", 5) && strncmp(*e, "NDSM_B200_TRACE", 15) && strncmp(*e, "NDSM_DEVICE", 11)) v.push_back(*e);
  std::sort(v.begin(), v.end());
  std::string s;
  for (auto& x : v) { s += x; s += ';'; }
  return s;
}
struct SolverSlot {
  int ndim = 0, shape[3] = {0, 0, 0}, device = -1;
  std::vector<double> mesh[3];
  Comm* base_comm = nullptr;
  bool cloned = false;
  cudaStream_t given_st = nullptr;
  std::string env;
  // resources
  cudaStream_t st = nullptr;
  bool own_stream = false;
  std::unique_ptr<Comm> own_comm;
  std::unique_ptr<MG> mg;
  bool used = false;
  // Work arrays that live as long as the hierarchy (the chi faces' iterate and right-hand side): their addresses
  // are baked into the captured V-cycle graph, so they must not change from call to call -- blocks taken from the
  // pool per call came back permuted every call and every chi graph was re-captured every call (0.6 ms each, and
  // every ~30 instantiations the driver took 25 ms for one).
  double* buf[2] = {nullptr, nullptr};
  size_t buf_n[2] = {0, 0};
  double* buffer(int i, size_t n) {
    if (buf_n[i] < n) {
      if (buf[i]) pool_free(buf[i]);
      buf[i] = static_cast<double*>(pool_alloc(n * sizeof(double)));
      buf_n[i] = n;
    }
    return buf[i];
  }
  void clear() {
    for (int i = 0; i < 2; ++i) {
      if (buf[i]) pool_free(buf[i]);
      buf[i] = nullptr;
      buf_n[i] = 0;
    }
    mg.reset();
    own_comm.reset();
    if (own_stream && st) cudaStreamDestroy(st);
    st = nullptr;
    own_stream = false;
    base_comm = nullptr;
    ndim = 0;
    used = false;
  }
};
// heap-allocated and never destroyed: at process exit the pool, the CUDA context and the communicators may already be
// gone when static destructors run (the cache is emptied explicitly by solver_cache_clear())
SolverSlot* const g_slots = new SolverSlot[9];
// the batched drivers over slots 6-8 (one per group of components): dropped before any of those hierarchies is
std::vector<std::unique_ptr<MGBatch>>* const g_batches = new std::vector<std::unique_ptr<MGBatch>>();

// hierarchy of slot `id` for this problem: the cached one when everything matches, a fresh one otherwise.
// given_st == nullptr: the slot owns its stream.  clone: the slot works on its own clone of base_comm.
SolverSlot& acquire_slot(int id, int ndim, const int* shape3, const double* const* mesh, Comm* base_comm, bool clone,
                         cudaStream_t given_st, cudaStream_t clone_st_hint) {
  SolverSlot& S = g_slots[id];
  int dev = 0;
  cudaGetDevice(&dev);
  const std::string env = env_signature();
  bool same = S.mg && S.ndim == ndim && S.device == dev && S.base_comm == base_comm && S.cloned == clone &&
              S.given_st == given_st && S.env == env;
  for (int d = 0; d < 3 && same; ++d) same = (S.shape[d] == shape3[d]);
  for (int d = 0; d < ndim && same; ++d)
    same = (S.mesh[d].size() == (size_t)shape3[d]) && !memcmp(S.mesh[d].data(), mesh[d], sizeof(double) * shape3[d]);
  if (!same) {
    if (id >= 6) {
      // component slots are acquired in ascending order and a later one may work on an earlier one's stream and
      // channel (a batched group): whatever comes after a rebuilt slot is rebuilt too
      g_batches->clear();
      for (int i = 8; i > id; --i) g_slots[i].clear();
    }
    S.clear();
    S.ndim = ndim;
    S.device = dev;
    S.base_comm = base_comm;
    S.cloned = clone;
    S.given_st = given_st;
    S.env = env;
    for (int d = 0; d < 3; ++d) S.shape[d] = shape3[d];
    for (int d = 0; d < ndim; ++d) S.mesh[d].assign(mesh[d], mesh[d] + shape3[d]);
    if (given_st) {
      S.st = given_st;
    } else {
      CUDA_CHECK(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
      S.own_stream = true;
    }
    Comm* c = base_comm;
    if (clone && base_comm) {
      S.own_comm = base_comm->clone(clone_st_hint ? clone_st_hint : S.st);
      c = S.own_comm.get();
    }
    S.mg.reset(new MG(ndim, shape3, -1, mesh, S.st, c));
  }
  S.used = true;
  return S;
}
bool cache_enabled(Comm* comm) {
  if (getenv("NDSM_B200_CACHE") && atoi(getenv("NDSM_B200_CACHE")) == 0) return false;
  if (getenv("NDSM_B200_WORKSPACE_CAP_MB") && strtoull(getenv("NDSM_B200_WORKSPACE_CAP_MB"), nullptr, 10) == 0 &&
      !(getenv("NDSM_B200_KEEP_WORKSPACE") && atoi(getenv("NDSM_B200_KEEP_WORKSPACE")) != 0))
    return false;
  return !comm || comm->persistent();
}
}  // namespace

// "012" | "01,2" | "0,12" | "02,1" | "0,1,2": every component exactly once; anything else gives no groups.  A group's
// first member (lowest component) owns the group's stream and channel: members and groups come back sorted.
std::vector<std::vector<int>> parse_component_groups(const char* spec_in) {
  std::vector<std::vector<int>> groups;
  const std::string spec = spec_in ? spec_in : "";
  int seen[3] = {0, 0, 0};
  bool ok = !spec.empty();
  std::vector<int> cur;
  for (size_t i = 0; ok && i <= spec.size(); ++i) {
    const char ch = i < spec.size() ? spec[i] : ',';
    if (ch == ',') {
      if (cur.empty()) { ok = false; break; }
      std::sort(cur.begin(), cur.end());
      groups.push_back(cur);
      cur.clear();
    } else if (ch >= '0' && ch <= '2' && !seen[ch - '0']) {
      seen[ch - '0'] = 1;
      cur.push_back(ch - '0');
    } else {
      ok = false;
    }
  }
  if (!ok || !(seen[0] && seen[1] && seen[2])) groups.clear();
  std::sort(groups.begin(), groups.end(), [](const std::vector<int>& x, const std::vector<int>& y) { return x[0] < y[0]; });
  return groups;
}

void solver_cache_clear() {
  g_batches->clear();
  for (int i = 8; i >= 0; --i) g_slots[i].clear();
}

int vector_solve_core(const int* nshape, const long long* iopt, const double* ropt, const double* x, const double* y,
                      const double* z, double* const* bn, const DenseIn& A0_in, Comm* comm, const std::vector<SlabOut>& outs_in,
                      cudaStream_t st, Report& rep, BcCapture* cap, bool stop_after_bc, const CoreHooks* hooks) {
  const int nx = nshape[0], ny = nshape[1], nz = nshape[2];
  const double* mesh[3] = {x, y, z};
  const bool use_du_max = (iopt[IOPT_DUMAX] == IOPT_TRUE);
  double Lq[3], dq[3];
  for (int d = 0; d < 3; ++d) {  // :201-221
    double lo = mesh[d][0], hi = mesh[d][0];
    for (int i = 1; i < nshape[d]; ++i) { lo = mesh[d][i] < lo ? mesh[d][i] : lo; hi = mesh[d][i] > hi ? mesh[d][i] : hi; }
    Lq[d] = hi - lo;
    dq[d] = mesh[d][1] - mesh[d][0];
  }
  const unsigned long long launches0 = g_launches;
  EvTimer tm(st), tall(st);
  tall.start();
  HostTrace trace;
  // hierarchies not worth (or not safe) keeping are dropped when this call returns
  struct CacheScope {
    bool keep;
    ~CacheScope() {
      if (!keep) solver_cache_clear();
      else for (int i = 0; i < 9; ++i) g_slots[i].used = false;
    }
  } cache_scope{cache_enabled(comm) && !cap && !g_debug};

  // mesh vectors on the device (flux-balance fields)
  DevBuf dmesh((size_t)nx + ny + nz);
  double* dx_ = dmesh.p;
  double* dy_ = dmesh.p + nx;
  double* dz_ = dmesh.p + nx + ny;
  CUDA_CHECK(cudaMemcpyAsync(dx_, x, sizeof(double) * nx, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(dy_, y, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(dz_, z, sizeof(double) * nz, cudaMemcpyHostToDevice, st));

  // ---------------- BC setup (K7) ----------------
  tm.start();
  int n1[6], n2[6];
  for (int f = 0; f < 6; ++f) { n1[f] = nshape[imap_nc[f][0]]; n2[f] = nshape[imap_nc[f][1]]; }
  DevBuf scr(reduce_scratch_doubles() + 16);
  double* d_phi = scr.p + reduce_scratch_doubles();
  for (int f = 0; f < 6; ++f)  // :300-306 -- always dq(1)*dq(2) (reference quirk)
    trapz_face(bn[f], n1[f], n2[f], dq[0], dq[1], scr.p, d_phi + f, st);
  double phi[6];
  CUDA_CHECK(cudaMemcpyAsync(phi, d_phi, sizeof phi, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  for (int f = 0; f < 6; ++f) rep.phi[f] = phi[f];
  trace.mark("face fluxes (trapz)");
  const double Aq[6] = {Lq[1] * Lq[2], Lq[1] * Lq[2], Lq[0] * Lq[2], Lq[0] * Lq[2], Lq[0] * Lq[1], Lq[0] * Lq[1]};

  // Multi-GPU: the six chi problems are independent (ndsm_vector_potential.f90:338-365), so face f is solved by
  // rank f mod world only and the resulting Dirichlet data At (two arrays per face) are broadcast to every rank
  // (peer-memory transport: written straight into the other ranks' copies).  One rank holding all slabs (virtual
  // ranks) and the BC-capture hook of the tests keep all six faces local.
  static const bool bc_dist_env = !(getenv("NDSM_BC_DISTRIBUTE") && atoi(getenv("NDSM_BC_DISTRIBUTE")) == 0);
  const bool dist_bc = comm && comm->world() > 1 && comm->nlocal() == 1 && !cap && bc_dist_env;
  const int bcW = dist_bc ? comm->world() : 1, bcme = dist_bc ? comm->first_rank() : 0;
  auto mine = [&](int f) { return !dist_bc || f % bcW == bcme; };
  // nothing of this call may reach a rank that is still working on the previous one
  if (comm && comm->nlocal() == 1) comm->barrier(st);
  struct BcBuf {  // dense face array that a broadcast may target
    Comm* c = nullptr;
    double* p = nullptr;
    void alloc(Comm* comm_, size_t n) {
      c = comm_;
      p = static_cast<double*>(c ? c->sym_alloc(n * sizeof(double)) : pool_alloc(n * sizeof(double)));
    }
    ~BcBuf() {
      if (!p) return;
      if (c) c->sym_free(p);
      else pool_free(p);
    }
  };
  BcBuf At[6][2], binfo;
  for (int f = 0; f < 6; ++f)
    for (int t = 0; t < 2; ++t) At[f][t].alloc(dist_bc ? comm : nullptr, (size_t)n1[f] * n2[f]);
  binfo.alloc(dist_bc ? comm : nullptr, 8);
  int ierr_last = 0;
  if (g_debug) debug_msg("compute_vector_potential", "Solve BVP on each boundary...");
  {
    // The chi problems are chains of tiny, latency-bound kernels, so the ones this rank owns run concurrently: one
    // hierarchy and one stream per face, V-cycles interleaved by this host thread.  (With the debug flag they
    // run one after the other so that the reference's message order is kept.)
    struct FaceSolve {
      MG* mg = nullptr;  // owned by the solver cache (slot f)
      struct { double* p = nullptr; } chi, rhs;  // owned by the solver slot
      cudaStream_t st = nullptr;
      int ierr = 0;
    } fs[6];
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (int f = 0; f < 6; ++f) {
      if (!mine(f)) continue;
      const int sh2[3] = {n1[f], n2[f], 1};
      const double* m2[3] = {mesh[imap_nc[f][0]], mesh[imap_nc[f][1]], nullptr};
      SolverSlot& slot = acquire_slot(f, 2, sh2, m2, nullptr, false, nullptr, nullptr);
      fs[f].mg = slot.mg.get();
      fs[f].st = slot.st;
      fs[f].mg->set_options((int)iopt[IOPT_MS], ropt[ROPT_CTOL], "NNNN", use_du_max, (int)iopt[IOPT_NMAXEX]);  // :355-357
      const Grid g2 = fs[f].mg->level(0).g;
      fs[f].chi.p = slot.buffer(0, 2 * (size_t)g2.cs);
      fs[f].rhs.p = slot.buffer(1, 2 * (size_t)g2.cs);
      CUDA_CHECK(cudaMemsetAsync(fs[f].rhs.p, 0, 2 * (size_t)g2.cs * sizeof(double), fs[f].st));
      CUDA_CHECK(cudaMemsetAsync(fs[f].chi.p, 0, 2 * (size_t)g2.cs * sizeof(double), fs[f].st));  // :345
      split_from_dense(bn[f], fs[f].rhs.p, g2, phi[f] / Aq[f], fs[f].st);                         // :348
    }
    trace.mark("2D hierarchy construction");
    auto run_face_to_end = [&](int f) {
      double du_last;
      fs[f].ierr = fs[f].mg->solve(fs[f].chi.p, fs[f].rhs.p, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &du_last, &rep.solves[f]);
    };
    if (g_debug) {
      for (int f = 0; f < 6; ++f)
        if (mine(f)) run_face_to_end(f);
    } else {
      for (int f = 0; f < 6; ++f)
        if (mine(f))
          fs[f].mg->solve_begin(fs[f].chi.p, fs[f].rhs.p, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &rep.solves[f]);
      trace.mark("chi solve_begin (graph capture)");
      for (int f = 0; f < 6; ++f)
        if (mine(f)) {
          fs[f].mg->solve_enqueue();
          if (trace.on && trace.verbose) trace.mark("  chi face first enqueue");
        }
      bool any = true;
      while (any) {  // a face is re-enqueued right after its own poll, so its stream never waits for the others
        any = false;
        for (int f = 0; f < 6; ++f) {
          if (!mine(f) || fs[f].mg->solve_done()) continue;
          const bool done = fs[f].mg->solve_poll();
          if (trace.on && trace.verbose) {
            char what[48];
            snprintf(what, sizeof what, "  chi face %d cycle polled", f);
            trace.mark(what);
          }
          if (!done) fs[f].mg->solve_enqueue();
          any = true;
        }
      }
      trace.mark("chi V-cycle loop");
      for (int f = 0; f < 6; ++f)
        if (mine(f)) { double du_last; fs[f].ierr = fs[f].mg->solve_end(&du_last); }
    }
    ierr_last = fs[5].ierr;  // :360,480 -- only the last chi solve's ierr survives (reference quirk)
    for (int f = 0; f < 6; ++f) {
      if (!mine(f)) continue;
      const Grid g2 = fs[f].mg->level(0).g;
      compute_At(fs[f].chi.p, g2, 1.0 / (2.0 * dq[imap_cp[f]]), f, At[f][0].p, At[f][1].p, fs[f].st);  // :394-398 (dq of normal dir)
      if (cap) {
        if (cap->chi[f]) {
          DevBuf dense((size_t)n1[f] * n2[f]);
          dense_from_split(fs[f].chi.p, dense.p, g2, fs[f].st);
          CUDA_CHECK(cudaMemcpyAsync(cap->chi[f], dense.p, sizeof(double) * n1[f] * n2[f], cudaMemcpyDeviceToHost, fs[f].st));
          CUDA_CHECK(cudaStreamSynchronize(fs[f].st));
        }
        if (cap->At1[f]) CUDA_CHECK(cudaMemcpyAsync(cap->At1[f], At[f][0].p, sizeof(double) * n1[f] * n2[f], cudaMemcpyDeviceToHost, fs[f].st));
        if (cap->At2[f]) CUDA_CHECK(cudaMemcpyAsync(cap->At2[f], At[f][1].p, sizeof(double) * n1[f] * n2[f], cudaMemcpyDeviceToHost, fs[f].st));
      }
      CUDA_CHECK(cudaStreamSynchronize(fs[f].st));
    }
    if (dist_bc) {  // all-gather of the At faces and of the chi solves' ierr
      double hinfo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int f = 0; f < 6; ++f)
        if (mine(f)) {
          hinfo[f] = (double)fs[f].ierr;
          CUDA_CHECK(cudaMemcpyAsync(binfo.p + f, &hinfo[f], sizeof(double), cudaMemcpyHostToDevice, st));
        }
      comm->begin(st);
      for (int f = 0; f < 6; ++f) {
        comm->bcast(f % bcW, At[f][0].p, (size_t)n1[f] * n2[f], st);
        comm->bcast(f % bcW, At[f][1].p, (size_t)n1[f] * n2[f], st);
        comm->bcast(f % bcW, binfo.p + f, 1, st);
      }
      comm->end(st);
      double hall[8];
      CUDA_CHECK(cudaMemcpyAsync(hall, binfo.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (comm->failed()) throw NdsmError(NDSM_ERR_INTERNAL);
      ierr_last = (int)hall[5];
    }
  }
  rep.ms_bc = tm.stop();
  trace.mark("At faces, BC setup done");
  if (stop_after_bc) {
    rep.launches = g_launches - launches0;
    return ierr_last;
  }

  // ---------------- three 3D solves (solve, :598-691) ----------------
  // initial guess of component c as received (the reference never zeroes A): the caller's dense array, or the
  // host entry's per-component hook
  auto guess_of = [&](int c) {
    if (hooks && hooks->guess) return hooks->guess(c);
    DenseIn g = A0_in;
    if (g.p) g.p += (i64)c * g.cstride;
    return g;
  };
  tm.start();
  if (g_debug) debug_msg("compute_vector_potential", "Solve BVP 3D...");
  const int sh3[3] = {nx, ny, nz};
  // Multi-GPU: the three component solves are independent (ndsm_vector_potential.f90:647-689), so every rank holds
  // a z-slab of all three and runs them concurrently on three streams, each with its own channel of the
  // peer-memory transport: while one component sits in its latency-bound tail (replicated coarse levels, halo
  // hand-shakes) the other two keep the HBM busy.  On one GPU the solves are bandwidth-bound and run one after
  // another on one hierarchy (a third of the memory, and the finished components overlap their D2H copies).
  // Over NCCL concurrently used communicators serialise badly (measured 10x slower), so there it is opt-in.
  const char* conc_env = getenv("NDSM_CONCURRENT_COMPONENTS");
  const bool conc_default = comm && comm->world() > 1 && comm->nlocal() == 1 && comm->one_sided();
  const bool conc_want = conc_env ? atoi(conc_env) != 0 : conc_default;
  const bool concurrent = conc_want && !prof_enabled() && !g_debug && (!comm || comm->nlocal() == 1);
  // ... or batched (mg_batch.cu): the components of a GROUP advance as one launch sequence -- a third of the launches
  // and hand-shakes, three times the work per launch, but nothing to overlap the latency-bound parts with; groups
  // run concurrently.  NDSM_COMPONENT_GROUPS names the groups ("012": one batch, "01,2": Ax+Ay batched next to
  // Az, "0,1,2" = three batches of one); NDSM_BATCH_COMPONENTS=1 is short for "012".  Also usable with virtual
  // slabs and on a single GPU (tests).
  std::vector<std::vector<int>> groups;
  {
    const char* ge = getenv("NDSM_COMPONENT_GROUPS");
    const char* be = getenv("NDSM_BATCH_COMPONENTS");
    groups = parse_component_groups(ge ? ge : ((be && atoi(be) != 0) ? "012" : ""));
    if (prof_enabled() || g_debug || (int)iopt[IOPT_NCYCLES] <= 1) groups.clear();
    if (groups.size() > 1 && comm && comm->nlocal() != 1) groups.clear();   // several channels: one rank per process
    if (groups.size() > 1 && comm && !comm->one_sided()) groups.clear();
  }
  const bool batch_want = !groups.empty();
  int group_of[3] = {0, 0, 0};
  for (size_t gi = 0; gi < groups.size(); ++gi)
    for (int c : groups[gi]) group_of[c] = (int)gi;
  struct Ctx {
    MG* mg = nullptr;  // owned by the solver cache (slots 6-8)
    cudaStream_t st = nullptr;
  } ctx[3];
  const int nctx = (concurrent || batch_want) ? 3 : 1;
  for (int q = 0; q < nctx; ++q) {
    // component 0 (and the sequential solves) work on the call's stream and communicator; concurrent
    // components 1 and 2 get a stream and a channel of their own; a batched group shares its first member's
    SolverSlot* slot;
    if (batch_want) {
      const std::vector<int>& G = groups[group_of[q]];
      if (G[0] == 0) {
        slot = &acquire_slot(6 + q, 3, sh3, mesh, comm, false, st, nullptr);
      } else if (G[0] == q) {
        slot = &acquire_slot(6 + q, 3, sh3, mesh, comm, comm != nullptr, nullptr, nullptr);
      } else {
        SolverSlot& first = g_slots[6 + G[0]];
        slot = &acquire_slot(6 + q, 3, sh3, mesh, first.own_comm ? first.own_comm.get() : comm, false, first.st, nullptr);
      }
    } else {
      slot = &acquire_slot(6 + q, 3, sh3, mesh, comm, q > 0 && comm != nullptr, q == 0 ? st : nullptr, nullptr);
    }
    ctx[q].mg = slot->mg.get();
    ctx[q].st = slot->st;
  }
  trace.mark("3D hierarchy construction");
  MG* mg3 = ctx[0].mg;
  auto mgc = [&](int c) { return ctx[nctx == 3 ? c : 0].mg; };
  const int ns = mg3->nslabs();
  std::vector<SlabOut> outs = outs_in;
  if (ns == 1 && outs.size() > 1) {  // every virtual rank shares one undivided solve: one contiguous output
    SlabOut o = outs.front();
    o.k1 = outs.back().k1;
    outs.assign(1, o);
  }
  if ((int)outs.size() != ns) throw NdsmError(6);
  rep.ndist = mg3->plan().ndist;
  rep.slab_points = 0;
  for (int s = 0; s < ns; ++s) rep.slab_points += (unsigned long long)nx * ny * mg3->level(0, s).g.nzl;
  std::vector<DevBuf> As(ns);
  std::vector<size_t> lvl(ns);
  std::vector<double*> Ap[3];  // per component: local plane 0 of every slab
  for (int s = 0; s < ns; ++s) {
    const Level& L0 = mg3->level(0, s);
    lvl[s] = mg3->level_doubles(0, s);
    As[s].alloc(3 * lvl[s]);
    CUDA_CHECK(cudaMemsetAsync(As[s].p, 0, 3 * lvl[s] * sizeof(double), st));
    for (int c = 0; c < 3; ++c) Ap[c].push_back(As[s].p + c * lvl[s] + (i64)L0.H * L0.g.ps);
  }
  static const int wf[3][4] = {{2, 3, 4, 5}, {0, 1, 4, 5}, {0, 1, 2, 3}};  // face write order :647-650,663-666,679-682
  static const int wa[3][4] = {{0, 0, 0, 0}, {0, 0, 1, 1}, {1, 1, 1, 1}};  // At(1,.) or At(2,.)
  static const char* cop[3] = {"NDDNDD", "DNDDND", "DDNDDN"};              // :655,671,687
  const std::vector<const double*> norhs(ns, nullptr);                     // rhs = 0 (:640-641)
  // component c's starting iterate: the guess, then the Dirichlet faces on top of it (:647-650,663-666,679-682)
  auto prepare = [&](int c, cudaStream_t sc) {
    const DenseIn A0 = guess_of(c);
    for (int s = 0; s < ns; ++s) {
      const Grid& g0 = mg3->level(0, s).g;
      if (A0.p) split_from_dense(A0.p + (i64)(g0.k0 - A0.kfirst) * nx * ny, Ap[c][s], g0, 0.0, sc);
      for (int w = 0; w < 4; ++w) {
        const int f = wf[c][w];
        const int layer = (f % 2 == 0) ? 0 : nshape[imap_cp[f]] - 1;
        write_face(Ap[c][s], g0, imap_cp[f], layer, At[f][wa[c][w]].p, sc);
      }
    }
  };
  CUDA_CHECK(cudaStreamSynchronize(st));  // the memsets above precede the component streams' work
  auto set_opts = [&](int c) {
    mgc(c)->set_options(c == 2 ? 5 : (int)iopt[IOPT_MS], ropt[ROPT_CTOL], cop[c], use_du_max, (int)iopt[IOPT_NMAXEX]);  // :685
  };
  // single slab writing the whole array in the default (flux first) order: a component of A is final as soon
  // as its solve has converged, so it is converted and handed out early
  const bool flux_first = (iopt[IOPT_FLXCRL] != 1);  // :453-477
  const bool early_out = nctx == 1 && ns == 1 && flux_first && outs[0].k0 == 0 && outs[0].k1 == nz &&
                         mg3->plan().ndist == 0;
  // ... and a component of B = curl A as soon as the two components of A it depends on are (Bz after Ay)
  const bool early_b = early_out && hooks && hooks->b_ready;
  bool batched = false;
  if (batch_want) {
    for (int c = 0; c < 3; ++c) set_opts(c);
    std::vector<std::vector<MG*>> mem(groups.size());
    batched = true;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      for (int c : groups[gi]) mem[gi].push_back(mgc(c));
      batched = batched && MGBatch::compatible(mem[gi]);
    }
    if (batched) {
      bool same = g_batches->size() == groups.size();
      for (size_t gi = 0; same && gi < groups.size(); ++gi) same = (*g_batches)[gi]->members() == mem[gi];
      if (!same) {
        g_batches->clear();
        for (size_t gi = 0; gi < groups.size(); ++gi) g_batches->emplace_back(new MGBatch(mem[gi]));
      }
      std::vector<std::vector<std::vector<double*>>> us(groups.size());
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        MGBatch& B = *(*g_batches)[gi];
        SolveTrace* trs[3] = {nullptr, nullptr, nullptr};
        for (size_t q = 0; q < groups[gi].size(); ++q) {
          const int c = groups[gi][q];
          prepare(c, B.stream());
          trs[q] = &rep.solves[6 + c];
          us[gi].push_back(Ap[c]);
        }
        B.solve_begin(us[gi], ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], trs);
      }
      // every group always has its next V-cycle queued (cf. the concurrent loop below)
      for (auto& B : *g_batches) B->solve_enqueue();
      bool any = true;
      while (any) {
        any = false;
        for (auto& B : *g_batches) {
          if (B->solve_done()) continue;
          if (!B->solve_poll()) B->solve_enqueue();
          any = true;
        }
      }
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        MGBatch& B = *(*g_batches)[gi];
        B.solve_end(nullptr, nullptr);
        B.exchange_level0(us[gi], 1);  // halo planes of the converged components (curl needs k-1, k+1)
        CUDA_CHECK(cudaStreamSynchronize(B.stream()));
      }
    }
  }
  rep.components_mode = batched ? 2 : ((concurrent && !batch_want) ? 1 : 0);
  if (batched) {
  } else if (!concurrent || batch_want) {  // (batch_want: the three hierarchies share one stream and one channel)
    for (int c = 0; c < 3; ++c) {
      set_opts(c);
      prepare(c, mgc(c)->stream());
      double du_last;
      mgc(c)->solve(Ap[c], norhs, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &du_last, &rep.solves[6 + c]);
      mgc(c)->exchange(0, 0, 3, 1, &Ap[c]);  // halo planes of the converged component (curl needs k-1, k+1)
      if (mgc(c)->stream() != st) CUDA_CHECK(cudaStreamSynchronize(mgc(c)->stream()));
      if (early_out) {
        unsplit_A(Ap[c][0], mg3->level(0, 0).g, c, dx_, dy_, dz_, phi, Lq, true, 0, nz, outs[0].A + c * outs[0].cstride, st);
        if (hooks && hooks->component_ready) hooks->component_ready(c);
        if (early_b && c == 1) {  // Bz = dAy/dx - dAx/dy (:804)
          curl_dense(outs[0].A, 0, outs[0].cstride, nx, ny, nz, dq[0], dq[1], dq[2], 0, nz, outs[0].B, outs[0].cstride, st, 2);
          hooks->b_ready(2);
        }
      }
    }
  } else {
    for (int c = 0; c < 3; ++c) {
      set_opts(c);
      prepare(c, ctx[c].st);
      mgc(c)->solve_begin(Ap[c], norhs, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &rep.solves[6 + c]);
    }
    // The three streams run the same program: started together they would sit in their latency-bound phases
    // (coarse levels, hand-shakes) at the same time and fight for HBM at the same time.  Component c starts
    // c * NDSM_STAGGER_US later, so that one stream's tail overlaps the others' smoothing.
    const double stagger_us = getenv("NDSM_STAGGER_US") ? atof(getenv("NDSM_STAGGER_US")) : 0.0;
    if (stagger_us > 0)
      for (int c = 1; c < 3; ++c) stream_delay(c * stagger_us, ctx[c].st);
    // every stream always has its next V-cycle queued: a component is re-enqueued right after its own poll
    for (int c = 0; c < 3; ++c) mgc(c)->solve_enqueue();
    bool any = true;
    while (any) {
      any = false;
      for (int c = 0; c < 3; ++c) {
        if (mgc(c)->solve_done()) continue;
        if (!mgc(c)->solve_poll()) mgc(c)->solve_enqueue();
        any = true;
      }
    }
    for (int c = 0; c < 3; ++c) {
      double du_last;
      mgc(c)->solve_end(&du_last);
      mgc(c)->exchange(0, 0, 3, 1, &Ap[c]);
      CUDA_CHECK(cudaStreamSynchronize(ctx[c].st));
    }
  }
  rep.ms_solve3d = tm.stop();
  trace.mark("three 3D solves");

  // ---------------- flux-balance fields + curl (K8) ----------------
  tm.start();
  if (g_debug) debug_msg("compute_vector_potential", "Compute B = curl(B) and flux correction...");
  if (!flux_first) printf(" FLAG SET: FLXCRL\n");
  const i64 pl = (i64)nx * ny;
  for (int s = 0; s < ns; ++s) {
    const Grid& g3 = mg3->level(0, s).g;
    // planes this slab has to deliver, and the planes of A the curl stencil touches
    int k0 = g3.k0, k1 = g3.k0 + g3.nzl;
    if (ns == 1 && mg3->plan().ndist == 0) { k0 = outs[0].k0; k1 = outs[0].k1; }  // replicated solve: own output range
    const SlabOut& o = outs[s];
    if (o.k0 != k0 || o.k1 != k1) throw NdsmError(6);
    const int glo = g3.k0 - mg3->level(0, s).H, ghi = g3.k0 + g3.nzl + mg3->level(0, s).H;
    int ka = (k0 == 0) ? 0 : k0 - 1, kb = (k1 == nz) ? nz : k1 + 1;
    if (k0 == 0 && kb < 3) kb = 3 < nz ? 3 : nz;       // one-sided stencil at the lower face needs planes 0,1,2
    if (k1 == nz && ka > nz - 3) ka = nz - 3 > 0 ? nz - 3 : 0;
    if (ka < (glo < 0 ? 0 : glo) || kb > (ghi > nz ? nz : ghi)) throw NdsmError(6);
    const bool in_place = (ka == k0 && kb == k1);
    DevBuf tmp;
    double* Ad = o.A;
    i64 csA = o.cstride;
    if (!in_place) {
      tmp.alloc((size_t)3 * (kb - ka) * pl);
      Ad = tmp.p;
      csA = (i64)(kb - ka) * pl;
    }
    if (!early_out)
      for (int c = 0; c < 3; ++c)
        unsplit_A(Ap[c][s], g3, c, dx_, dy_, dz_, phi, Lq, flux_first, ka, kb, Ad + c * csA, st);
    if (early_b) {  // Bz left after the Ay solve; Bx and By need Az
      for (int c = 0; c < 2; ++c) {
        curl_dense(Ad, ka, csA, nx, ny, nz, dq[0], dq[1], dq[2], k0, k1, o.B, o.cstride, st, c);
        hooks->b_ready(c);
      }
    } else {
      curl_dense(Ad, ka, csA, nx, ny, nz, dq[0], dq[1], dq[2], k0, k1, o.B, o.cstride, st);
    }
    if (!in_place)
      for (int c = 0; c < 3; ++c)
        CUDA_CHECK(cudaMemcpyAsync(o.A + c * o.cstride, Ad + c * csA + (i64)(k0 - ka) * pl, (size_t)(k1 - k0) * pl * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
    if (!flux_first) add_flux_dense(o.A, o.cstride, o.B, o.cstride, nx, ny, k0, k1, dx_, dy_, dz_, phi, Lq, st);
    CUDA_CHECK(cudaStreamSynchronize(st));  // tmp goes out of scope
  }
  rep.ms_post = tm.stop();
  trace.mark("flux fields + curl");
  rep.ms_device = tall.stop();
  rep.launches = g_launches - launches0;
  return ierr_last;  // :480 -- ierr of the LAST chi solve (reference quirk)
}

void output_range(int nz, int world, int rank, int* k0, int* k1) {
  *k0 = (int)((i64)nz * rank / world);
  *k1 = (int)((i64)nz * (rank + 1) / world);
}

}  // namespace ndsm
