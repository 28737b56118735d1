// vecpot.cu -- compute_vector_potential on the GPU (ndsm_vector_potential.f90:130-497):
// face fluxes -> six 2D pure-Neumann chi solves -> At Dirichlet data -> three 3D Laplace
// solves -> flux-balance fields -> curl.  Everything between the face upload and the final
// dense A/B arrays stays resident in HBM.
#include "vecpot.hpp"
#include "pool.hpp"

#include <chrono>
#include <algorithm>
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

int vector_solve_core(const int* nshape, const long long* iopt, const double* ropt, const double* x, const double* y,
                      const double* z, double* const* bn, const DenseIn& A0_in, Comm* comm, const std::vector<SlabOut>& outs_in,
                      cudaStream_t st, Report& rep, BcCapture* cap, bool stop_after_bc, const CoreHooks* hooks,
                      const Hybrid* hyb) {
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
  const double Aq[6] = {Lq[1] * Lq[2], Lq[1] * Lq[2], Lq[0] * Lq[2], Lq[0] * Lq[2], Lq[0] * Lq[1], Lq[0] * Lq[1]};

  DevBuf At[6][2];
  int ierr_last = 0;
  if (g_debug) debug_msg("compute_vector_potential", "Solve BVP on each boundary...");
  {
    // The six chi problems are independent (ndsm_vector_potential.f90:338-365) and each is a chain of tiny,
    // latency-bound kernels, so they run concurrently: one hierarchy and one stream per face, V-cycles
    // interleaved by this host thread.  (With the debug flag they run one after the other so that the
    // reference's message order is kept.)
    struct FaceSolve {
      std::unique_ptr<MG> mg;
      DevBuf chi, rhs;
      cudaStream_t st = nullptr;
      int ierr = 0;
    } fs[6];
    struct StreamGuard {
      FaceSolve* f;
      ~StreamGuard() { for (int i = 0; i < 6; ++i) if (f[i].st) cudaStreamDestroy(f[i].st); }
    } guard{fs};
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (int f = 0; f < 6; ++f) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&fs[f].st, cudaStreamNonBlocking));
      const int sh2[3] = {n1[f], n2[f], 1};
      const double* m2[2] = {mesh[imap_nc[f][0]], mesh[imap_nc[f][1]]};
      fs[f].mg.reset(new MG(2, sh2, -1, m2, fs[f].st));
      fs[f].mg->set_options((int)iopt[IOPT_MS], ropt[ROPT_CTOL], "NNNN", use_du_max, (int)iopt[IOPT_NMAXEX]);  // :355-357
      const Grid g2 = fs[f].mg->level(0).g;
      fs[f].chi.alloc(2 * (size_t)g2.cs);
      fs[f].rhs.alloc(2 * (size_t)g2.cs);
      CUDA_CHECK(cudaMemsetAsync(fs[f].rhs.p, 0, 2 * (size_t)g2.cs * sizeof(double), fs[f].st));
      CUDA_CHECK(cudaMemsetAsync(fs[f].chi.p, 0, 2 * (size_t)g2.cs * sizeof(double), fs[f].st));  // :345
      split_from_dense(bn[f], fs[f].rhs.p, g2, phi[f] / Aq[f], fs[f].st);                         // :348
      At[f][0].alloc((size_t)n1[f] * n2[f]);
      At[f][1].alloc((size_t)n1[f] * n2[f]);
    }
    auto run_face_to_end = [&](int f) {
      double du_last;
      fs[f].ierr = fs[f].mg->solve(fs[f].chi.p, fs[f].rhs.p, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &du_last, &rep.solves[f]);
    };
    if (g_debug) {
      for (int f = 0; f < 6; ++f) run_face_to_end(f);
    } else {
      for (int f = 0; f < 6; ++f)
        fs[f].mg->solve_begin(fs[f].chi.p, fs[f].rhs.p, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &rep.solves[f]);
      bool any = true;
      while (any) {
        any = false;
        for (int f = 0; f < 6; ++f)
          if (!fs[f].mg->solve_done()) fs[f].mg->solve_enqueue();
        for (int f = 0; f < 6; ++f)
          if (!fs[f].mg->solve_done()) { fs[f].mg->solve_poll(); any = true; }
      }
      for (int f = 0; f < 6; ++f) { double du_last; fs[f].ierr = fs[f].mg->solve_end(&du_last); }
    }
    ierr_last = fs[5].ierr;  // :360,480 -- only the last chi solve's ierr survives (reference quirk)
    for (int f = 0; f < 6; ++f) {
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
  }
  rep.ms_bc = tm.stop();
  if (stop_after_bc) {
    rep.launches = g_launches - launches0;
    return ierr_last;
  }

  // ---------------- hybrid: my group solves ONE component, then all-to-all of dense A planes ------------
  if (hyb) {
    tm.start();
    const int c = hyb->comp, me = hyb->world->first_rank(), W = hyb->world->world();
    const int sh3h[3] = {nx, ny, nz};
    const i64 plh = (i64)nx * ny;
    static const int wf[3][4] = {{2, 3, 4, 5}, {0, 1, 4, 5}, {0, 1, 2, 3}};
    static const int wa[3][4] = {{0, 0, 0, 0}, {0, 0, 1, 1}, {1, 1, 1, 1}};
    static const char* cop[3] = {"NDDNDD", "DNDDND", "DDNDDN"};
    const bool flux_first = (iopt[IOPT_FLXCRL] != 1);
    if (!flux_first && me == 0) printf(" FLAG SET: FLXCRL\n");
    int a0, a1;  // planes of component c this rank converts and sends out
    output_range(nz, hyb->gsize[c], me - hyb->gfirst[c], &a0, &a1);
    DevBuf mine((size_t)std::max(a1 - a0, 1) * plh);
    {
      MG mg(3, sh3h, -1, mesh, st, hyb->gsize[c] > 1 ? hyb->group : nullptr);
      rep.ndist = mg.plan().ndist;
      const Level& L0 = mg.level(0, 0);
      rep.slab_points = (unsigned long long)nx * ny * L0.g.nzl;
      const size_t lv0 = mg.level_doubles(0, 0);
      DevBuf As1(lv0);
      CUDA_CHECK(cudaMemsetAsync(As1.p, 0, lv0 * sizeof(double), st));
      double* p0 = As1.p + (i64)L0.H * L0.g.ps;
      for (int w = 0; w < 4; ++w) {
        const int f = wf[c][w];
        write_face(p0, L0.g, imap_cp[f], (f % 2 == 0) ? 0 : nshape[imap_cp[f]] - 1, At[f][wa[c][w]].p, st);
      }
      mg.set_options(c == 2 ? 5 : (int)iopt[IOPT_MS], ropt[ROPT_CTOL], cop[c], use_du_max, (int)iopt[IOPT_NMAXEX]);
      double du_last;
      mg.solve(p0, nullptr, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &du_last, &rep.solves[6 + c]);
      if (L0.g.k0 > a0 || L0.g.k0 + L0.g.nzl < a1) throw NdsmError(6);
      unsplit_A(p0, L0.g, c, dx_, dy_, dz_, phi, Lq, flux_first, a0, a1, mine.p, st);
      CUDA_CHECK(cudaStreamSynchronize(st));
    }
    rep.ms_solve3d = tm.stop();
    tm.start();
    // planes every rank needs: its output range plus the planes the curl stencil touches
    auto need = [&](int d, int* ka, int* kb, int* k0, int* k1) {
      output_range(nz, W, d, k0, k1);
      *ka = (*k0 == 0) ? 0 : *k0 - 1;
      *kb = (*k1 == nz) ? nz : *k1 + 1;
      if (*k0 == 0 && *kb < 3) *kb = 3 < nz ? 3 : nz;
      if (*k1 == nz && *ka > nz - 3) *ka = nz - 3 > 0 ? nz - 3 : 0;
    };
    if (outs_in.size() != 1) throw NdsmError(6);
    const SlabOut& o = outs_in[0];
    int ka, kb, k0, k1;
    need(me, &ka, &kb, &k0, &k1);
    if (o.k0 != k0 || o.k1 != k1) throw NdsmError(6);
    const i64 csA = (i64)(kb - ka) * plh;
    DevBuf tmp((size_t)3 * csA);
    hyb->world->begin(st);
    for (int cc = 0; cc < 3; ++cc)
      for (int q = hyb->gfirst[cc]; q < hyb->gfirst[cc] + hyb->gsize[cc]; ++q) {
        int qa0, qa1;
        output_range(nz, hyb->gsize[cc], q - hyb->gfirst[cc], &qa0, &qa1);
        for (int d = 0; d < W; ++d) {
          int dka, dkb, dk0, dk1;
          need(d, &dka, &dkb, &dk0, &dk1);
          const int v0 = std::max(qa0, dka), v1 = std::min(qa1, dkb);
          if (v1 <= v0) continue;
          const size_t n = (size_t)(v1 - v0) * plh;
          if (q == me && d == me) {
            CUDA_CHECK(cudaMemcpyAsync(tmp.p + cc * csA + (i64)(v0 - ka) * plh, mine.p + (i64)(v0 - a0) * plh,
                                       n * sizeof(double), cudaMemcpyDeviceToDevice, st));
          } else if (q == me) {
            hyb->world->send(me, d, mine.p + (i64)(v0 - a0) * plh, n, st);
          } else if (d == me) {
            hyb->world->recv(me, q, tmp.p + cc * csA + (i64)(v0 - ka) * plh, n, st);
          }
        }
      }
    hyb->world->end(st);
    curl_dense(tmp.p, ka, csA, nx, ny, nz, dq[0], dq[1], dq[2], k0, k1, o.B, o.cstride, st);
    for (int cc = 0; cc < 3; ++cc)
      CUDA_CHECK(cudaMemcpyAsync(o.A + cc * o.cstride, tmp.p + cc * csA + (i64)(k0 - ka) * plh,
                                 (size_t)(k1 - k0) * plh * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (!flux_first) add_flux_dense(o.A, o.cstride, o.B, o.cstride, nx, ny, k0, k1, dx_, dy_, dz_, phi, Lq, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    rep.ms_post = tm.stop();
    rep.ms_device = tall.stop();
    rep.launches = g_launches - launches0;
    return ierr_last;
  }

  // ---------------- three 3D solves (solve, :598-691) ----------------
  const DenseIn A0 = (hooks && hooks->guess) ? hooks->guess() : A0_in;
  tm.start();
  if (g_debug) debug_msg("compute_vector_potential", "Solve BVP 3D...");
  const int sh3[3] = {nx, ny, nz};
  // Multi-GPU: the three component solves are independent (ndsm_vector_potential.f90:647-689), so they run
  // concurrently on three streams with three communicators: the halo-exchange latency of one solve is hidden
  // behind the kernels of the other two.  On one GPU the solves are bandwidth-bound and run one after another
  // on one hierarchy (a third of the memory).
  // opt-in: fine with virtual ranks, but NCCL serialises kernels of concurrently used communicators badly
  // (measured 10x slower on 2 GPUs), so the default is one solve at a time
  static const bool conc_env = getenv("NDSM_CONCURRENT_COMPONENTS") && atoi(getenv("NDSM_CONCURRENT_COMPONENTS")) != 0;
  const bool concurrent = comm && comm->world() > 1 && conc_env && !prof_enabled() && !g_debug;
  struct Ctx {
    std::unique_ptr<MG> mg;
    std::unique_ptr<Comm> own_comm;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    ~Ctx() {
      mg.reset();
      own_comm.reset();
      if (own_stream && st) cudaStreamDestroy(st);
    }
  } ctx[3];
  const int nctx = concurrent ? 3 : 1;
  for (int q = 0; q < nctx; ++q) {
    Comm* cq = comm;
    ctx[q].st = st;
    if (q > 0) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&ctx[q].st, cudaStreamNonBlocking));
      ctx[q].own_stream = true;
      ctx[q].own_comm = comm->clone(st);
      cq = ctx[q].own_comm.get();
    }
    ctx[q].mg.reset(new MG(3, sh3, -1, mesh, ctx[q].st, cq));
  }
  MG* mg3 = ctx[0].mg.get();
  auto mgc = [&](int c) { return ctx[concurrent ? c : 0].mg.get(); };
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
    for (int c = 0; c < 3; ++c) {
      double* p0 = As[s].p + c * lvl[s] + (i64)L0.H * L0.g.ps;
      Ap[c].push_back(p0);
      if (A0.p)  // initial guess as received (reference never zeroes A)
        split_from_dense(A0.p + c * A0.cstride + (i64)(L0.g.k0 - A0.kfirst) * nx * ny, p0, L0.g, 0.0, st);
    }
  }
  static const int wf[3][4] = {{2, 3, 4, 5}, {0, 1, 4, 5}, {0, 1, 2, 3}};  // face write order :647-650,663-666,679-682
  static const int wa[3][4] = {{0, 0, 0, 0}, {0, 0, 1, 1}, {1, 1, 1, 1}};  // At(1,.) or At(2,.)
  static const char* cop[3] = {"NDDNDD", "DNDDND", "DDNDDN"};              // :655,671,687
  const std::vector<const double*> norhs(ns, nullptr);                     // rhs = 0 (:640-641)
  for (int c = 0; c < 3; ++c)
    for (int s = 0; s < ns; ++s)
      for (int w = 0; w < 4; ++w) {
        const int f = wf[c][w];
        const int layer = (f % 2 == 0) ? 0 : nshape[imap_cp[f]] - 1;
        write_face(Ap[c][s], mg3->level(0, s).g, imap_cp[f], layer, At[f][wa[c][w]].p, st);
      }
  CUDA_CHECK(cudaStreamSynchronize(st));
  auto set_opts = [&](int c) {
    mgc(c)->set_options(c == 2 ? 5 : (int)iopt[IOPT_MS], ropt[ROPT_CTOL], cop[c], use_du_max, (int)iopt[IOPT_NMAXEX]);  // :685
  };
  // single slab writing the whole array in the default (flux first) order: a component of A is final as soon
  // as its solve has converged, so it is converted and handed out early
  const bool flux_first = (iopt[IOPT_FLXCRL] != 1);  // :453-477
  const bool early_out = !concurrent && ns == 1 && flux_first && outs[0].k0 == 0 && outs[0].k1 == nz &&
                         mg3->plan().ndist == 0;
  if (!concurrent) {
    for (int c = 0; c < 3; ++c) {
      set_opts(c);
      double du_last;
      mg3->solve(Ap[c], norhs, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &du_last, &rep.solves[6 + c]);
      mg3->exchange(0, 0, 3, 1, &Ap[c]);  // halo planes of the converged component (curl needs k-1, k+1)
      if (early_out) {
        unsplit_A(Ap[c][0], mg3->level(0, 0).g, c, dx_, dy_, dz_, phi, Lq, true, 0, nz, outs[0].A + c * outs[0].cstride, st);
        if (hooks && hooks->component_ready) hooks->component_ready(c);
      }
    }
  } else {
    for (int c = 0; c < 3; ++c) {
      set_opts(c);
      mgc(c)->solve_begin(Ap[c], norhs, ropt[ROPT_VTOL], (int)iopt[IOPT_NCYCLES], &rep.solves[6 + c]);
    }
    bool any = true;
    while (any) {
      any = false;
      for (int c = 0; c < 3; ++c)
        if (!mgc(c)->solve_done()) mgc(c)->solve_enqueue();
      for (int c = 0; c < 3; ++c)
        if (!mgc(c)->solve_done()) { mgc(c)->solve_poll(); any = true; }
    }
    for (int c = 0; c < 3; ++c) {
      double du_last;
      mgc(c)->solve_end(&du_last);
      mgc(c)->exchange(0, 0, 3, 1, &Ap[c]);
      CUDA_CHECK(cudaStreamSynchronize(ctx[c].st));
    }
  }
  rep.ms_solve3d = tm.stop();

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
    curl_dense(Ad, ka, csA, nx, ny, nz, dq[0], dq[1], dq[2], k0, k1, o.B, o.cstride, st);
    if (!in_place)
      for (int c = 0; c < 3; ++c)
        CUDA_CHECK(cudaMemcpyAsync(o.A + c * o.cstride, Ad + c * csA + (i64)(k0 - ka) * pl, (size_t)(k1 - k0) * pl * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
    if (!flux_first) add_flux_dense(o.A, o.cstride, o.B, o.cstride, nx, ny, k0, k1, dx_, dy_, dz_, phi, Lq, st);
    CUDA_CHECK(cudaStreamSynchronize(st));  // tmp goes out of scope
  }
  rep.ms_post = tm.stop();
  rep.ms_device = tall.stop();
  rep.launches = g_launches - launches0;
  return ierr_last;  // :480 -- ierr of the LAST chi solve (reference quirk)
}

// contiguous groups of ranks, sizes as equal as possible (larger groups first)
void hybrid_groups(int world, int* gfirst3, int* gsize3) {
  int at = 0;
  for (int c = 0; c < 3; ++c) {
    gsize3[c] = world / 3 + (c < world % 3 ? 1 : 0);
    gfirst3[c] = at;
    at += gsize3[c];
  }
}

void output_range(int nz, int world, int rank, int* k0, int* k1) {
  *k0 = (int)((i64)nz * rank / world);
  *k1 = (int)((i64)nz * (rank + 1) / world);
}

}  // namespace ndsm
