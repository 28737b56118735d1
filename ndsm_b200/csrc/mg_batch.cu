// Batched V-cycles: several independent Poisson solves on the SAME grids (the components Ax, Ay, Az of the vector
// potential, ndsm_vector_potential.f90:647-689) advance as one launch sequence on one stream.
//
// Why: on a z-slab of a multi-GPU solve every kernel is short (a 513^3 grid on 8 GPUs leaves 64 planes per rank;
// a colour pass takes ~30 us of which ~10 us are launch ramp and tail) and every halo exchange is a latency-bound
// hand-shake.  Three concurrent streams hide part of that; merging the three components into one launch and one
// hand-shake removes it: a third of the launches, exchanges and flag round trips, three times the work per launch.
//
// Every member keeps its own hierarchy (arrays, Dirichlet pattern, first colour, halo validity); the batch only
// drives them in lock-step.  The arithmetic of every member is exactly that of MG::v_cycle (same kernels' device
// code, same summation orders), so the results are bit-identical to the member-by-member solves.  A member that
// has converged drops out of the batch (its V-cycle count is its own, as in the reference's three sequential
// solve_poisson_bvp calls); the remaining ones continue with a re-captured graph.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mg.hpp"
#include "pool.hpp"

namespace ndsm {

extern unsigned long long g_launches;

bool MGBatch::compatible(const std::vector<MG*>& m) {
  if (m.empty() || m.size() > NDSM_BATCH_MAX) return false;
  const MG* a = m[0];
  if (a->ndim_ != 3 || a->ngrids() < 2 || a->ms_ < 1) return false;
  for (const MG* b : m) {
    if (b->ndim_ != 3 || b->ngrids() != a->ngrids() || b->slabs_.size() != a->slabs_.size()) return false;
    if (b->ms_ != a->ms_ || b->ex_tol_ != a->ex_tol_ || b->du_max_ != a->du_max_ || b->nmax_exact_ != a->nmax_exact_) return false;
    if (b->all_neumann_ || b->small_from_ != a->small_from_ || a->small_from_ <= 0) return false;
    if (b->plan_.world != a->plan_.world || b->plan_.ndist != a->plan_.ndist || b->plan_.halo != a->plan_.halo) return false;
    if (b->plan_.zs != a->plan_.zs) return false;
    for (int g = 0; g < a->ngrids(); ++g) {
      const Grid& x = a->slabs_[0].lv[g].g;
      const Grid& y = b->slabs_[0].lv[g].g;
      if (x.nx != y.nx || x.ny != y.ny || x.nz != y.nz || x.cs != y.cs || x.k0 != y.k0 || x.nzl != y.nzl) return false;
    }
    if (b != a && (b->slabs_[0].arena == a->slabs_[0].arena)) return false;
  }
  return true;
}

MGBatch::MGBatch(const std::vector<MG*>& members) : m_(members) {
  if (!compatible(m_)) throw NdsmError(NDSM_ERR_INTERNAL);
  lead_ = m_[0];
  st_ = lead_->st_;
  comm_ = lead_->comm_;
  const MG& L = *lead_;
  if (comm_ && L.plan_.ndist > 0)  // one staged message carries both colours of a full halo of every member
    comm_->reserve((size_t)m_.size() * 2 * L.plan_.halo * (size_t)L.slabs_[0].lv[0].g.ps);
  const int world = L.plan_.world;
  const size_t nd = (size_t)m_.size() * 2 * world + 8;
  d_all_ = static_cast<double*>(comm_ ? comm_->sym_alloc(nd * sizeof(double)) : pool_alloc(nd * sizeof(double)));
  CUDA_CHECK(cudaMemsetAsync(d_all_, 0, nd * sizeof(double), st_));
  d_small_ = static_cast<SmallArgs*>(pool_alloc(m_.size() * sizeof(SmallArgs)));
  h_out_ = static_cast<double*>(pool_alloc_host((nd + 8) * sizeof(double)));
  void* dp = nullptr;
  if (2 * m_.size() * world <= 120 && cudaHostGetDevicePointer(&dp, h_out_, 0) == cudaSuccess) h_out_dev_ = static_cast<double*>(dp);
  else cudaGetLastError();
}

MGBatch::~MGBatch() {
  drop_graphs();
  cudaStreamSynchronize(st_);
  if (d_all_) { if (comm_) comm_->sym_free(d_all_); else pool_free(d_all_); }
  if (d_small_) pool_free(d_small_);
  if (h_out_) pool_free_host(h_out_);
}

void MGBatch::drop_graphs() {
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
}

std::vector<MG*> MGBatch::active() const {
  std::vector<MG*> v;
  for (size_t i = 0; i < m_.size(); ++i)
    if (on_[i]) v.push_back(m_[i]);
  return v;
}

// one grouped exchange of halo planes for several members (cf. MG::exchange)
void MGBatch::exchange(int g, int which, const std::vector<Item>& items, int np) {
  const MG& L = *lead_;
  if (g >= L.plan_.ndist || !comm_ || items.empty()) return;
  const int world = L.plan_.world;
  comm_->begin(st_);
  for (size_t s = 0; s < L.slabs_.size(); ++s) {
    for (const Item& it : items) {
      Slab& S = it.m->slabs_[s];
      const Level& V = S.lv[g];
      double* P = it.arr ? (*it.arr)[s] : (which == 0 ? V.u : it.m->r_scratch(g, (int)s));
      const size_t n = (size_t)np * V.g.ps;
      for (int c = 0; c < 2; ++c) {
        if (!(it.mask & (1 << c))) continue;
        double* pc = P + (i64)c * V.g.cs;
        if (S.rank + 1 < world) {
          comm_->send(S.rank, S.rank + 1, pc + (i64)(V.g.nzl - np) * V.g.ps, n, st_);
          comm_->recv(S.rank, S.rank + 1, pc + (i64)V.g.nzl * V.g.ps, n, st_);
        }
        if (S.rank > 0) {
          comm_->send(S.rank, S.rank - 1, pc, n, st_);
          comm_->recv(S.rank, S.rank - 1, pc - (i64)np * V.g.ps, n, st_);
        }
      }
    }
  }
  comm_->end(st_);
}

void MGBatch::need_halo(int g, int depth) {
  const MG& L = *lead_;
  if (g >= L.plan_.ndist || !comm_) return;
  std::vector<Item> items;
  for (MG* m : active())
    if (m->valid_[g][0] < depth || m->valid_[g][1] < depth) items.push_back(Item{m, 3, nullptr});
  if (items.empty()) return;
  exchange(g, 0, items, L.plan_.halo);
  for (const Item& it : items) {
    it.m->valid_[g][0] = it.m->valid_[g][1] = L.plan_.halo;
    it.m->static_ok_[g][0] = it.m->static_ok_[g][1] = true;
  }
}

void MGBatch::relax(int g) {  // cf. MG::relax (3D, not pure Neumann)
  const MG& L = *lead_;
  const bool dist = g < L.plan_.ndist && comm_;
  const size_t ns = dist ? L.slabs_.size() : 1;
  const std::vector<MG*> act = active();
  static const bool one_colour = !(getenv("NDSM_HALO_ONE_COLOUR") && atoi(getenv("NDSM_HALO_ONE_COLOUR")) == 0);
  for (int pass = 0; pass < 2; ++pass) {
    int ext = 0;
    if (dist) {
      std::vector<Item> items;
      for (MG* m : act) {
        const int need = 1 - (m->first_colour_ ^ pass);  // the colour this pass reads
        if (m->valid_[g][need] >= 1) continue;
        const bool both = !one_colour || !m->static_ok_[g][1 - need];  // see MG::need_halo_colour
        items.push_back(Item{m, both ? 3 : (1 << need), nullptr});
      }
      exchange(g, 0, items, L.plan_.halo);
      for (const Item& it : items)
        for (int c = 0; c < 2; ++c)
          if (it.mask & (1 << c)) { it.m->valid_[g][c] = L.plan_.halo; it.m->static_ok_[g][c] = true; }
      ext = L.plan_.halo;
      for (MG* m : act) ext = std::min(ext, m->valid_[g][1 - (m->first_colour_ ^ pass)] - 1);
    }
    for (size_t s = 0; s < ns; ++s) {
      RelaxBatch bt;
      memset(&bt, 0, sizeof bt);
      bt.n = (int)act.size();
      for (int q = 0; q < bt.n; ++q) {
        Level& V = act[q]->slabs_[s].lv[g];
        bt.u[q] = V.u;
        bt.rhs[q] = (g == 0) ? nullptr : V.rhs;
        bt.b[q] = V.b;
        bt.colour[q] = act[q]->first_colour_ ^ pass;
      }
      const Level& V0 = act[0]->slabs_[s].lv[g];
      relax3d_half_batch(bt, V0.g, V0.w, ext, st_);
    }
    if (dist)
      for (MG* m : act) m->valid_[g][m->first_colour_ ^ pass] = ext;
  }
}

void MGBatch::residual(int g) {  // cf. MG::residual
  need_halo(g, 1);
  const MG& L = *lead_;
  const size_t ns = (g < L.plan_.ndist) ? L.slabs_.size() : 1;
  const std::vector<MG*> act = active();
  for (size_t s = 0; s < ns; ++s) {
    ResidualBatch bt;
    memset(&bt, 0, sizeof bt);
    bt.n = (int)act.size();
    for (int q = 0; q < bt.n; ++q) {
      Level& V = act[q]->slabs_[s].lv[g];
      bt.u[q] = V.u;
      bt.rhs[q] = (g == 0) ? nullptr : V.rhs;
      bt.r[q] = act[q]->r_scratch(g, (int)s);
      bt.b[q] = V.b;
    }
    const Level& V0 = act[0]->slabs_[s].lv[g];
    residual3d_batch(bt, V0.g, V0.w, st_);
  }
}

void MGBatch::restrict_to(int g) {  // cf. MG::restrict_to + MG::finish_restrict
  const MG& L = *lead_;
  const int c = g + 1;
  const bool fdist = g < L.plan_.ndist, cdist = c < L.plan_.ndist;
  const size_t ns = fdist ? L.slabs_.size() : 1;
  const std::vector<MG*> act = active();
  if (fdist) {
    std::vector<Item> items;
    for (MG* m : act) items.push_back(Item{m, 3, nullptr});
    exchange(g, 2, items, L.rneed_[g]);
  }
  for (size_t s = 0; s < ns; ++s) {
    const Level& F = L.slabs_[s].lv[g];
    const Level& C = L.slabs_[s].lv[c];
    Grid gv = C.g;
    i64 shift = 0;
    if (fdist && !cdist) {  // this rank produces planes [zs[c][r], zs[c][r+1]) of the replicated array
      const int r = L.slabs_[s].rank;
      gv.k0 = L.plan_.zs[c][r];
      gv.nzl = L.plan_.zs[c][r + 1] - gv.k0;
      shift = (i64)gv.k0 * C.g.ps;
    }
    if (gv.nzl <= 0) continue;
    if (F.rdirect) {
      TransferBatch bt;
      memset(&bt, 0, sizeof bt);
      bt.n = (int)act.size();
      for (int q = 0; q < bt.n; ++q) {
        bt.src[q] = act[q]->r_scratch(g, (int)s);
        bt.dst[q] = act[q]->slabs_[s].lv[c].rhs + shift;
      }
      restrict_direct_batch(bt, F.g, gv, F.rt[0], F.rt[1], F.rt[2], st_);
      continue;
    }
    for (MG* m : act) {  // other operators: member by member (same choice as MG::restrict_to)
      const double* rf = m->r_scratch(g, (int)s);
      double* out = m->slabs_[s].lv[c].rhs + shift;
      if (F.rsep) restrict_sep(rf, F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], st_);
      else if (F.fused) restrict_tiled(rf, F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], F.rr_hwp, F.rr_fyw, st_);
      else restrict_level(rf, F.g, out, gv, F.rt[0], F.rt[1], F.rt[2], st_);
    }
  }
  if (fdist && !cdist && comm_) {  // all-gather the replicated rhs of every member in one group
    comm_->begin(st_);
    for (MG* m : act) {
      Level& C = m->slabs_[0].lv[c];
      for (int q = 0; q < L.plan_.world; ++q) {
        const int k0 = L.plan_.zs[c][q], cnt = L.plan_.zs[c][q + 1] - k0;
        if (cnt <= 0) continue;
        for (int col = 0; col < 2; ++col)
          comm_->bcast(q, C.rhs + (i64)col * C.g.cs + (i64)k0 * C.g.ps, (size_t)cnt * C.g.ps, st_);
      }
    }
    comm_->end(st_);
  }
  if (cdist) {  // extended colour passes read rhs in the halo planes
    std::vector<std::vector<double*>> rp(act.size());
    std::vector<Item> items;
    for (size_t q = 0; q < act.size(); ++q) {
      for (auto& S : act[q]->slabs_) rp[q].push_back(S.lv[c].rhs);
      items.push_back(Item{act[q], 3, &rp[q]});
    }
    exchange(c, 0, items, L.plan_.halo);
    for (MG* m : act) {
      m->valid_[c][0] = m->valid_[c][1] = L.plan_.halo;  // u[c] = 0 everywhere, halos included
      m->static_ok_[c][0] = m->static_ok_[c][1] = true;
    }
  }
  const size_t nc = cdist ? L.slabs_.size() : 1;
  for (MG* m : act)
    for (size_t s = 0; s < nc; ++s) {  // ndsm_multigrid_core.f90:557-558
      Level& C = m->slabs_[s].lv[c];
      CUDA_CHECK(cudaMemsetAsync(m->level_base(C.u, c, (int)s), 0, (size_t)2 * C.g.cs * sizeof(double), st_));
    }
}

void MGBatch::interp_add_from(int c) {  // cf. MG::interp_add_from
  const MG& L = *lead_;
  const int f = c - 1;
  const bool fdist = f < L.plan_.ndist, cdist = c < L.plan_.ndist;
  const size_t ns = fdist ? L.slabs_.size() : 1;
  const std::vector<MG*> act = active();
  if (cdist) {
    std::vector<Item> items;
    for (MG* m : act)
      if (m->valid_[c][0] < L.ineed_[c] || m->valid_[c][1] < L.ineed_[c]) items.push_back(Item{m, 3, nullptr});
    exchange(c, 0, items, L.ineed_[c]);
    for (const Item& it : items) {
      it.m->valid_[c][0] = std::max(it.m->valid_[c][0], L.ineed_[c]);
      it.m->valid_[c][1] = std::max(it.m->valid_[c][1], L.ineed_[c]);
    }
  }
  for (size_t s = 0; s < ns; ++s) {
    const Level& C = L.slabs_[s].lv[c];
    const Level& F = L.slabs_[s].lv[f];
    if (F.icols) {
      TransferBatch bt;
      memset(&bt, 0, sizeof bt);
      bt.n = (int)act.size();
      for (int q = 0; q < bt.n; ++q) {
        bt.src[q] = act[q]->slabs_[s].lv[c].u;
        bt.dst[q] = act[q]->slabs_[s].lv[f].u;
      }
      interp_add_zt_batch(bt, C.g, F.g, F.it[0], F.it[1], F.it[2], st_);
      continue;
    }
    for (MG* m : act) {
      const double* uc = m->slabs_[s].lv[c].u;
      double* uf = m->slabs_[s].lv[f].u;
      if (F.itiled) interp_add_tiled(uc, C.g, uf, F.g, F.it[0], F.it[1], F.it[2], st_);
      else interp_add(uc, C.g, uf, F.g, F.it[0], F.it[1], F.it[2], st_);
    }
  }
  if (fdist)
    for (MG* m : act) {
      m->valid_[f][0] = m->valid_[f][1] = 0;
      m->static_ok_[f][0] = m->static_ok_[f][1] = false;
    }
}

void MGBatch::v_cycle() {  // cf. MG::v_cycle with small_from_ > 0
  const MG& L = *lead_;
  const int ls = L.small_from_, ms = L.ms_;
  for (int g = 0; g < ls; ++g) {
    for (int s = 0; s < ms; ++s) relax(g);
    residual(g);
    restrict_to(g);
  }
  {
    SmallBatch sb;
    memset(&sb, 0, sizeof sb);
    int n = 0;
    for (size_t i = 0; i < m_.size(); ++i) {
      if (!on_[i]) continue;
      Level& V = m_[i]->slabs_[0].lv[ls];
      sb.rhs_in[n] = V.rhs;
      sb.u_out[n] = V.u;
      sb.info[n] = m_[i]->d_info_;
      sb.slot[n] = (int)i;
      ++n;
    }
    sb.n = n;
    vcycle_small_batch(sb, L.slabs_[0].lv[ls].g, L.small_args_, d_small_, st_);
  }
  for (int c = ls; c >= 1; --c) {
    if (c != ls)
      for (int s = 0; s < ms; ++s) relax(c);
    interp_add_from(c);
    for (int s = 0; s < ms; ++s) relax(c - 1);
  }
}

// V-cycle + update_u of every active member + results to pinned memory (cf. MG::enqueue_cycle)
void MGBatch::enqueue_cycle() {
  const MG& L = *lead_;
  v_cycle();
  const std::vector<MG*> act = active();
  const int na = (int)act.size();
  const bool dist = L.plan_.ndist > 0 && comm_;
  const int world = dist ? L.plan_.world : 1;
  // member q's (max,sum) pair of slab s lands in the slab's send buffer at [2q, 2q+1]
  for (size_t s = 0; s < L.slabs_.size(); ++s)
    for (int q = 0; q < na; ++q) {
      MG* m = act[q];
      Level& V = m->slabs_[s].lv[0];
      diff_reduce(m->ss_.u[s], V.u, V.g, true, m->scratch_, lead_->slabs_[s].d_out + 2 * q, st_);  // update_u :122
    }
  if (dist)
    for (auto& S : lead_->slabs_) comm_->gathern(S.rank, S.d_out, 2 * na, d_all_, st_);
  const double* pairs = dist ? d_all_ : lead_->slabs_[0].d_out;
  const int* infos[NDSM_BATCH_MAX] = {nullptr, nullptr, nullptr};
  for (int q = 0; q < na; ++q) infos[q] = act[q]->d_info_;
  if (h_out_dev_) {
    publish_results_batch(pairs, na * world, infos, na, h_out_dev_, st_);
  } else {
    CUDA_CHECK(cudaMemcpyAsync(h_out_, pairs, 2 * (size_t)na * world * sizeof(double), cudaMemcpyDeviceToHost, st_));
    for (int q = 0; q < na; ++q)
      CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<int*>(h_out_ + 2 * na * world) + 2 * q, infos[q], 2 * sizeof(int),
                                 cudaMemcpyDeviceToHost, st_));
  }
}

unsigned MGBatch::mask() const {
  unsigned k = 0;
  for (size_t i = 0; i < m_.size(); ++i) k |= (on_[i] ? 1u : 0u) << i;
  return k;
}

std::vector<unsigned long long> MGBatch::graph_key() const {
  std::vector<unsigned long long> k;
  k.push_back(mask());
  for (size_t i = 0; i < m_.size(); ++i) {
    if (!on_[i]) continue;
    const std::vector<unsigned long long> km = m_[i]->graph_key(0);
    k.insert(k.end(), km.begin(), km.end());
    k.push_back((unsigned long long)(size_t)m_[i]->d_info_);
  }
  k.push_back(comm_ ? comm_->epoch() : 0ull);
  k.push_back((unsigned long long)(size_t)d_small_);
  k.push_back((unsigned long long)(size_t)d_all_);
  k.push_back((unsigned long long)(size_t)h_out_dev_);
  return k;
}

void MGBatch::capture() {
  Slot& gs = graphs_[mask()];
  if (gs.exec) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
  vcycle_small_prepare();
  // replayed with whatever halo state the previous cycle left: capture the pessimistic pattern
  for (MG* m : active()) {
    for (auto& v : m->valid_) v = {{0, 0}};
    for (auto& v : m->static_ok_) v = {{false, false}};
  }
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
  g_launches = l0;
  g_peer_bytes = b0;
  g_peer_msgs = m0;
  cudaError_t e = cudaGraphInstantiate(&gs.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { gs.exec = nullptr; cudaGetLastError(); }
  gs.key = graph_key();
}

// The members' solve_poisson_bvp loops (ndsm_poisson.f90:63-155) in lock-step, as a state machine like MG's (begin /
// enqueue / poll / end) so that several groups can be interleaved by one host thread.  u[i]: member i's iterate per
// slab (level-0 layout, local plane 0), in/out; rhs == 0 for every member.
void MGBatch::solve_begin(const std::vector<std::vector<double*>>& u, double vc_tol, int nmax, SolveTrace* const* tr) {
  const MG& L = *lead_;
  const size_t nm = m_.size();
  if (u.size() != nm) throw NdsmError(NDSM_ERR_INTERNAL);
  if (!compatible(m_)) throw NdsmError(NDSM_ERR_INTERNAL);  // options may have changed since construction
  on_.assign(nm, nmax > 0);
  du_.assign(nm, 1.7976931348623157e308);
  its_.assign(nm, 0);
  conv_.assign(nm, 0);
  tr_.assign(nm, nullptr);
  vc_tol_ = vc_tol;
  nmax_ = nmax;
  for (size_t i = 0; i < nm; ++i) {
    MG* m = m_[i];
    if (tr) tr_[i] = tr[i];
    if (u[i].size() != m->slabs_.size()) throw NdsmError(NDSM_ERR_INTERNAL);
    m->ss_ = MG::SolveState();
    m->ss_.u = u[i];
    m->ss_.zero_rhs.assign(m->slabs_.size(), nullptr);
    if (m->slabs_.size() == 1) m->slabs_[0].lv[0].u = m->u0_home_;
    m->pp_read_ = nullptr;
    for (size_t s = 0; s < m->slabs_.size(); ++s) {
      Level& V = m->slabs_[s].lv[0];
      m->rhs0_[s] = nullptr;
      CUDA_CHECK(cudaMemcpyAsync(m->level_base(V.u, 0, (int)s), m->level_base(u[i][s], 0, (int)s),
                                 (size_t)2 * V.g.cs * sizeof(double), cudaMemcpyDeviceToDevice, st_));  // :100
    }
    for (auto& v : m->valid_) v = {{0, 0}};
    for (auto& v : m->static_ok_) v = {{false, false}};
    // the small-level argument block of this member (boundary pattern, first colour) for the batched kernel;
    // uploaded only when it changed (the source is pageable memory: the copy is synchronous then)
    if (small_host_.size() != nm) small_host_.resize(nm);
    if (memcmp(&small_host_[i], &m->small_args_, sizeof(SmallArgs)) != 0) {
      memcpy(&small_host_[i], &m->small_args_, sizeof(SmallArgs));
      CUDA_CHECK(cudaMemcpyAsync(d_small_ + i, &small_host_[i], sizeof(SmallArgs), cudaMemcpyHostToDevice, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
    }
  }
  if (comm_ && L.plan_.ndist > 0) comm_->barrier(st_);  // see MG::solve_begin
  static const bool graphs_on = !(getenv("NDSM_B200_GRAPH") && atoi(getenv("NDSM_B200_GRAPH")) == 0);
  use_graph_ = graphs_on && !prof_enabled() && nmax > 1;
}

void MGBatch::solve_enqueue() {
  if (mask() == 0) return;
  if (use_graph_) {
    {
      Slot& gs = graphs_[mask()];
      if (!gs.exec || gs.key != graph_key()) capture();
    }
    Slot& gs = graphs_[mask()];
    if (gs.exec) {
      CUDA_CHECK(cudaGraphLaunch(gs.exec, st_));
      g_launches += gs.launches;
      g_peer_bytes += gs.peer_bytes;
      g_peer_msgs += gs.peer_msgs;
      return;
    }
    use_graph_ = false;  // instantiation failed: launch directly from here on
  }
  enqueue_cycle();
}

bool MGBatch::solve_poll() {
  if (mask() == 0) return true;
  const MG& L = *lead_;
  const size_t nm = m_.size();
  const Grid& g0 = L.slabs_[0].lv[0].g;
  const double N = (double)((i64)g0.nx * g0.ny * g0.nz);
  const bool dist = L.plan_.ndist > 0 && comm_;
  const int world = dist ? L.plan_.world : 1;
  CUDA_CHECK(cudaStreamSynchronize(st_));
  prof_collect();
  if (comm_ && comm_->failed()) {
    fprintf(stderr, "ERROR(solve_poisson_bvp):a peer did not answer within the time-out (NDSM_P2P_TIMEOUT_MS):NDSM_B200_ERR_INTERNAL\n");
    throw NdsmError(NDSM_ERR_INTERNAL);
  }
  // gathered layout: [rank][active member][max, sum], then two ints per active member
  std::vector<size_t> act;
  for (size_t i = 0; i < nm; ++i)
    if (on_[i]) act.push_back(i);
  const int na = (int)act.size();
  const int* info = reinterpret_cast<const int*>(h_out_ + 2 * na * world);
  for (int q = 0; q < na; ++q) {
    const size_t i = act[q];
    double dmax = 0.0, dsum = 0.0;
    for (int r = 0; r < world; ++r) {  // fixed rank order: every rank takes the same decision
      const double* p = h_out_ + ((size_t)r * na + q) * 2;
      dmax = p[0] > dmax ? p[0] : dmax;
      dsum += p[1];
    }
    du_[i] = L.du_max_ ? dmax : dsum / N;
    if (tr_[i]) { tr_[i]->du.push_back(du_[i]); tr_[i]->nexact.push_back(info[2 * q]); }
    if (!info[2 * q + 1]) printf(" Warning: IOPT_NMAXEX exceeded. Coarse-mesh solution may not have converged\n");
    ++its_[i];
    if (du_[i] < vc_tol_) { conv_[i] = 1; on_[i] = false; }  // :136 strict <
    else if (its_[i] >= nmax_) on_[i] = false;
  }
  return mask() == 0;
}

void MGBatch::solve_end(double* du_last, int* ierr) {
  for (size_t i = 0; i < m_.size(); ++i) {
    MG* m = m_[i];
    m->ss_.it = its_[i];
    m->ss_.du = du_[i];
    m->ss_.converged = conv_[i] != 0;
    m->ss_.done = true;
    if (du_last) du_last[i] = du_[i];
    const int e = conv_[i] ? 0 : 1;
    if (e) printf(" Warning: IOPT_NCYCLES exceeded. V-cycle iteration may not have converged\n");
    if (ierr) ierr[i] = e;
    if (tr_[i]) tr_[i]->ierr = e;
  }
}

void MGBatch::solve(const std::vector<std::vector<double*>>& u, double vc_tol, int nmax, SolveTrace* const* tr,
                    double* du_last, int* ierr) {
  solve_begin(u, vc_tol, nmax, tr);
  while (!solve_done()) {
    solve_enqueue();
    solve_poll();
  }
  solve_end(du_last, ierr);
}

// halo planes of the converged components (the curl stencil reads k-1, k+1), one exchange for all of them
void MGBatch::exchange_level0(const std::vector<std::vector<double*>>& u, int np) {
  std::vector<Item> items;
  for (size_t i = 0; i < m_.size(); ++i) items.push_back(Item{m_[i], 3, &u[i]});
  exchange(0, 0, items, np);
}

}  // namespace ndsm
