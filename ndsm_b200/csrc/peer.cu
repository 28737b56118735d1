// peer.cu -- peer-memory transport for the z-slab decomposition: one process per GPU, data written straight
// into the other ranks' HBM over NVLink 5 / NVSwitch, handed over with flags that live in peer memory.
//
// Why not NCCL for the data path: a halo exchange is 6-13 MB to each z-neighbour, 13-20 times per V-cycle.  NVLink
// moves that in ~10 us; one grouped ncclSend/ncclRecv costs 60-160 us of launch and proxy latency (measured in
// round 1), which made the exchanges 18 % of the 8-GPU step.  Here an exchange is two small kernels on the
// solve's own stream (so both are captured into the V-cycle's CUDA graph):
//   k_push        copies this rank's boundary planes into the neighbour's inbox (remote stores through a CUDA IPC
//                 mapping), then the last block publishes a sequence number in the neighbour's flag word
//                 (__threadfence_system + st.release.sys);
//   k_wait_unpack spins on the local flag word (ld.acquire.sys, with a time-out instead of a hang), then copies
//                 the inbox into the halo planes.
// Inboxes are double-buffered by the parity of a per-direction message counter kept in device memory, so a
// replayed graph needs no host-side state; a neighbour can never be two staged messages ahead because every
// staged send is paired with a staged receive from the same rank (see end()).  Replicated coarse levels and
// the (max,sum) pairs of update_u are written directly into the peers' copies of the same buffer: those
// buffers come from a SYMMETRIC HEAP (same offset on every rank), and the per-cycle all-gather of the pairs
// orders the cycles globally, which makes the direct writes race-free.
//
// NCCL is only the bootstrap: it carries the 64-byte IPC handles when a heap segment is created.
#include <cstring>
#include <iterator>
#include <map>
#include <vector>

#include "mg.hpp"
#include "pool.hpp"
#include "sym_alloc.hpp"

namespace ndsm {

extern unsigned long long g_launches;

namespace {

typedef unsigned long long u64;
constexpr int MAX_WORLD = 16;
constexpr int MAX_CH = 8;    // communicators alive at the same time (one per concurrent solve)
constexpr int MAX_SEG = 64;  // copy segments per launch (three batched components: 3 x 2 colours x 7 peers = 42)
constexpr size_t SEG_MIN = (size_t)256 << 20;

// ---------------------------------------------------------------------------------------------
// symmetric heap: segments of cudaMalloc memory, every rank maps every other rank's copy (CUDA IPC)
// ---------------------------------------------------------------------------------------------
// Allocation inside a segment: sym_alloc.hpp (first fit, coalescing; the same call sequence on every rank gives the
// same offsets on every rank).
struct Segment : SegmentAllocator {
  char* base[MAX_WORLD];
  size_t bytes = 0;
};

// control block at the start of segment 0 (u64 words, zero-initialised)
enum { CW_FLAGS = 0, CW_SENT = 1, CW_SENT_ST = 2, CW_EXPECT = 3, CW_EXPECT_ST = 4, CW_NARR = 5 };
constexpr size_t CTRL_WORDS = (size_t)CW_NARR * MAX_CH * MAX_WORLD + 2 * MAX_CH + 8;

struct Fabric {
  Comm* boot = nullptr;
  int rank = 0, world = 1;
  std::vector<Segment> segs;
  u64* ctrl = nullptr;
  int* h_err = nullptr;  // pinned + mapped: a device-side wait timed out
  int* d_err_map = nullptr;
  bool ch_used[MAX_CH];
  bool ok = false;
  u64 timeout_ns = 30ull * 1000000000ull;

  u64* word(int arr, int ch, int r) const { return ctrl + ((size_t)arr * MAX_CH + ch) * MAX_WORLD + r; }
  unsigned* ticket(int ch, int which) const {
    return reinterpret_cast<unsigned*>(ctrl + (size_t)CW_NARR * MAX_CH * MAX_WORLD + 2 * ch + which);
  }
  int* d_err() const { return reinterpret_cast<int*>(ctrl + (size_t)CW_NARR * MAX_CH * MAX_WORLD + 2 * MAX_CH); }

  // address of my buffer `p` in rank r's copy of the heap
  template <typename T>
  T* peer(const T* p, int r) const {
    const char* c = reinterpret_cast<const char*>(p);
    for (const Segment& s : segs)
      if (c >= s.base[rank] && c < s.base[rank] + s.bytes)
        return reinterpret_cast<T*>(s.base[r] + (c - s.base[rank]));
    fprintf(stderr, "ERROR(peer):buffer is not in the symmetric heap:NDSM_B200_ERR_INTERNAL\n");
    throw NdsmError(NDSM_ERR_INTERNAL);
  }

  void new_segment(size_t bytes, cudaStream_t st) {
    Segment s;
    memset(s.base, 0, sizeof s.base);
    s.bytes = bytes;
    char* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {  // the pool's cache may hold the memory
      cudaGetLastError();
      pool_trim(0);
      CUDA_CHECK(cudaMalloc(&p, bytes));
    }
    CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
    std::vector<cudaIpcMemHandle_t> all(world);
    boot->allgather_host(&h, all.data(), sizeof h, st);  // synchronises st: the memset is complete
    s.base[rank] = p;
    int good = 1;
    for (int r = 0; r < world && good; ++r) {
      if (r == rank) continue;
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        good = 0;
      }
      s.base[r] = static_cast<char*>(q);
    }
    // every rank must take the same decision
    std::vector<int> goods(world, 0);
    boot->allgather_host(&good, goods.data(), sizeof(int), st);
    for (int v : goods) good = good && v;
    if (!good) {
      for (int r = 0; r < world; ++r)
        if (r != rank && s.base[r]) cudaIpcCloseMemHandle(s.base[r]);
      cudaFree(p);
      fprintf(stderr, "ERROR(peer):cudaIpcOpenMemHandle failed on some rank:NDSM_B200_ERR_CUDA\n");
      throw NdsmError(NDSM_ERR_CUDA);
    }
    s.reset(bytes);
    segs.push_back(s);
  }

  void* alloc(size_t bytes, cudaStream_t st) {
    const size_t b = (bytes + 511) / 512 * 512;
    for (Segment& s : segs) {
      const size_t off = s.take(b);
      if (off != (size_t)-1) return s.base[rank] + off;
    }
    new_segment(b > SEG_MIN ? b : SEG_MIN, st);
    return segs.back().base[rank] + segs.back().take(b);
  }
  void free(void* p) {
    const char* c = static_cast<const char*>(p);
    for (Segment& s : segs)
      if (c >= s.base[rank] && c < s.base[rank] + s.bytes) {
        s.give((size_t)(c - s.base[rank]));
        return;
      }
  }
  static size_t ctrl_bytes() { return (CTRL_WORDS * sizeof(u64) + 511) / 512 * 512; }

  void shutdown() {
    if (!ok) return;
    cudaDeviceSynchronize();
    for (Segment& s : segs) {
      for (int r = 0; r < world; ++r)
        if (r != rank && s.base[r]) cudaIpcCloseMemHandle(s.base[r]);
    }
    // peers may still be unmapping: the allocations themselves are released after a last handshake
    if (boot) {
      int one = 1;
      std::vector<int> all(world);
      cudaStream_t st = nullptr;
      if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess) {
        try { boot->allgather_host(&one, all.data(), sizeof(int), st); } catch (...) {}
        cudaStreamDestroy(st);
      }
    }
    for (Segment& s : segs) cudaFree(s.base[rank]);
    segs.clear();
    if (h_err) cudaFreeHost(h_err);
    h_err = nullptr;
    ctrl = nullptr;
    ok = false;
  }
};
// heap-allocated and never destroyed: communicators held in other translation units' statics may outlive it otherwise
Fabric* const g_fab_ptr = new Fabric();
#define g_fab (*g_fab_ptr)

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
struct CopySeg {
  const double* src;
  double* dst;
  u64 n;              // doubles
  u64 parity_stride;  // doubles added to the inbox side (dst of a push, src of an unpack) for odd messages
  int peer;           // index into the launch's peer list (-1: local copy, no parity)
};
struct PushArgs {
  int nseg, npeer;
  CopySeg seg[MAX_SEG];
  u64* peer_flag[MAX_WORLD];  // the peer's flag word for (channel, me) -- remote address
  u64* sent[MAX_WORLD];       // my counter of signals to that peer
  u64* sent_st[MAX_WORLD];    // my counter of STAGED messages to that peer (inbox parity)
  unsigned staged_mask;       // peers that receive a staged message in this launch
  unsigned* ticket;
};
struct WaitArgs {
  int nseg, npeer;
  CopySeg seg[MAX_SEG];
  const u64* flag[MAX_WORLD];  // my flag word for (channel, peer)
  u64* expect[MAX_WORLD];
  u64* expect_st[MAX_WORLD];
  unsigned staged_mask;
  unsigned* ticket;
  int* d_err;      // device: set once a wait timed out (later waits return at once)
  int* h_err_map;  // the same for the host (mapped pinned memory)
  u64 timeout_ns;
};

__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(u64* p, u64 v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_timer_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// grid-strided copy of one segment, 16-byte accesses when both sides allow it (plane strides are 256-byte
// multiples; face arrays and the 2-double reduction pairs take the scalar tail / path)
__device__ __forceinline__ void copy_segment(const double* __restrict__ src, double* __restrict__ dst, const u64 n) {
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nth = (u64)gridDim.x * blockDim.x;
  if ((((u64)src | (u64)dst) & 15ull) == 0) {
    const u64 n2 = n >> 1;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
    double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
    u64 i = tid;
    for (; i + 3 * nth < n2; i += 4 * nth) {
      const double2 a = __ldcg(s2 + i), b = __ldcg(s2 + i + nth), c = __ldcg(s2 + i + 2 * nth),
                    d = __ldcg(s2 + i + 3 * nth);
      d2[i] = a; d2[i + nth] = b; d2[i + 2 * nth] = c; d2[i + 3 * nth] = d;
    }
    for (; i < n2; i += nth) d2[i] = __ldcg(s2 + i);
    if ((n & 1ull) && tid == 0) dst[n - 1] = __ldcg(src + n - 1);
  } else {
    for (u64 i = tid; i < n; i += nth) dst[i] = __ldcg(src + i);
  }
}

__global__ void __launch_bounds__(256) k_push(const PushArgs a) {
  pdl_enter();
  __shared__ u64 par[MAX_WORLD];
  __shared__ bool last;
  if (threadIdx.x < a.npeer)  // parity of the staged message this launch carries to each peer
    par[threadIdx.x] = ((a.staged_mask >> threadIdx.x) & 1u) ? ((*a.sent_st[threadIdx.x] + 1ull) & 1ull) : 0ull;
  __syncthreads();
  for (int s = 0; s < a.nseg; ++s) {
    const CopySeg& g = a.seg[s];
    copy_segment(g.src, g.dst + (g.peer >= 0 ? par[g.peer] * g.parity_stride : 0ull), g.n);
  }
  if (a.npeer == 0) return;
  __threadfence_system();  // my remote stores are visible before the ticket says so
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < a.npeer) {
    const int p = threadIdx.x;
    __threadfence_system();
    const u64 seq = *a.sent[p] + 1ull;
    *a.sent[p] = seq;
    if ((a.staged_mask >> p) & 1u) *a.sent_st[p] += 1ull;
    st_release_sys(a.peer_flag[p], seq);
  }
  if (threadIdx.x == 0) *a.ticket = 0u;
}

// Only the wait of k_wait_unpack, one block: with NDSM_P2P_SPLIT_WAIT=1 it runs in front of k_wait_unpack, whose up
// to 128 blocks then find the flags already set instead of sitting on SM slots that the other component streams'
// kernels could use while the neighbour's message is still on its way.
__global__ void __launch_bounds__(32) k_wait_only(const WaitArgs a) {
  pdl_enter_no_trigger();
  if (threadIdx.x < a.npeer) {
    const int p = threadIdx.x;
    const u64 want = *a.expect[p] + 1ull;
    if (*a.d_err == 0) {
      const u64 t0 = global_timer_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(a.flag[p]) < want) {
        if ((++spins & 255u) == 0 && global_timer_ns() - t0 > a.timeout_ns) {
          *a.d_err = 1;
          *a.h_err_map = 1;
          break;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_wait_unpack(const WaitArgs a) {
  pdl_enter_no_trigger();
  __shared__ u64 par[MAX_WORLD];
  __shared__ bool last;
  if (threadIdx.x < a.npeer) {
    const int p = threadIdx.x;
    const u64 want = *a.expect[p] + 1ull;
    if (*a.d_err == 0) {
      const u64 t0 = global_timer_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(a.flag[p]) < want) {
        if ((++spins & 255u) == 0 && global_timer_ns() - t0 > a.timeout_ns) {
          *a.d_err = 1;
          *a.h_err_map = 1;
          break;
        }
      }
    }
    par[p] = ((a.staged_mask >> p) & 1u) ? ((*a.expect_st[p] + 1ull) & 1ull) : 0ull;
  }
  __syncthreads();
  for (int s = 0; s < a.nseg; ++s) {
    const CopySeg& g = a.seg[s];
    copy_segment(g.src + (g.peer >= 0 ? par[g.peer] * g.parity_stride : 0ull), g.dst, g.n);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;  // every block has read the counters before the last one advances them
  if (threadIdx.x < a.npeer) {
    const int p = threadIdx.x;
    *a.expect[p] += 1ull;
    if ((a.staged_mask >> p) & 1u) *a.expect_st[p] += 1ull;
  }
  if (threadIdx.x == 0) *a.ticket = 0u;
}

int blocks_for(const CopySeg* seg, int nseg) {
  u64 tot = 0;
  for (int s = 0; s < nseg; ++s) tot += seg[s].n;
  // 16 KB per block keeps ~100 blocks of remote stores in flight for a 6-plane halo (NVLink needs the
  // parallelism, the SMs do not notice 100 blocks), one block for the flag-only launches
  const u64 b = (tot * sizeof(double) + 16383) / 16384;
  return (int)(b < 1 ? 1 : (b > 128 ? 128 : b));
}

// ---------------------------------------------------------------------------------------------
// the communicator
// ---------------------------------------------------------------------------------------------
struct PeerComm : Comm {
  int ch;
  double* inbox = nullptr;  // [2 neighbours][2 parities][cap]
  size_t cap = 0;
  struct Send { int to; const double* src; size_t n; };
  struct Recv { int from; double* dst; size_t n; };
  struct Bc { int root; double* buf; size_t n; };
  std::vector<Send> sends;
  std::vector<Recv> recvs;
  std::vector<Bc> bcs;
  cudaStream_t alloc_stream;

  PeerComm(int channel, cudaStream_t st) : ch(channel), alloc_stream(st) { g_fab.ch_used[ch] = true; }
  ~PeerComm() override {
    if (inbox && g_fab.ok) g_fab.free(inbox);
    g_fab.ch_used[ch] = false;
  }
  int world() const override { return g_fab.world; }
  int first_rank() const override { return g_fab.rank; }
  int nlocal() const override { return 1; }
  const char* transport() const override { return "peer-memory stores over NVLink (CUDA IPC symmetric heap) + flags"; }
  bool failed() override { return g_fab.h_err && *g_fab.h_err != 0; }
  bool one_sided() const override { return true; }
  bool persistent() const override { return true; }
  void* sym_alloc(size_t bytes) override { return g_fab.alloc(bytes, alloc_stream); }
  void sym_free(void* p) override { g_fab.free(p); }
  void allgather_host(const void* s, void* r, size_t b, cudaStream_t st) override { g_fab.boot->allgather_host(s, r, b, st); }
  std::unique_ptr<Comm> split(int) override { return nullptr; }
  std::unique_ptr<Comm> clone(cudaStream_t st) override {
    for (int c = 0; c < MAX_CH; ++c)
      if (!g_fab.ch_used[c]) return std::unique_ptr<Comm>(new PeerComm(c, st));
    fprintf(stderr, "ERROR(peer):out of channels:NDSM_B200_ERR_INTERNAL\n");
    throw NdsmError(NDSM_ERR_INTERNAL);
  }
  void reserve(size_t doubles) override {
    doubles = (doubles + 63) / 64 * 64;
    if (doubles <= cap) return;
    if (inbox) g_fab.free(inbox);
    cap = doubles;
    inbox = static_cast<double*>(g_fab.alloc(4 * cap * sizeof(double), alloc_stream));
    ++epoch_;
  }
  unsigned long long epoch_ = 0;
  unsigned long long epoch() const override { return epoch_ ^ ((unsigned long long)(size_t)inbox << 8); }

  void begin(cudaStream_t) override { sends.clear(); recvs.clear(); bcs.clear(); }
  void send(int, int to, const double* src, size_t n, cudaStream_t) override { sends.push_back(Send{to, src, n}); }
  void recv(int, int from, double* dst, size_t n, cudaStream_t) override { recvs.push_back(Recv{from, dst, n}); }
  void bcast(int root, double* buf, size_t n, cudaStream_t) override { bcs.push_back(Bc{root, buf, n}); }

  // peer list bookkeeping of one launch
  struct PeerSet {
    int n = 0, rank_of[MAX_WORLD], idx_of[MAX_WORLD];
    PeerSet() { for (int r = 0; r < MAX_WORLD; ++r) idx_of[r] = -1; }
    int add(int r) {
      if (idx_of[r] < 0) { idx_of[r] = n; rank_of[n++] = r; }
      return idx_of[r];
    }
  };

  void launch_push(const std::vector<CopySeg>& segs, const PeerSet& ps, unsigned staged_mask, cudaStream_t st) {
    if (segs.size() > (size_t)MAX_SEG) {
      fprintf(stderr, "ERROR(peer):message has more than %d segments:NDSM_B200_ERR_INTERNAL\n", MAX_SEG);
      throw NdsmError(NDSM_ERR_INTERNAL);
    }
    PushArgs a;
    memset(&a, 0, sizeof a);
    a.nseg = (int)segs.size();
    for (size_t s = 0; s < segs.size(); ++s) a.seg[s] = segs[s];
    a.staged_mask = staged_mask;
    a.npeer = ps.n;
    for (int p = 0; p < ps.n; ++p) {
      const int r = ps.rank_of[p];
      a.peer_flag[p] = g_fab.peer(g_fab.word(CW_FLAGS, ch, g_fab.rank), r);  // rank r's flag word for (channel, me)
      a.sent[p] = g_fab.word(CW_SENT, ch, r);
      a.sent_st[p] = g_fab.word(CW_SENT_ST, ch, r);
    }
    a.ticket = g_fab.ticket(ch, 0);
    for (const CopySeg& g : segs)
      if (g.peer >= 0) g_peer_bytes += g.n * sizeof(double);
    if (ps.n > 0) ++g_peer_msgs;
    launch_k(k_push, blocks_for(a.seg, a.nseg), 256, 0, st, a);
    CUDA_CHECK(cudaGetLastError());
    ++g_launches;
  }

  void launch_wait(const std::vector<CopySeg>& segs, const PeerSet& ps, unsigned staged_mask, cudaStream_t st) {
    if (segs.size() > (size_t)MAX_SEG) {
      fprintf(stderr, "ERROR(peer):message has more than %d segments:NDSM_B200_ERR_INTERNAL\n", MAX_SEG);
      throw NdsmError(NDSM_ERR_INTERNAL);
    }
    WaitArgs a;
    memset(&a, 0, sizeof a);
    a.nseg = (int)segs.size();
    for (size_t s = 0; s < segs.size(); ++s) a.seg[s] = segs[s];
    a.npeer = ps.n;
    for (int p = 0; p < ps.n; ++p) {
      const int r = ps.rank_of[p];
      a.flag[p] = g_fab.word(CW_FLAGS, ch, r);
      a.expect[p] = g_fab.word(CW_EXPECT, ch, r);
      a.expect_st[p] = g_fab.word(CW_EXPECT_ST, ch, r);
    }
    a.staged_mask = staged_mask;
    a.ticket = g_fab.ticket(ch, 1);
    a.d_err = g_fab.d_err();
    a.h_err_map = g_fab.d_err_map;
    a.timeout_ns = g_fab.timeout_ns;
    const int nb = blocks_for(a.seg, a.nseg);
    const char* sw = getenv("NDSM_P2P_SPLIT_WAIT");
    if (nb > 1 && sw && atoi(sw) != 0) {
      launch_k(k_wait_only, 1, 32, 0, st, a);
      ++g_launches;
    }
    launch_k(k_wait_unpack, nb, 256, 0, st, a);
    CUDA_CHECK(cudaGetLastError());
    ++g_launches;
  }

  void end(cudaStream_t st) override {
    const int me = g_fab.rank, W = g_fab.world;
    std::vector<CopySeg> out, in;
    PeerSet pto, pfrom;
    unsigned staged_to = 0, staged_from = 0;
    size_t off_to[2] = {0, 0}, off_from[2] = {0, 0};  // running offsets inside the inbox of / from the lower, upper neighbour
    for (const Send& s : sends) {
      if (s.to != me - 1 && s.to != me + 1) throw NdsmError(NDSM_ERR_INTERNAL);  // staged messages go to z-neighbours
      const int d = (s.to == me + 1) ? 1 : 0;
      if (off_to[d] + s.n > cap) {
        fprintf(stderr, "ERROR(peer):staged message exceeds the reserved inbox:NDSM_B200_ERR_INTERNAL\n");
        throw NdsmError(NDSM_ERR_INTERNAL);
      }
      const int p = pto.add(s.to);
      staged_to |= 1u << p;
      // at the peer I am its lower neighbour (slot 0) when it sits above me, else its upper neighbour (slot 1)
      const int slot_at_peer = (s.to == me + 1) ? 0 : 1;
      CopySeg g;
      g.src = s.src;
      g.dst = g_fab.peer(inbox, s.to) + (size_t)slot_at_peer * 2 * cap + off_to[d];
      g.n = s.n;
      g.parity_stride = cap;
      g.peer = p;
      out.push_back(g);
      off_to[d] += s.n;
    }
    for (const Recv& r : recvs) {
      if (r.from != me - 1 && r.from != me + 1) throw NdsmError(NDSM_ERR_INTERNAL);
      const int d = (r.from == me + 1) ? 1 : 0;
      const int p = pfrom.add(r.from);
      staged_from |= 1u << p;
      CopySeg g;
      g.src = inbox + (size_t)d * 2 * cap + off_from[d];
      g.dst = r.dst;
      g.n = r.n;
      g.parity_stride = cap;
      g.peer = p;
      in.push_back(g);
      off_from[d] += r.n;
    }
    // a staged send must be answered by a staged receive from the same rank (double-buffer argument, file header)
    for (int p = 0; p < pto.n; ++p)
      if (pfrom.idx_of[pto.rank_of[p]] < 0) throw NdsmError(NDSM_ERR_INTERNAL);
    for (int p = 0; p < pfrom.n; ++p)
      if (pto.idx_of[pfrom.rank_of[p]] < 0) throw NdsmError(NDSM_ERR_INTERNAL);
    // broadcasts: the root writes its buffer straight into every other rank's copy
    for (const Bc& b : bcs) {
      if (b.root == me) {
        for (int r = 0; r < W; ++r) {
          if (r == me) continue;
          CopySeg g;
          g.src = b.buf;
          g.dst = g_fab.peer(b.buf, r);
          g.n = b.n;
          g.parity_stride = 0;
          g.peer = pto.add(r);
          out.push_back(g);
        }
      } else {
        pfrom.add(b.root);
      }
    }
    if (pto.n > 0) launch_push(out, pto, staged_to, st);
    if (pfrom.n > 0) launch_wait(in, pfrom, staged_from, st);
    sends.clear(); recvs.clear(); bcs.clear();
  }

  void gathern(int, const double* send, int n, double* recv_all, cudaStream_t st) override {
    const int me = g_fab.rank, W = g_fab.world;
    std::vector<CopySeg> out, none;
    PeerSet pall;
    for (int r = 0; r < W; ++r) {
      CopySeg g;
      g.src = send;
      g.n = (u64)n;
      g.parity_stride = 0;
      if (r == me) {
        g.dst = recv_all + (size_t)n * me;
        g.peer = -1;
      } else {
        g.dst = g_fab.peer(recv_all, r) + (size_t)n * me;
        g.peer = pall.add(r);
      }
      out.push_back(g);
    }
    launch_push(out, pall, 0u, st);
    launch_wait(none, pall, 0u, st);
  }

  void barrier(cudaStream_t st) override {
    std::vector<CopySeg> none;
    PeerSet pall;
    for (int r = 0; r < g_fab.world; ++r)
      if (r != g_fab.rank) pall.add(r);
    if (pall.n == 0) return;
    launch_push(none, pall, 0u, st);
    launch_wait(none, pall, 0u, st);
  }
};

}  // namespace

std::unique_ptr<Comm> make_peer_comm(Comm* boot, cudaStream_t st) {
  if (!boot || boot->world() < 2 || boot->nlocal() != 1) return nullptr;
  if (boot->world() > MAX_WORLD) return nullptr;
  if (g_fab.ok) return nullptr;  // one fabric per process: ndsm_b200_dist_finalize() first
  g_fab = Fabric();
  g_fab.boot = boot;
  g_fab.rank = boot->first_rank();
  g_fab.world = boot->world();
  for (int c = 0; c < MAX_CH; ++c) g_fab.ch_used[c] = false;
  if (const char* e = getenv("NDSM_P2P_TIMEOUT_MS")) g_fab.timeout_ns = (u64)atoll(e) * 1000000ull;
  try {
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_fab.h_err), 64, cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      throw NdsmError(NDSM_ERR_CUDA);
    }
    *g_fab.h_err = 0;
    CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_fab.d_err_map), g_fab.h_err, 0));
    size_t seg0 = SEG_MIN;
    if (const char* e = getenv("NDSM_P2P_HEAP_MB")) seg0 = (size_t)atoll(e) << 20;
    if (seg0 < Fabric::ctrl_bytes() + (1u << 20)) seg0 = Fabric::ctrl_bytes() + (1u << 20);
    g_fab.new_segment(seg0, st);
    g_fab.ctrl = reinterpret_cast<u64*>(g_fab.segs[0].base[g_fab.rank]);
    if (g_fab.segs[0].take(Fabric::ctrl_bytes()) != 0) throw NdsmError(NDSM_ERR_INTERNAL);  // control block at offset 0
    g_fab.ok = true;
  } catch (const NdsmError&) {
    fprintf(stderr, "WARNING(make_peer_comm):peer-memory transport unavailable, staying on %s\n", boot->transport());
    if (g_fab.h_err) cudaFreeHost(g_fab.h_err);
    g_fab = Fabric();
    return nullptr;
  }
  return std::unique_ptr<Comm>(new PeerComm(0, st));
}

void peer_fabric_shutdown() {
  g_fab.shutdown();
  g_fab = Fabric();
}

}  // namespace ndsm
