// pool.cu -- see pool.hpp
#include "pool.hpp"

#include <cstdlib>
#include <iterator>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ndsm {

namespace {
// Device blocks are keyed by (device ordinal, size): a block cached while GPU a was current is never handed
// out for kernels on GPU b (NDSM_DEVICE / the current device may change between calls).  Pinned host blocks
// are usable from every device and share key 0.
typedef std::pair<int, size_t> Key;
struct Pool {
  bool host;
  std::multimap<Key, void*> free_blocks;      // (device, size) -> block
  std::unordered_map<void*, Key> live;        // block -> (device, size)
  size_t cached = 0;
  size_t cap = (size_t)-1;                    // cached bytes above this are returned to the driver on free

  int device_key() const {
    if (host) return 0;
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
    return d;
  }

  static size_t round(size_t b) { return (b + 511) / 512 * 512; }

  void* raw_alloc(size_t b) {
    void* p = nullptr;
    cudaError_t e = host ? cudaMallocHost(&p, b) : cudaMalloc(&p, b);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  void raw_free(void* p) { host ? cudaFreeHost(p) : cudaFree(p); }

  void* alloc(size_t bytes) {
    const size_t b = round(bytes ? bytes : 1);
    const Key key(device_key(), b);
    // among cached blocks of this size take the one freed LAST: scoped buffers are released in reverse order of
    // their allocation, so a call sequence that repeats gets the same block for the same buffer every time
    // (addresses are baked into captured graphs; first-in-first-out permuted them from call to call)
    auto range = free_blocks.equal_range(key);
    auto it = (range.first == range.second) ? free_blocks.end() : std::prev(range.second);
    void* p = nullptr;
    if (it != free_blocks.end()) {
      p = it->second;
      free_blocks.erase(it);
      cached -= b;
    } else {
      p = raw_alloc(b);
      if (!p) {  // out of memory: drop the cache and retry once
        release();
        p = raw_alloc(b);
        if (!p) throw NdsmError(NDSM_ERR_CUDA, (int)cudaErrorMemoryAllocation);
      }
    }
    live[p] = key;
    return p;
  }
  void free(void* p) {
    if (!p) return;
    auto it = live.find(p);
    if (it == live.end()) { raw_free(p); return; }
    const Key key = it->second;
    live.erase(it);
    if (cached + key.second > cap) {  // over the cap: hand the block back instead of caching it
      raw_free(p);
      return;
    }
    free_blocks.emplace(key, p);
    cached += key.second;
  }
  void trim(size_t keep) {  // release cached blocks, largest first, until at most `keep` bytes stay cached
    while (cached > keep && !free_blocks.empty()) {
      auto big = free_blocks.begin();
      for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it)
        if (it->first.second > big->first.second) big = it;
      raw_free(big->second);
      cached -= big->first.second;
      free_blocks.erase(big);
    }
  }
  void release() {
    for (auto& kv : free_blocks) raw_free(kv.second);
    free_blocks.clear();
    cached = 0;
  }
};
Pool g_dev{false}, g_host{true};
std::mutex g_mu;  // the copy workers of HostSink allocate their pinned slots from their own threads
using Lock = std::lock_guard<std::mutex>;
}  // namespace

void* pool_alloc(size_t bytes) { Lock l(g_mu); return g_dev.alloc(bytes); }
void pool_free(void* p) { Lock l(g_mu); g_dev.free(p); }
void* pool_alloc_host(size_t bytes) { Lock l(g_mu); return g_host.alloc(bytes); }
void pool_free_host(void* p) { Lock l(g_mu); g_host.free(p); }
void pool_release() { Lock l(g_mu); g_dev.release(); g_host.release(); }
size_t pool_cached_bytes() { Lock l(g_mu); return g_dev.cached; }
void pool_trim(size_t keep_bytes) { Lock l(g_mu); g_dev.trim(keep_bytes); }
void pool_trim_to_cap() {
  size_t cap_mb = 24576;
  if (const char* e = getenv("NDSM_B200_WORKSPACE_CAP_MB")) cap_mb = (size_t)strtoull(e, nullptr, 10);
  if (getenv("NDSM_B200_KEEP_WORKSPACE") && atoi(getenv("NDSM_B200_KEEP_WORKSPACE")) != 0) return;
  pool_trim(cap_mb << 20);
}

}  // namespace ndsm
