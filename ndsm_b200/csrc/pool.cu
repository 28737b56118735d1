// pool.cu -- see pool.hpp
#include "pool.hpp"

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ndsm {

namespace {
struct Pool {
  bool host;
  std::multimap<size_t, void*> free_blocks;   // size -> block
  std::unordered_map<void*, size_t> live;     // block -> size
  size_t cached = 0;

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
    auto it = free_blocks.find(b);
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
        if (!p) throw NdsmError(3);
      }
    }
    live[p] = b;
    return p;
  }
  void free(void* p) {
    if (!p) return;
    auto it = live.find(p);
    if (it == live.end()) { raw_free(p); return; }
    free_blocks.emplace(it->second, p);
    cached += it->second;
    live.erase(it);
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

}  // namespace ndsm
