// pool.hpp -- caching allocator for device and pinned host memory.
// A vector-potential solve allocates the same set of buffers every call (multi-GB level arrays at 513^3);
// cudaMalloc/cudaFree of such sizes costs 10s-100s of ms and synchronises the device, so freed blocks are
// kept and handed out again on an exact size match (same device).  ndsm_b200_release_workspace() returns them
// to CUDA; the public entry points trim the cache to NDSM_B200_WORKSPACE_CAP_MB (default 24576 MiB; 0 = the
// reference's ownership: everything is freed before the call returns, ndsm_vector_potential.f90:489-495).
#pragma once
#include <cstddef>

namespace ndsm {
void* pool_alloc(size_t bytes);       // device memory; throws NdsmError(3) when CUDA is out of memory
void pool_free(void* p);              // caller guarantees no GPU work still uses p
void* pool_alloc_host(size_t bytes);  // pinned host memory
void pool_free_host(void* p);
void pool_release();                  // give every cached block back to CUDA
size_t pool_cached_bytes();
void pool_trim(size_t keep_bytes);    // release cached device blocks (largest first) down to keep_bytes
void pool_trim_to_cap();              // pool_trim(NDSM_B200_WORKSPACE_CAP_MB), read on every call
}  // namespace ndsm
