// sym_alloc.hpp -- offset allocator of one symmetric-heap segment (peer.cu).  Host-only, no CUDA.
//
// Every rank drives its copy with the same sequence of take / give calls and the same sizes, so every rank gets the
// same offsets: a buffer then sits at the same offset of every rank's copy of the segment, which is what lets a rank
// compute a peer's address as base[peer] + offset.  First fit over an offset-ordered free list, with coalescing.
#pragma once
#include <cstddef>
#include <iterator>
#include <map>

namespace ndsm {

struct SegmentAllocator {
  std::map<size_t, size_t> free_list;  // offset -> size
  std::map<size_t, size_t> live;       // offset -> size
  void reset(size_t bytes) {
    free_list.clear();
    live.clear();
    free_list[0] = bytes;
  }
  size_t take(size_t b) {  // returns the offset or (size_t)-1
    for (auto it = free_list.begin(); it != free_list.end(); ++it)
      if (it->second >= b) {
        const size_t off = it->first, rest = it->second - b;
        free_list.erase(it);
        if (rest) free_list[off + b] = rest;
        live[off] = b;
        return off;
      }
    return (size_t)-1;
  }
  void give(size_t off) {
    auto lv = live.find(off);
    if (lv == live.end()) return;
    size_t sz = lv->second;
    live.erase(lv);
    auto nx = free_list.lower_bound(off);
    if (nx != free_list.end() && off + sz == nx->first) {  // merge with the following block
      sz += nx->second;
      nx = free_list.erase(nx);
    }
    if (nx != free_list.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == off) {  // merge with the preceding block
        pv->second += sz;
        return;
      }
    }
    free_list[off] = sz;
  }
};

}  // namespace ndsm
