// hostsink.hpp -- device-to-host delivery of result arrays into caller memory of unknown kind.
// The reference's callers (ndsm.py:160-175) hand the library ordinary numpy arrays, i.e. PAGEABLE memory.
// A cudaMemcpyAsync into pageable memory is staged by the driver inside the calling thread and, measured on
// the B200 box at 513^3, both runs at ~20 GB/s and slows the kernel launches of the solve that is in flight.
// HostSink does the staging itself: worker threads pull 4 MB chunks into their own pinned double buffers
// over their own streams and memcpy them out, several chunks in flight; page-locked destinations (detected
// with cudaPointerGetAttributes) are copied directly.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace ndsm {

class HostSink {
 public:
  explicit HostSink(int device);
  ~HostSink();
  HostSink(const HostSink&) = delete;
  HostSink& operator=(const HostSink&) = delete;
  // copy n doubles from device `src` to host `dst` once the work enqueued so far on `producer` has finished;
  // returns immediately
  void push(double* dst, const double* src, size_t n, cudaStream_t producer);
  // block until everything pushed has landed in host memory; throws NdsmError on a CUDA failure
  void wait();

 private:
  struct Job { char* dst; const char* src; size_t bytes; };
  struct Worker {
    std::thread th;
    cudaStream_t st = nullptr;
    std::deque<Job> q;
  };
  void start_workers();
  void run(int w);
  const int dev_;
  cudaStream_t direct_ = nullptr;  // copies into page-locked destinations
  cudaEvent_t ev_ = nullptr;
  std::vector<Worker> workers_;
  std::mutex mu_;
  std::condition_variable cv_, idle_;
  size_t pending_ = 0;
  bool stop_ = false;
  std::atomic<int> err_{0};
};

}  // namespace ndsm
