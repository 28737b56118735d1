// hostsink.cu -- see hostsink.hpp
#include "hostsink.hpp"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "pool.hpp"

namespace ndsm {

namespace {
constexpr size_t CHUNK = 4u << 20;  // bytes per staged chunk
int worker_count() {
  if (const char* e = std::getenv("NDSM_B200_D2H_THREADS")) return std::max(1, std::min(32, std::atoi(e)));
  const unsigned hc = std::thread::hardware_concurrency();
  return (int)std::max(2u, std::min(8u, hc / 2));
}
bool page_locked(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
}  // namespace

HostSink::HostSink(int device) : dev_(device) {
  CUDA_CHECK(cudaStreamCreateWithFlags(&direct_, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_, cudaEventDisableTiming));
}

HostSink::~HostSink() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto& w : workers_) if (w.th.joinable()) w.th.join();
  for (auto& w : workers_) if (w.st) cudaStreamDestroy(w.st);
  if (direct_) { cudaStreamSynchronize(direct_); cudaStreamDestroy(direct_); }
  if (ev_) cudaEventDestroy(ev_);
}

void HostSink::start_workers() {
  const int n = worker_count();
  workers_.resize(n);
  for (int w = 0; w < n; ++w) CUDA_CHECK(cudaStreamCreateWithFlags(&workers_[w].st, cudaStreamNonBlocking));
  for (int w = 0; w < n; ++w) workers_[w].th = std::thread([this, w] { run(w); });
}

void HostSink::push(double* dst, const double* src, size_t n, cudaStream_t producer) {
  if (n == 0) return;
  CUDA_CHECK(cudaEventRecord(ev_, producer));
  const size_t bytes = n * sizeof(double);
  if (page_locked(dst)) {
    CUDA_CHECK(cudaStreamWaitEvent(direct_, ev_, 0));
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, direct_));
    return;
  }
  if (workers_.empty()) start_workers();
  const int nw = (int)workers_.size();
  // every worker stream waits for the producer here, in the caller's thread, so the event can be re-recorded
  for (int w = 0; w < nw; ++w) CUDA_CHECK(cudaStreamWaitEvent(workers_[w].st, ev_, 0));
  // contiguous, chunk-aligned shares: each worker streams through its own part of the array
  const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
  {
    std::lock_guard<std::mutex> lk(mu_);
    size_t c0 = 0;
    for (int w = 0; w < nw; ++w) {
      const size_t c1 = nchunks * (size_t)(w + 1) / nw;
      if (c1 > c0) {
        const size_t b0 = c0 * CHUNK, b1 = std::min(bytes, c1 * CHUNK);
        workers_[w].q.push_back(Job{reinterpret_cast<char*>(dst) + b0, reinterpret_cast<const char*>(src) + b0, b1 - b0});
        ++pending_;
      }
      c0 = c1;
    }
  }
  cv_.notify_all();
}

void HostSink::run(int w) {
  cudaSetDevice(dev_);
  Worker& me = workers_[w];
  char* slot[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  try {
    for (int s = 0; s < 2; ++s) {
      slot[s] = static_cast<char*>(pool_alloc_host(CHUNK));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming | cudaEventBlockingSync));
    }
  } catch (const NdsmError& e) {
    err_.store(e.code ? e.code : 3);
  }
  while (true) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return stop_ || !me.q.empty(); });
      if (me.q.empty()) break;  // stop requested and nothing left
      job = me.q.front();
      me.q.pop_front();
    }
    if (err_.load() == 0) {
      const size_t nc = (job.bytes + CHUNK - 1) / CHUNK;
      auto issue = [&](size_t i) {
        const size_t off = i * CHUNK, len = std::min(CHUNK, job.bytes - off);
        cudaError_t e = cudaMemcpyAsync(slot[i & 1], job.src + off, len, cudaMemcpyDeviceToHost, me.st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], me.st);
        if (e != cudaSuccess) err_.store(3);
      };
      issue(0);
      for (size_t i = 0; i < nc && err_.load() == 0; ++i) {
        if (i + 1 < nc) issue(i + 1);
        if (cudaEventSynchronize(ev[i & 1]) != cudaSuccess) { err_.store(3); break; }
        const size_t off = i * CHUNK, len = std::min(CHUNK, job.bytes - off);
        std::memcpy(job.dst + off, slot[i & 1], len);
      }
      if (err_.load() != 0) cudaStreamSynchronize(me.st);  // nothing may still write into the slots
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      --pending_;
    }
    idle_.notify_all();
  }
  cudaStreamSynchronize(me.st);
  for (int s = 0; s < 2; ++s) {
    if (ev[s]) cudaEventDestroy(ev[s]);
    if (slot[s]) pool_free_host(slot[s]);
  }
}

void HostSink::wait() {
  {
    std::unique_lock<std::mutex> lk(mu_);
    idle_.wait(lk, [&] { return pending_ == 0; });
  }
  CUDA_CHECK(cudaStreamSynchronize(direct_));
  if (int e = err_.load()) throw NdsmError(e);
}

}  // namespace ndsm
