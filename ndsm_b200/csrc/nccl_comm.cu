// nccl_comm.cu -- slab communication over NCCL (NVLink 5 / NVSwitch), one process per GPU.
// NCCL is resolved at run time with dlopen so that ndsmf.so has no link-time dependency on it (the
// single-GPU drop-in path never touches NCCL).  Inside a torchrun worker the already-loaded
// libnccl.so.2 of PyTorch is picked up; otherwise the system library is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "mg.hpp"

namespace ndsm {

namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

bool load_nccl() {
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return false;
#define NDSM_SYM(field, name)                                            \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.h, name);      \
  if (!g_nccl.field) return false;
  NDSM_SYM(GetUniqueId, "ncclGetUniqueId")
  NDSM_SYM(CommInitRank, "ncclCommInitRank")
  NDSM_SYM(CommDestroy, "ncclCommDestroy")
  NDSM_SYM(Send, "ncclSend")
  NDSM_SYM(Recv, "ncclRecv")
  NDSM_SYM(AllGather, "ncclAllGather")
  NDSM_SYM(Broadcast, "ncclBroadcast")
  NDSM_SYM(GroupStart, "ncclGroupStart")
  NDSM_SYM(GroupEnd, "ncclGroupEnd")
  NDSM_SYM(GetErrorString, "ncclGetErrorString")
  NDSM_SYM(CommSplit, "ncclCommSplit")
  NDSM_SYM(CommCount, "ncclCommCount")
  NDSM_SYM(CommUserRank, "ncclCommUserRank")
#undef NDSM_SYM
  g_nccl.ok = true;
  return true;
}

#define NCCL_CHECK(call)                                                                         \
  do {                                                                                           \
    ncclResult_t r__ = (call);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      fprintf(stderr, "ERROR(%s):NCCL %s:%s:%d\n", __func__, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
      throw NdsmError(3);                                                                        \
    }                                                                                            \
  } while (0)

struct NcclComm : Comm {
  int rank_, world_;
  ncclComm_t comm = nullptr;
  NcclComm(int rank, int world, const ncclUniqueId& id) : rank_(rank), world_(world) {
    NCCL_CHECK(g_nccl.CommInitRank(&comm, world, id, rank));
  }
  explicit NcclComm(ncclComm_t c) : rank_(0), world_(1), comm(c) {
    NCCL_CHECK(g_nccl.CommCount(comm, &world_));
    NCCL_CHECK(g_nccl.CommUserRank(comm, &rank_));
  }
  std::unique_ptr<Comm> split(int colour) override {
    ncclComm_t nc = nullptr;
    NCCL_CHECK(g_nccl.CommSplit(comm, colour, rank_, &nc, nullptr));
    return std::unique_ptr<Comm>(new NcclComm(nc));
  }
  ~NcclComm() override {
    if (bar_) cudaFree(bar_);
    if (comm) g_nccl.CommDestroy(comm);
  }
  double* bar_ = nullptr;
  int world() const override { return world_; }
  int first_rank() const override { return rank_; }
  int nlocal() const override { return 1; }
  void begin(cudaStream_t) override { NCCL_CHECK(g_nccl.GroupStart()); }
  void send(int, int to_rank, const double* src, size_t n, cudaStream_t st) override {
    NCCL_CHECK(g_nccl.Send(src, n, ncclDouble, to_rank, comm, st));
  }
  void recv(int, int from_rank, double* dst, size_t n, cudaStream_t st) override {
    NCCL_CHECK(g_nccl.Recv(dst, n, ncclDouble, from_rank, comm, st));
  }
  void end(cudaStream_t) override { NCCL_CHECK(g_nccl.GroupEnd()); }
  void gathern(int, const double* send, int n, double* recv_all, cudaStream_t st) override {
    NCCL_CHECK(g_nccl.AllGather(send, recv_all, (size_t)n, ncclDouble, comm, st));
  }
  void bcast(int root_rank, double* buf, size_t n, cudaStream_t st) override {
    NCCL_CHECK(g_nccl.Broadcast(buf, buf, n, ncclDouble, root_rank, comm, st));
  }
  const char* transport() const override { return "NCCL send/recv groups"; }
  bool persistent() const override { return true; }
  void barrier(cudaStream_t st) override {  // a 2-double all-gather orders the ranks on the stream
    if (!bar_) CUDA_CHECK(cudaMalloc(&bar_, (2 * (size_t)world_ + 2) * sizeof(double)));
    NCCL_CHECK(g_nccl.AllGather(bar_ + 2 * world_, bar_, 2, ncclDouble, comm, st));
  }
  void allgather_host(const void* send, void* recv_all, size_t bytes, cudaStream_t st) override {
    char* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, bytes * ((size_t)world_ + 1)));
    CUDA_CHECK(cudaMemcpyAsync(d + bytes * world_, send, bytes, cudaMemcpyHostToDevice, st));
    NCCL_CHECK(g_nccl.AllGather(d + bytes * world_, d, bytes, ncclChar, comm, st));
    CUDA_CHECK(cudaMemcpyAsync(recv_all, d, bytes * world_, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d);
  }
  std::unique_ptr<Comm> clone(cudaStream_t st) override {
    // rank 0 draws a fresh unique id and hands it to the others through this communicator
    ncclUniqueId id;
    memset(&id, 0, sizeof id);
    if (rank_ == 0) NCCL_CHECK(g_nccl.GetUniqueId(&id));
    char* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, sizeof id));
    CUDA_CHECK(cudaMemcpyAsync(d, &id, sizeof id, cudaMemcpyHostToDevice, st));
    NCCL_CHECK(g_nccl.Broadcast(d, d, sizeof id, ncclChar, 0, comm, st));
    CUDA_CHECK(cudaMemcpyAsync(&id, d, sizeof id, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d);
    return std::unique_ptr<Comm>(new NcclComm(rank_, world_, id));
  }
};
}  // namespace

bool nccl_unique_id(void* out128) {
  if (!load_nccl()) return false;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return false;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return true;
}

std::unique_ptr<Comm> make_nccl_comm(int rank, int world, const void* id128) {
  if (!load_nccl()) {
    fprintf(stderr, "ERROR(make_nccl_comm):libnccl.so.2 not found:NDSM_B200_ERR_CUDA\n");
    throw NdsmError(3);
  }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  return std::unique_ptr<Comm>(new NcclComm(rank, world, id));
}

}  // namespace ndsm
