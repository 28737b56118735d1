// common.cuh -- shared types for the B200 (sm_100a) NDSM hot path.
//
// Data layout in HBM ("colour-split" / checkerboard-compressed):
//   A level of global shape (nx,ny,nz) stores its red and black points in two separate
//   sub-arrays so that one colour pass of the red/black Gauss-Seidel smoother
//   (reference: ndsm_optimized.f90:103-167) streams each array exactly once:
//     colour c = (i+j+k)&1                     (0-based global indices)
//     offset   = c*cs + (k-k0)*ps + j*hp + (i>>1)
//   hp = half-row pitch (multiple of 8 doubles = 64 B), ps = plane stride (multiple of 32
//   doubles), cs = colour stride.  Neighbours of a point always have the other colour, and
//   within the other colour's row they sit at compressed index m-1+s / m+s (x) or m (y,z),
//   where m = i>>1 and s = i&1.  Padding entries are never written and stay zero.
//   2D faces are levels with nz == 1.
//   [k0,k0+nzl) are the z-planes owned by this rank (z-slab decomposition); pointers handed
//   to kernels point at local plane 0, halo planes live at plane -1 and nzl.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef long long i64;

struct Grid {
  int nx, ny, nz;  // global level shape (nz == 1 for a 2D face)
  int k0, nzl;     // owned z-planes [k0, k0+nzl)
  int hp;          // half-row pitch (doubles)
  int mcnt;        // compressed row length = (nx+1)/2
  i64 ps, cs;      // plane stride, colour stride (doubles)
};

struct Bounds {    // inclusive 0-based index range of non-Dirichlet points (ndsm_optimized.f90:68-76)
  int lb[3], ub[3];
};

// Restriction stencil capacity per dimension (see tables in mg.cu; ratio <= 8/3 -> <= 7 points)
#define NDSM_RMAX 8

#define CUDA_CHECK(call)                                                                      \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      fprintf(stderr, "ERROR(%s):%s:%s:%d\n", __func__, cudaGetErrorString(e__), __FILE__, __LINE__); \
      throw NdsmError(NDSM_ERR_CUDA, (int)e__);                                               \
    }                                                                                         \
  } while (0)

// Library error kinds carried by NdsmError.  The first five are the public NDSM_B200_ERR_* codes of
// include/ndsm_b200.h; NDSM_ERR_INTERNAL is a broken invariant of this library (reported as its own message).
// A CUDA runtime error never travels as a library code: it is NDSM_ERR_CUDA with the cudaError_t beside it.
enum { NDSM_ERR_SHAPE = 2, NDSM_ERR_CUDA = 3, NDSM_ERR_STENCIL = 4, NDSM_ERR_ARG = 5, NDSM_ERR_INTERNAL = 6 };
struct NdsmError {
  int code;      // one of the NDSM_ERR_* kinds
  int cuda_err;  // cudaError_t when code == NDSM_ERR_CUDA and the CUDA runtime reported it, else 0
  explicit NdsmError(int c, int ce = 0) : code(c), cuda_err(ce) {}
};

// Programmatic dependent launch (PDL): every kernel of this library starts with pdl_enter() and is launched with
// the programmatic-stream-serialization attribute (launch_k), in streams and in captured graphs alike.  The
// dependent grid's blocks are scheduled and run up to griddepcontrol.wait while the previous kernel is still
// draining; wait returns when that kernel has completed and its writes are visible.  Nothing above pdl_enter()
// touches global memory.  The trigger comes AFTER the wait, so at most one dependent grid is ever parked behind
// the running one.  The V-cycle's small levels are chains of 2-5 us kernels: the launch gap is a large share.
__device__ __forceinline__ void pdl_enter() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// For kernels that may SPIN on something another GPU produces (peer.cu, k_wait_unpack): no early trigger, so that
// no dependent grid is parked on the SMs while this one waits -- a parked full wave would keep the kernels of the
// other component streams (whose messages the peer is waiting for) off this GPU: a cross-GPU deadlock.
__device__ __forceinline__ void pdl_enter_no_trigger() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__host__ __device__ inline i64 gidx(const Grid& g, int i, int j, int k) {
  return (i64)((i + j + k) & 1) * g.cs + (i64)(k - g.k0) * g.ps + (i64)j * g.hp + (i >> 1);
}

// host side: launch with the PDL attribute (NDSM_B200_PDL=0 launches plainly); throws on a launch error
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}
