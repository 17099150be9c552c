// pg_common.cuh — error plumbing and small device helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pagegeom.h"
#include "pg_math.h"

void pg_set_error(const char* fmt, ...);

#define PG_CUDA_TRY(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      pg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PG_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define PG_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) {                            \
      pg_set_error("invalid argument: %s", msg); \
      return PG_ERR_INVALID;                  \
    }                                         \
  } while (0)

#define PG_LAUNCH_CHECK()                                                       \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      pg_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PG_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

// Checked build (-DPG_CHECKED; multimodal_embeddings_b200/build.py --checked writes _variants/libpagegeom_checked.so):
// device-side assertions on every computed index into a workspace, a staging ring or a stream.  A violation traps:
// the launch fails with cudaErrorAssert and the caller sees PG_ERR_CUDA.  compute-sanitizer is closed on the GPU
// pool this was developed on; the GPU test-suite is run once per round against this build instead
// (PAGEGEOM_LIB=..., profiles/r02_checked_build.md).  The default build compiles the checks away.
#ifdef PG_CHECKED
#include <assert.h>
#define PG_DEV_ASSERT(c) assert(c)
#else
#define PG_DEV_ASSERT(c) ((void)0)
#endif

#ifdef __CUDACC__
__device__ __forceinline__ int pg_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int pg_warp() { return threadIdx.x >> 5; }

// Block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the block sum.  `smem` needs 33 ints.
__device__ __forceinline__ int pg_block_exscan(int v, int* smem, int* total) {
  const int lane = pg_lane(), warp = pg_warp();
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;           // exclusive warp offsets
    if (lane == 31) smem[32] = winc; // block total
  }
  __syncthreads();
  const int res = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}
#endif
