// Shared helpers for the d3rlpy_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/d3rlpy_b200.h"

namespace d3b {

// thread-local last-error text, readable through d3b_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);
void count_launch();
long long launches();

inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(D3B_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return D3B_OK;
}

#define D3B_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return d3b::set_err(D3B_ERR_ARG, __VA_ARGS__);  \
  } while (0)

#define D3B_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) return d3b::set_err(D3B_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

constexpr int kNumSM = 148;  // B200

// ---- programmatic dependent launch (PDL): batch-256 updates are launch-latency bound, so consecutive kernels of
// the update graph overlap the prologue of kernel N+1 (barrier init, TMEM allocation, descriptor prefetch, block
// scheduling) with the tail of kernel N.  Every kernel launched through launch_pdl() calls pdl_trigger() first
// and pdl_wait() before its first global-memory access (griddepcontrol.wait returns once all prerequisite grids
// have completed and their writes are visible).
bool pdl_enabled();
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// the same with a thread-block cluster of `cluster_z` CTAs along grid z (grid.z must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      unsigned cluster_z, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 1;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = cluster_z;
  cfg.attrs = attr;
  cfg.numAttrs = cluster_z > 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < (int)((blockDim.x + 31) >> 5)) ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

}  // namespace d3b
