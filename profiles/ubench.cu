// Micro-benchmarks behind the fused-MLP design decisions: how fast can ONE SM
//  (1) stream weight K blocks from L2 into shared memory with TMA (latency of one box, pipelined throughput vs ring
//      depth, alone and with all 148 SMs streaming the same matrix), and
//  (2) drain a TMEM accumulator with tcgen05.ld (bytes/clk vs number of warps and loads in flight).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/ubench profiles/ubench.cu -lcudart
// (no libcuda link: cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint).  Run on the B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ int g_poll_mode = 0;  // 0: try_wait (may suspend the thread), 1: test_wait (pure polling)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (g_poll_mode == 0) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
      if (spins > (1u << 24)) __trap();
  } else {
    for (uint32_t spins = 0; !mbar_test_wait(bar, parity); ++spins)
      if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// thread 0 streams `n_loads` boxes ({64 cols, box_rows}) of a [rows_total x 256] bf16 matrix through a ring of `stages`
// buffers; a buffer is re-armed as soon as its previous load has landed (an infinitely fast consumer).
__global__ void tma_stream_kernel(const __grid_constant__ CUtensorMap map, int stages, int box_bytes, int n_loads,
                                  int n_row_boxes, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bars[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // warm-up: descriptor fetch + one box
    mbar_expect_tx(bars, box_bytes);
    tma_load_3d(smem, &map, bars, 0, 0, 0);
    mbar_wait(bars, 0);
    long long t0 = clock64();
    uint32_t par = 1;  // stage 0 has completed one phase
    for (int i = 0; i < n_loads; ++i) {
      const int s = i % stages;
      if (i >= stages || s == 0) {
        if (i >= stages) {
          mbar_wait(bars + s, (par >> s) & 1);
          par ^= 1u << s;
        }
      }
      mbar_expect_tx(bars + s, box_bytes);
      tma_load_3d(smem + (size_t)s * box_bytes, &map, bars + s, (i & 3) * 64, ((i >> 2) % n_row_boxes) * (box_bytes / 128), 0);
    }
    for (int s = 0; s < stages && s < n_loads; ++s) {
      mbar_wait(bars + s, (par >> s) & 1);
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

// thread 0 issues `n` TMA tensor stores of one {64 cols, 128 rows} box (16 KB), at most `inflight` bulk groups pending
__global__ void tma_store_kernel(const __grid_constant__ CUtensorMap map, int n, int inflight, int n_row_boxes,
                                 long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&map),
                   "r"(smem_u32(smem + (i & 3) * 16384)), "r"((i & 3) * 64), "r"(((i >> 2) % n_row_boxes) * 128), "r"(0)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (inflight == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (inflight == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    long long t1 = clock64();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
}

// 1-D bulk copies (no tensor map): `n` copies of `bytes` shared -> global / global -> shared through `stages` buffers
__global__ void bulk_1d_kernel(uint8_t* g, int bytes, int n, int stages, int store, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint8_t* mine = g + (size_t)blockIdx.x * 4 * 65536;
    long long t0 = clock64();
    if (store) {
      for (int i = 0; i < n; ++i) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((uint64_t)(mine + (size_t)(i & 3) * bytes)),
                     "r"(smem_u32(smem + (size_t)(i % stages) * bytes)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (stages == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else {
      uint32_t par = 0;
      for (int i = 0; i < n; ++i) {
        const int s = i % stages;
        if (i >= stages) {
          mbar_wait(bars + s, (par >> s) & 1);
          par ^= 1u << s;
        }
        mbar_expect_tx(bars + s, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + (size_t)s * bytes)), "l"((uint64_t)(mine + (size_t)(i & 3) * bytes)), "r"(bytes),
                     "r"(smem_u32(bars + s)) : "memory");
      }
      for (int s = 0; s < stages && s < n; ++s) mbar_wait(bars + s, (par >> s) & 1);
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

// `warps` warps (x4 TMEM lane quarters) each read `iters` x `depth` 16-column chunks; depth loads are issued back to back
// before one tcgen05.wait::ld.
template <int DEPTH>
__global__ void tmem_read_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[DEPTH][16];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const uint32_t col = (uint32_t)(((warp >> 2) * DEPTH + d) * 16 + it * 64) & 511u & ~15u;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[d][0]), "=r"(v[d][1]), "=r"(v[d][2]), "=r"(v[d][3]), "=r"(v[d][4]), "=r"(v[d][5]), "=r"(v[d][6]),
                     "=r"(v[d][7]), "=r"(v[d][8]), "=r"(v[d][9]), "=r"(v[d][10]), "=r"(v[d][11]), "=r"(v[d][12]),
                     "=r"(v[d][13]), "=r"(v[d][14]), "=r"(v[d][15])
                   : "r"(base + col));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= v[d][i];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

// epilogue skeleton: 16 warps, per iteration one 16-column TMEM chunk -> bias/ReLU/bf16 -> 2 x 16-byte swizzled shared
// stores, then optionally fence.proxy.async and/or a per-warp mbarrier arrive / a 512-thread named barrier
__global__ void epi_skeleton_kernel(int iters, int fence_every, int sync_mode, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) mbar_init(&bar, 1u << 20);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3, sub = warp >> 2, row = q * 32 + lane;
  const uint32_t base = slot + ((uint32_t)(q * 32) << 16);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int kb = it & 3;
    const int c = kb * 64 + sub * 16;
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(base + (uint32_t)c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float f0 = fmaxf(__uint_as_float(v[2 * i]) + 0.25f, 0.f), f1 = fmaxf(__uint_as_float(v[2 * i + 1]) + 0.5f, 0.f);
      uint32_t r;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f1), "f"(f0));
      pk[i] = r;
    }
    const int j0 = (c & 63) >> 3;
    uint8_t* b = smem + kb * 16384 + row * 128;
    *reinterpret_cast<uint4*>(b + ((j0 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(b + (((j0 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    if (fence_every && (it % fence_every) == fence_every - 1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (sync_mode == 1) {
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
      } else if (sync_mode == 2) {
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (smem[threadIdx.x] == 77 && iters < 0) sink[0] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

// wide epilogue step: 16 warps, per iteration one 64-column K block per thread-row (2 x tcgen05.ld.x32), bias from
// BIAS = 0 immediate / 1 scalar __ldg / 2 float4 __ldg / 3 float4 shared broadcast, ReLU, bf16, 8 swizzled 16-byte stores,
// one fence.proxy.async
template <int BIAS, int CVT>
__global__ void epi_wide_kernel(int iters, const float* __restrict__ bias_g, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t slot;
  __shared__ __align__(16) float bias_s[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 256) bias_s[threadIdx.x] = bias_g[threadIdx.x];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3, sub = warp >> 2, row = q * 32 + lane;
  const uint32_t base = slot + ((uint32_t)(q * 32) << 16);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint8_t* b = smem + sub * 16384 + row * 128;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = sub * 64 + half * 32;
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(base + (uint32_t)c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float bb[8];
        if (BIAS == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = 0.25f;
        } else if (BIAS == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = __ldg(bias_g + c + 8 * j + i);
        } else {
          const float4* src = BIAS == 2 ? reinterpret_cast<const float4*>(bias_g + c + 8 * j)
                                        : reinterpret_cast<const float4*>(bias_s + c + 8 * j);
          float4 x0 = BIAS == 2 ? __ldg(src) : src[0], x1 = BIAS == 2 ? __ldg(src + 1) : src[1];
          bb[0] = x0.x; bb[1] = x0.y; bb[2] = x0.z; bb[3] = x0.w; bb[4] = x1.x; bb[5] = x1.y; bb[6] = x1.z; bb[7] = x1.w;
        }
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f0 = fmaxf(__uint_as_float(v[8 * j + 2 * i]) + bb[2 * i], 0.f);
          float f1 = fmaxf(__uint_as_float(v[8 * j + 2 * i + 1]) + bb[2 * i + 1], 0.f);
          if (CVT) {
            uint32_t r;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f1), "f"(f0));
            pk[i] = r;
          } else {
            pk[i] = (__float_as_uint(f0) >> 16) | (__float_as_uint(f1) & 0xFFFF0000u);
          }
        }
        const int chunk = half * 4 + j;
        *reinterpret_cast<uint4*>(b + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (smem[threadIdx.x] == 77 && iters < 0) sink[0] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int rows_total = 2048, cols = 256;  // 1 MB of "weights": L2 resident
  void* w;
  CK(cudaMalloc(&w, (size_t)rows_total * cols * 2));
  CK(cudaMemset(w, 1, (size_t)rows_total * cols * 2));
  long long* out;
  CK(cudaMalloc(&out, 148 * sizeof(long long)));
  uint32_t* sink;
  CK(cudaMalloc(&sink, 4));
  long long h[148];
  CK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int n_loads = 96;
  printf("== TMA stream: 64-column SWIZZLE_128B boxes of an L2-resident bf16 matrix, cycles per box / bytes per clk per SM\n");
  for (int box_rows : {256, 128, 64}) {
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows_total, 1};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * rows_total};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap map;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = box_rows * 128;
    for (int mode : {0, 1}) {
     CK(cudaMemcpyToSymbol(g_poll_mode, &mode, sizeof(int)));
     printf("-- mbarrier wait = %s\n", mode ? "test_wait polling" : "try_wait");
     for (int grid : {1, 148}) {
      for (int stages : {1, 2, 4, 6, 12, 24}) {
        if ((size_t)stages * box_bytes > 192 * 1024) continue;
        for (int rep = 0; rep < 2; ++rep) {
          tma_stream_kernel<<<grid, 32, 1024 + (size_t)stages * box_bytes>>>(map, stages, box_bytes, n_loads,
                                                                             rows_total / box_rows, out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        double mean = 0;
        for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mean += (double)h[i] / grid; }
        printf("box %3d rows (%2d KB) grid %3d stages %d: %7.0f cycles/box (max CTA %7.0f)  %6.1f B/clk/SM\n", box_rows,
               box_bytes / 1024, grid, stages, mean / n_loads, (double)mx / n_loads, (double)box_bytes * n_loads / mean);
      }
     }
    }
  }
  {
    void* hbuf;
    const int rows_h = 128 * 64;
    CK(cudaMalloc(&hbuf, (size_t)rows_h * cols * 2));
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows_h, 1};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * rows_h};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap map;
    enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, hbuf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    printf("== TMA tensor stores, 16 KB boxes {64 cols x 128 rows}: cycles per box until smem read / until complete\n");
    for (int inflight : {1, 4, 64}) {
      for (int rep = 0; rep < 2; ++rep) {
        tma_store_kernel<<<1, 32, 1024 + 65536>>>(map, 64, inflight, 64, out);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, out, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      printf("in flight <= %2d: %6.0f cycles/box (read), %6.0f (complete)\n", inflight, h[0] / 64.0, h[1] / 64.0);
    }
    uint8_t* gb;
    CK(cudaMalloc(&gb, (size_t)148 * 4 * 65536));
    CK(cudaMemset(gb, 1, (size_t)148 * 4 * 65536));
    CK(cudaFuncSetAttribute(bulk_1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("== 1-D bulk copies (cp.async.bulk, no tensor map): cycles per copy / bytes per clk per SM\n");
    for (int store : {0, 1}) {
      for (int bytes : {8192, 16384, 32768, 65536}) {
        for (int grid : {1, 148}) {
          for (int stages : {1, 2}) {
            for (int rep = 0; rep < 2; ++rep) {
              bulk_1d_kernel<<<grid, 32, 1024 + (size_t)stages * bytes>>>(gb, bytes, 64, stages, store, out);
              CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int i = 0; i < grid; ++i) mean += (double)h[i] / grid;
            printf("%s %2d KB grid %3d buffers %d: %7.0f cycles/copy  %6.1f B/clk/SM\n", store ? "smem->global" : "global->smem",
                   bytes / 1024, grid, stages, mean / 64, 64.0 * bytes / mean);
          }
        }
      }
    }
  }
  CK(cudaFuncSetAttribute(epi_skeleton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  printf("== epilogue skeleton, 16 warps: cycles per 16-column piece (x4 = one 256-column layer)\n");
  {
    float* bg;
    CK(cudaMalloc(&bg, 1024));
    CK(cudaMemset(bg, 0, 1024));
    CK(cudaFuncSetAttribute(epi_wide_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    CK(cudaFuncSetAttribute(epi_wide_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    CK(cudaFuncSetAttribute(epi_wide_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    CK(cudaFuncSetAttribute(epi_wide_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    CK(cudaFuncSetAttribute(epi_wide_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    const char* names[5] = {"immediate bias", "scalar __ldg bias", "float4 __ldg bias", "float4 shared bias", "immediate bias, no cvt"};
    for (int v = 0; v < 5; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        if (v == 0) epi_wide_kernel<0, 1><<<1, 512, 1024 + 65536>>>(256, bg, out, sink);
        if (v == 1) epi_wide_kernel<1, 1><<<1, 512, 1024 + 65536>>>(256, bg, out, sink);
        if (v == 2) epi_wide_kernel<2, 1><<<1, 512, 1024 + 65536>>>(256, bg, out, sink);
        if (v == 3) epi_wide_kernel<3, 1><<<1, 512, 1024 + 65536>>>(256, bg, out, sink);
        if (v == 4) epi_wide_kernel<0, 0><<<1, 512, 1024 + 65536>>>(256, bg, out, sink);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost));
      printf("wide epilogue step (128x256 tile, 16 warps), %s: %6.0f cycles/tile\n", names[v], h[0] / 256.0);
    }
  }
  for (int fe : {0, 1, 2, 4}) {
    for (int sm : {0, 2}) {
      if (fe == 0 && sm) continue;
      for (int rep = 0; rep < 2; ++rep) {
        epi_skeleton_kernel<<<1, 512, 1024 + 65536>>>(1024, fe, sm, out, sink);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost));
      printf("fence.proxy.async every %d pieces, sync %s: %6.0f cycles/piece\n", fe,
             sm == 0 ? "none" : sm == 1 ? "warp mbarrier arrive" : "bar.sync 512", h[0] / 1024.0);
    }
  }
  printf("== TMEM drain: tcgen05.ld.32x32b.x16, bytes per clk per SM\n");
  const int iters = 256;
  for (int warps : {4, 8, 16}) {
    for (int depth : {1, 2, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (depth == 1) tmem_read_kernel<1><<<1, warps * 32>>>(iters, out, sink);
        if (depth == 2) tmem_read_kernel<2><<<1, warps * 32>>>(iters, out, sink);
        if (depth == 4) tmem_read_kernel<4><<<1, warps * 32>>>(iters, out, sink);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost));
      double bytes = (double)warps * iters * depth * 2048.0;
      printf("warps %2d loads in flight %d: %8lld cycles  %6.1f B/clk\n", warps, depth, h[0], bytes / (double)h[0]);
    }
  }
  return 0;
}
