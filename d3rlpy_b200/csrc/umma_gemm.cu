// K2/K3 (bf16 mode): batched-ensemble dense layers on the 5th-gen tensor cores.
//
//   C[e][m][n] = sum_k A[e][m][k] * B[e][n][k]          (both operands K-major bf16, fp32 accumulate)
//
// tcgen05.mma (cta_group::1, M=128, N=BN<=256, K=16) issued by one elected thread, operands staged
// by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-stage mbarrier ring, accumulator in TMEM,
// epilogue warps read it back with tcgen05.ld and fuse bias / ReLU / ReLU-mask / bf16 conversion /
// transposed copy / fp32 RED accumulation.  All three layer GEMMs are expressed in this one form by
// keeping K-major shadows (DESIGN.md §bf16 mode):
//   forward : A = H_{l-1} [rows,K]      B = W_l   [N,K]      -> H_l   (+ H_l^T)
//   dgrad   : A = dZ_l    [rows,N]      B = W_l^T [K,N]      -> dZ_{l-1} * [H_{l-1}>0] (+ transposed)
//   wgrad   : A = dZ_l^T  [N,rows]      B = H_{l-1}^T [K,rows], split over rows -> RED.ADD into dW_l (fp32)
// Replaces nn.Linear fwd/bwd of d3rlpy/models/torch/encoders.py:265-275 for every ensemble member
// (q_functions/ensemble_q_function.py:144-146,168-170) in one launch.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace d3b {
namespace umma {

constexpr int BM = 128;       // UMMA_M
constexpr int BK = 64;        // one 128-byte swizzle atom of bf16 along K
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct Params {
  int M, N, K;            // logical sizes (C is M x N, reduction K)
  int BN;                 // tile N (32/64/128/256)
  int lg_bn;              // log2(BN)
  int stages;             // operand ring depth (3 when a mask tile needs the space, else 4)
  int splits, kb_per_split;
  int csplit;             // > 1: the `splits` CTAs of a tile are a thread-block cluster; partial tiles are reduced through
                          // distributed shared memory and the full epilogue runs on the sum (latency configuration)
  int part_off;           // byte offset of the fp32 partial tile inside the operand stages (csplit > 1)
  int a_shared, b_shared; // operand shared by all members -> member coordinate 0
  int mn_major;           // 1: both operands are MN-major in memory (A [K][M], B [K][N]) — the weight-gradient form
  // epilogue
  const float* bias; long long sBias; int relu;
  const __nv_bfloat16* mask; long long ldmask, sMask;  // keep value where mask > 0
  __nv_bfloat16* out_bf16; long long ldo, sO;
  __nv_bfloat16* outT_bf16; long long ldt, sT;          // transposed copy [N][M]
  float* out_f32; long long ldf, sF; int atomic;        // fp32 store or RED.ADD
  long long* dbg;                                       // optional: per-CTA phase timestamps (8 slots)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a pipeline bug becomes a trapped launch error instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100): 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)0 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// MN-major, SWIZZLE_128B: atoms of 64 MN-elements (128 B) x 8 K-rows (1024 B); consecutive 8-row groups along K
// are SBO = 1024 B apart, consecutive 64-element groups along MN are LBO = one TMA box (BK rows x 128 B) apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((BK * 128) >> 4) << 16;             // leading byte offset: next 64-wide MN group
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: next 8 K-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// thread-block cluster helpers (cluster split-K, see gemm_tf32x3.cu for the measurements behind the relaxed arrive)
__device__ __forceinline__ void cluster_sync() {
  asm volatile(
      "fence.acq_rel.cta;\n\t"
      "barrier.cluster.arrive.relaxed.aligned;\n\t"
      "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(ra));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// warps 0-7: epilogue (TMEM lanes 32(w%4).., column half w/4), warp 8: TMA producer, warp 9: TMEM alloc + MMA issuer
constexpr int EPI_THREADS = 256;
constexpr int NTHREADS = EPI_THREADS + 64;

__device__ __forceinline__ void gemm_body(const CUtensorMap* tmA_p, const CUtensorMap* tmB_p, const Params& p,
                                          const int bx, const int by, const int bz, const int cta_linear) {
  const CUtensorMap& tmA = *tmA_p;
  const CUtensorMap& tmB = *tmB_p;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  const int b_stage_bytes = (p.mn_major && p.BN < 64 ? 64 : p.BN) * BK * 2;
  uint8_t* smA = smem;
  const int STAGES = p.stages;
  uint8_t* smB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(smB + STAGES * b_stage_bytes);  // 256 B reserved: barriers + TMEM slot
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + 8 * cta_linear : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const int m0 = bx * BM, n0 = by * p.BN;
  const int e = bz / p.splits, split = bz % p.splits;
  const int num_kb = (p.K + BK - 1) / BK;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
  const uint32_t tmem_cols = p.BN < 32 ? 32 : p.BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();  // setup done

  if (warp == 8) {
    if (lane == 0) {
      const int ea = p.a_shared ? 0 : e, eb = p.b_shared ? 0 : e;
      const uint32_t bytes = A_STAGE_BYTES + b_stage_bytes;
      for (int kb = kb_begin, i = 0; kb < kb_end; ++kb, ++i) {
        int s = i % STAGES;
        uint32_t phase = (i / STAGES) & 1;
        mbar_wait(empty + s, phase ^ 1);
        mbar_expect_tx(full + s, bytes);
        if (p.mn_major) {
          // boxes of {64 MN-elements, BK reduction rows}: coordinates (mn, k, member)
          for (int j = 0; j < BM / 64; ++j)
            tma_load_3d(smA + s * A_STAGE_BYTES + j * (BK * 128), &tmA, full + s, m0 + 64 * j, kb * BK, ea);
          for (int j = 0; j < (p.BN + 63) / 64; ++j)
            tma_load_3d(smB + s * b_stage_bytes + j * (BK * 128), &tmB, full + s, n0 + 64 * j, kb * BK, eb);
        } else {
          tma_load_3d(smA + s * A_STAGE_BYTES, &tmA, full + s, kb * BK, m0, ea);
          tma_load_3d(smB + s * b_stage_bytes, &tmB, full + s, kb * BK, n0, eb);
        }
      }
      if (dbg) dbg[2] = clock64();  // all TMA issued
    }
    if (p.csplit > 1) { __syncwarp(); cluster_sync(); }   // partial tiles staged (epilogue warps arrive below)
  } else if (warp == 9) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24) | (p.mn_major ? ((1u << 15) | (1u << 16)) : 0u);
      for (int kb = kb_begin, i = 0; kb < kb_end; ++kb, ++i) {
        int s = i % STAGES;
        uint32_t phase = (i / STAGES) & 1;
        mbar_wait(full + s, phase);
        if (dbg && i == 0) dbg[3] = clock64();  // first stage landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.mn_major) {
          uint64_t adesc = make_desc_mn(smem_u32(smA + s * A_STAGE_BYTES));
          uint64_t bdesc = make_desc_mn(smem_u32(smB + s * b_stage_bytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 reduction rows (16 x 128 B = 2048 B): +128 in the (>>4) address field
            mma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
        } else {
          uint64_t adesc = make_desc(smem_u32(smA + s * A_STAGE_BYTES));
          uint64_t bdesc = make_desc(smem_u32(smB + s * b_stage_bytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (>>4) address field
            mma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
        }
        mma_commit(empty + s);  // frees the smem stage when these MMAs retire
      }
      mma_commit(tmem_full);    // accumulator complete
      if (dbg) dbg[4] = clock64();  // all MMAs issued
    }
    if (p.csplit > 1) { __syncwarp(); cluster_sync(); }
  } else {
    // ---- epilogue (warps 0-7, 256 threads).  Phase 0 (overlaps the mainloop): stage bias and the ReLU
    // mask tile in shared memory with coalesced loads.  Phase 1: thread owns tile row 32*warp+lane, reads
    // the accumulator from TMEM, applies bias/ReLU/mask and writes the tile (row-major bf16, transposed
    // bf16 or fp32) into the now-free operand stages.  Phase 2: fully coalesced 16-byte global stores / REDs.
    const int t = threadIdx.x;  // 0..EPI_THREADS-1
    const int BN = p.BN;
    const int S = p.csplit > 1 ? p.csplit : 1;
    const int r_lo = S > 1 ? (int)cluster_ctarank() * (BM / S) : 0;   // tile rows this CTA finishes and stores
    const int r_hi = S > 1 ? r_lo + BM / S : BM;
    const int ldc = BN + 8, ldts = BM + 8, ldfs = BN + 4;
    float* bias_s = (float*)(smem + p.stages * (A_STAGE_BYTES + b_stage_bytes) + 256);
    __nv_bfloat16* mask_s = (__nv_bfloat16*)((uint8_t*)bias_s + 1024);
    __nv_bfloat16* c_s = (__nv_bfloat16*)smem;
    __nv_bfloat16* t_s = c_s + BM * ldc;
    float* f_s = (float*)smem;
    const int lg_nvec = p.lg_bn - 3;           // BN/8 bf16 vectors per tile row (BN is a power of two)
    const int nvec_mask = (1 << lg_nvec) - 1;
    if (p.bias) {
      const float* bias = p.bias + (long long)e * p.sBias;
      for (int j = t; j < BN; j += EPI_THREADS) bias_s[j] = (n0 + j < p.N) ? __ldg(bias + n0 + j) : 0.f;
    }
    if (p.mask) {
      // coalesced 16-byte loads, 8 in flight per thread
      const __nv_bfloat16* mk = p.mask + (long long)e * p.sMask;
      const bool vec_ok = ((p.ldmask & 7) == 0) && ((((uintptr_t)mk) & 15) == 0);
      const int total = BM << lg_nvec;
      for (int v0 = t; v0 < total; v0 += EPI_THREADS * 8) {
        uint4 val[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int v = v0 + u * EPI_THREADS;
          int r = v >> lg_nvec, cv = (v & nvec_mask) << 3;
          int m = m0 + r, n = n0 + cv;
          val[u] = make_uint4(0, 0, 0, 0);
          if (v < total && m < p.M) {
            const __nv_bfloat16* src = mk + (long long)m * p.ldmask + n;
            if (vec_ok && n + 7 < p.N) {
              val[u] = __ldg((const uint4*)src);
            } else {
              __nv_bfloat16 tmp[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) tmp[j] = (n + j < p.N) ? src[j] : __float2bfloat16_rn(0.f);
              val[u] = *reinterpret_cast<uint4*>(tmp);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int v = v0 + u * EPI_THREADS;
          if (v < total) *reinterpret_cast<uint4*>(mask_s + (v >> lg_nvec) * ldc + ((v & nvec_mask) << 3)) = val[u];
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (kb_end > kb_begin) mbar_wait(tmem_full, 0);
    if (dbg && threadIdx.x == 0) dbg[5] = clock64();  // accumulator ready
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // thread owns tile row 32*(warp%4)+lane and the column half warp/4 (BN=32: a single half)
    const int quarter = warp & 3;
    const int half_cols = BN >= 64 ? (BN >> 1) : BN;
    const int c_begin = (warp >> 2) * half_cols;
    const int c_end = (BN >= 64 || warp < 4) ? c_begin + half_cols : c_begin;
    // 32 accumulator columns [c, c + 32) of tile row `row`: bias / ReLU / mask, then into the staging tiles
    auto finish = [&](int row, int c, const uint32_t (&v)[32]) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 bv = p.bias ? *reinterpret_cast<const float4*>(bias_s + c + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[j] = __uint_as_float(v[j]) + bv.x;
        f[j + 1] = __uint_as_float(v[j + 1]) + bv.y;
        f[j + 2] = __uint_as_float(v[j + 2]) + bv.z;
        f[j + 3] = __uint_as_float(v[j + 3]) + bv.w;
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (p.mask) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 mv = *reinterpret_cast<const uint4*>(mask_s + row * ldc + c + j);
          const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&mv);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (!(__bfloat162float(mb[q]) > 0.f)) f[j + q] = 0.f;
        }
      }
      if (p.out_f32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(f_s + row * ldfs + c + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      } else {
        if (p.out_bf16) {
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(c_s + row * ldc + c + j) =
                make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]), pack_bf16(f[j + 4], f[j + 5]),
                           pack_bf16(f[j + 6], f[j + 7]));
        }
        if (p.outT_bf16) {
#pragma unroll
          for (int j = 0; j < 32; ++j) t_s[(c + j) * ldts + row] = __float2bfloat16_rn(f[j]);
        }
      }
    };
    if (S == 1) {
      const int row = quarter * 32 + lane;
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, v);
        finish(row, c, v);
      }
    } else {
      // cluster split-K: stage the raw partial tile, meet the other CTAs of the tile, then CTA q finishes rows
      // [q BM/S, (q+1) BM/S) on the sum of the S partial tiles (rank order: deterministic)
      float* part_s = (float*)(smem + p.part_off);
      {
        const int row = quarter * 32 + lane;
        for (int c = c_begin; c < c_end; c += 32) {
          uint32_t v[32];
          if (kb_end > kb_begin) {
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(part_s + row * ldfs + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      __syncwarp();
      cluster_sync();
      const int slab = BM / S, nchunk = BN >> 5;
      if (t < slab * nchunk) {
        const int row = r_lo + (t % slab), c = (t / slab) << 5;
        const uint32_t la = smem_u32(part_s + row * ldfs + c);
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        for (int r = 0; r < S; ++r) {
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = ld_dsmem_f4(la + 16 * j, (uint32_t)r);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[4 * j] += x[j].x; acc[4 * j + 1] += x[j].y; acc[4 * j + 2] += x[j].z; acc[4 * j + 3] += x[j].w;
          }
        }
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(acc[j]);
        finish(row, c, v);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- phase 2: coalesced global traffic (shift/mask indexing, 4 vectors in flight per thread)
    if (p.out_f32) {
      float* o = p.out_f32 + (long long)e * p.sF;
      const bool vec_ok = ((p.ldf & 3) == 0) && ((((uintptr_t)o) & 15) == 0);
      const int lg4 = p.lg_bn - 2, m4 = (1 << lg4) - 1;
      const int total = BM << lg4;
      for (int v0 = t; v0 < total; v0 += EPI_THREADS * 4) {
        float4 val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int v = v0 + u * EPI_THREADS;
          if (v < total) val[u] = *reinterpret_cast<const float4*>(f_s + (v >> lg4) * ldfs + ((v & m4) << 2));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int v = v0 + u * EPI_THREADS;
          int m = m0 + (v >> lg4), n = n0 + ((v & m4) << 2);
          if (v >= total || m >= p.M || n >= p.N || (v >> lg4) < r_lo || (v >> lg4) >= r_hi) continue;
          float* dst = o + (long long)m * p.ldf + n;
          if (vec_ok && n + 3 < p.N) {
            if (p.atomic)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(val[u].x), "f"(val[u].y),
                           "f"(val[u].z), "f"(val[u].w)
                           : "memory");
            else
              *reinterpret_cast<float4*>(dst) = val[u];
          } else {
            float tmp[4] = {val[u].x, val[u].y, val[u].z, val[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n + j < p.N) {
                if (p.atomic) atomicAdd(dst + j, tmp[j]);
                else dst[j] = tmp[j];
              }
            }
          }
        }
      }
    } else {
      if (p.out_bf16) {
        __nv_bfloat16* o = p.out_bf16 + (long long)e * p.sO;
        const bool vec_ok = ((p.ldo & 7) == 0) && ((((uintptr_t)o) & 15) == 0);
        const int total = BM << lg_nvec;
        for (int v0 = t; v0 < total; v0 += EPI_THREADS * 4) {
          uint4 val[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int v = v0 + u * EPI_THREADS;
            if (v < total) val[u] = *reinterpret_cast<const uint4*>(c_s + (v >> lg_nvec) * ldc + ((v & nvec_mask) << 3));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int v = v0 + u * EPI_THREADS;
            int m = m0 + (v >> lg_nvec), n = n0 + ((v & nvec_mask) << 3);
            if (v >= total || m >= p.M || n >= p.N || (v >> lg_nvec) < r_lo || (v >> lg_nvec) >= r_hi) continue;
            __nv_bfloat16* dst = o + (long long)m * p.ldo + n;
            if (vec_ok && n + 7 < p.N) {
              *reinterpret_cast<uint4*>(dst) = val[u];
            } else {
              const __nv_bfloat16* tb = reinterpret_cast<const __nv_bfloat16*>(&val[u]);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n + j < p.N) dst[j] = tb[j];
            }
          }
        }
      }
      if (p.outT_bf16) {
        __nv_bfloat16* o = p.outT_bf16 + (long long)e * p.sT;
        const bool vec_ok = ((p.ldt & 7) == 0) && ((((uintptr_t)o) & 15) == 0);
        const int total = BN << 4;  // BM/8 = 16 vectors of 8 rows per output row
        for (int v0 = t; v0 < total; v0 += EPI_THREADS * 4) {
          uint4 val[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int v = v0 + u * EPI_THREADS;
            if (v < total) val[u] = *reinterpret_cast<const uint4*>(t_s + (v >> 4) * ldts + ((v & 15) << 3));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int v = v0 + u * EPI_THREADS;
            int n = n0 + (v >> 4), m = m0 + ((v & 15) << 3);
            if (v >= total || n >= p.N || m >= p.M || ((v & 15) << 3) < r_lo || ((v & 15) << 3) >= r_hi) continue;
            __nv_bfloat16* dst = o + (long long)n * p.ldt + m;
            if (vec_ok && m + 7 < p.M) {
              *reinterpret_cast<uint4*>(dst) = val[u];
            } else {
              const __nv_bfloat16* tb = reinterpret_cast<const __nv_bfloat16*>(&val[u]);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (m + j < p.M) dst[j] = tb[j];
            }
          }
        }
      }
    }
  }
  if (dbg && threadIdx.x == 0) dbg[6] = clock64();  // epilogue (warp 0) done
  if (p.csplit > 1) { __syncwarp(); cluster_sync(); }   // peers may still be reading this CTA's partial tile
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[7] = clock64();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, Params p) {
  gemm_body(&tmA, &tmB, p, blockIdx.x, blockIdx.y, blockIdx.z,
            blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
}

// Several independent weight-gradient GEMMs (one per layer of a network) in ONE launch: blockIdx.z enumerates
// (problem, member, split); CTAs outside a problem's tile range exit immediately.
constexpr int MAX_BATCH = 5;
struct BatchMaps {
  CUtensorMap a[MAX_BATCH];
  CUtensorMap b[MAX_BATCH];
};
struct BatchParams {
  Params p[MAX_BATCH];
  int z_begin[MAX_BATCH + 1];
  int mtiles[MAX_BATCH], ntiles[MAX_BATCH];
  int n;
};
__global__ void __launch_bounds__(NTHREADS, 2) umma_gemm_batched_kernel(const __grid_constant__ BatchMaps maps,
                                                                   const __grid_constant__ BatchParams bp) {
  int l = 0;
  while (l + 1 < bp.n && (int)blockIdx.z >= bp.z_begin[l + 1]) ++l;
  if ((int)blockIdx.x >= bp.mtiles[l] || (int)blockIdx.y >= bp.ntiles[l]) return;
  gemm_body(&maps.a[l], &maps.b[l], bp.p[l], blockIdx.x, blockIdx.y, blockIdx.z - bp.z_begin[l],
            blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// K-major bf16 operand [members][rows][k] with leading dim ld and member stride `stride` (elements).
static int make_map(CUtensorMap* map, const void* base, int k, int rows, int members, long long ld, long long stride,
                    int box_rows, const char* what) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (((uintptr_t)base & 15) || (ld & 7) || (members > 1 && (stride & 7)))
    return set_err(D3B_ERR_ARG, "umma_gemm: %s must be 16-byte aligned with ld/stride multiples of 8 bf16", what);
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)(members < 1 ? 1 : members)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)((members > 1 ? stride : ld * (long long)rows) * 2)};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r);
  return D3B_OK;
}

// MN-major bf16 operand [members][k_rows][mn] (row-major, reduction index on rows) with leading dim ld.
static int make_map_mn(CUtensorMap* map, const void* base, int mn, int k_rows, int members, long long ld,
                       long long stride, const char* what) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (((uintptr_t)base & 15) || (ld & 7) || (members > 1 && (stride & 7)))
    return set_err(D3B_ERR_ARG, "umma_gemm_tn: %s must be 16-byte aligned with ld/stride multiples of 8 bf16", what);
  cuuint64_t dims[3] = {(cuuint64_t)mn, (cuuint64_t)k_rows, (cuuint64_t)(members < 1 ? 1 : members)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)((members > 1 ? stride : ld * (long long)k_rows) * 2)};
  cuuint32_t box[3] = {64, (cuuint32_t)BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r);
  return D3B_OK;
}

}  // namespace umma
}  // namespace d3b

using namespace d3b;
using namespace d3b::umma;

static long long* g_umma_dbg = nullptr;
static bool g_umma_cluster = true;
// profiling hook: 0 switches the cluster split-K latency configuration of d3b_umma_gemm off
extern "C" int d3b_umma_set_cluster(int enabled) {
  g_umma_cluster = enabled != 0;
  return D3B_OK;
}
// test/profiling hook: device buffer receiving 8 clock64() phase stamps per CTA of the next launches
extern "C" int d3b_umma_set_debug(void* device_buffer) {
  g_umma_dbg = (long long*)device_buffer;
  return D3B_OK;
}

// C[e] (M x N) = A[e] (M x K, ld lda) * B[e] (N x K, ldb)^T, bf16 operands.  stride_* == 0 => shared.
// Outputs are optional: bf16 row-major, bf16 transposed, fp32 (store or RED.ADD when `atomic`).
// splits > 1 partitions K (only meaningful with atomic fp32 output).
extern "C" int d3b_umma_gemm(const void* a, int64_t lda, int64_t stride_a, const void* b, int64_t ldb,
                             int64_t stride_b, int m, int n, int k, int members, int splits, const float* bias,
                             int64_t stride_bias, int relu, const void* mask, int64_t ld_mask, int64_t stride_mask,
                             void* out_bf16, int64_t ldo, int64_t stride_o, void* out_t_bf16, int64_t ldt,
                             int64_t stride_t, float* out_f32, int64_t ldf, int64_t stride_f, int atomic,
                             void* stream) {
  D3B_REQUIRE(m >= 0 && n > 0 && k > 0 && members > 0, "umma_gemm: bad sizes");
  if (m == 0) return D3B_OK;
  D3B_REQUIRE(a && b, "umma_gemm: null operand");
  D3B_REQUIRE(out_bf16 || out_t_bf16 || out_f32, "umma_gemm: no output requested");
  D3B_REQUIRE(!(out_f32 && (out_bf16 || out_t_bf16)), "umma_gemm: fp32 and bf16 outputs are mutually exclusive");
  D3B_REQUIRE(splits >= 1 && (splits == 1 || (atomic && out_f32 && !out_bf16 && !out_t_bf16 && !bias && !relu)),
              "umma_gemm: split-K needs a pure fp32 RED epilogue");
  int BN = n > 128 ? 256 : (n > 64 ? 128 : (n > 32 ? 64 : 32));
  // Few tiles: halve the tile so that twice as many CTAs exist and two of them fit on one SM
  // (~100 KB shared memory, 128 TMEM columns each) — one CTA's epilogue overlaps the other's mainloop.
  if (BN == 256 && (long long)ceil_div(m, BM) * ceil_div(n, 256) * members * splits < 2LL * kNumSM) BN = 128;
  int num_kb = ceil_div(k, BK);
  // Long reductions over few tiles (the fc layer of the pixel encoder: 32 x 512 x 3136): the launch is cut into
  // 128 x 64 tiles and each tile's reduction is split over a thread-block cluster of up to 8 CTAs (>= 6 K blocks of
  // 64 each); see gemm_body.  Measured (profiles/r2/umma_small.py): 21.3 -> 13.3 us for that layer; for K <= 752 the
  // TMA pipeline already streams the reduction in 5-9 us and the cluster epilogue costs more than it saves, so
  // shorter reductions keep the single-CTA tiles.
  int csplit = 1;
  if (splits == 1 && !atomic && g_umma_cluster) {
    long long tiles = (long long)ceil_div(m, BM) * ceil_div(n, BN) * members;
    if (tiles * 2 <= kNumSM && num_kb >= 16) {
      if (BN > 64) BN = 64;
      tiles = (long long)ceil_div(m, BM) * ceil_div(n, BN) * members;
      while (csplit * 2 <= 8 && tiles * csplit * 2 <= kNumSM && num_kb >= csplit * 2 * 6) csplit *= 2;
    }
  }
  Params p{};
  p.M = m; p.N = n; p.K = k; p.BN = BN;
  p.lg_bn = BN == 256 ? 8 : (BN == 128 ? 7 : (BN == 64 ? 6 : 5));
  if (csplit > 1) {
    p.splits = csplit;
    p.kb_per_split = ceil_div(num_kb, csplit);
  } else {
    if (splits > num_kb) splits = num_kb;
    p.kb_per_split = ceil_div(num_kb, splits);
    p.splits = ceil_div(num_kb, p.kb_per_split);
  }
  p.csplit = csplit;
  p.a_shared = stride_a == 0; p.b_shared = stride_b == 0;
  p.bias = bias; p.sBias = stride_bias; p.relu = relu;
  p.mask = (const __nv_bfloat16*)mask; p.ldmask = ld_mask; p.sMask = stride_mask;
  p.out_bf16 = (__nv_bfloat16*)out_bf16; p.ldo = ldo; p.sO = stride_o;
  p.outT_bf16 = (__nv_bfloat16*)out_t_bf16; p.ldt = ldt; p.sT = stride_t;
  p.out_f32 = out_f32; p.ldf = ldf; p.sF = stride_f; p.atomic = atomic;
  p.dbg = g_umma_dbg;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, a, k, m, p.a_shared ? 1 : members, lda, stride_a, BM, "A");
  if (rc) return rc;
  rc = make_map(&tmB, b, k, n, p.b_shared ? 1 : members, ldb, stride_b, BN, "B");
  if (rc) return rc;
  // epilogue staging needs: bf16 tile (+ transposed tile) or fp32 tile inside the operand stages, and
  // bias (1 KB) + mask tile after them
  size_t stage_bytes = (size_t)A_STAGE_BYTES + (size_t)BN * BK * 2;
  size_t need_epi = out_f32 ? (size_t)BM * (BN + 4) * 4
                            : (size_t)BM * (BN + 8) * 2 + (out_t_bf16 ? (size_t)BN * (BM + 8) * 2 : 0);
  if (csplit > 1) {  // the fp32 partial tile sits behind the output staging tiles
    p.part_off = (int)((need_epi + 127) / 128 * 128);
    need_epi = (size_t)p.part_off + (size_t)BM * (BN + 4) * 4;
  }
  size_t tail = 256 + 1024 + (mask ? (size_t)BM * (BN + 8) * 2 : 0);
  int stages = BN == 256 ? MAX_STAGES : 3;
  if (mask && !out_t_bf16 && BN <= 128) stages = 2;  // keeps two CTAs per SM with the mask tile resident
  while (stages > 2 && 1024 + stages * stage_bytes + tail > 227 * 1024) --stages;
  while ((size_t)stages * stage_bytes < need_epi) ++stages;  // tiny tiles: stages are smaller than the C tile
  D3B_REQUIRE(stages <= MAX_STAGES && 1024 + stages * stage_bytes + tail <= 227 * 1024,
              "umma_gemm: tile does not fit shared memory (BN=%d)", BN);
  p.stages = stages;
  size_t smem = 1024 + stages * stage_bytes + tail;
  static bool attr_set = false;
  if (!attr_set) {
    D3B_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(ceil_div(m, BM), ceil_div(n, BN), members * p.splits);
  launch_pdl_cluster(umma_gemm_kernel, grid, dim3(NTHREADS), smem, (cudaStream_t)stream, (unsigned)csplit, tmA, tmB, p);
  return check_launch("umma_gemm");
}

// Weight-gradient form: C[e] (m x n) (+)= A[e]^T B[e] with A [k][m] and B [k][n] row-major bf16 (the saved
// dZ_l and H_{l-1} of a layer, reduction over the k = minibatch rows).  Both operands are fed to
// tcgen05.mma as MN-major tiles straight from those row-major tensors: no transposed copies.
extern "C" int d3b_umma_gemm_tn(const void* a, int64_t lda, int64_t stride_a, const void* b, int64_t ldb,
                                int64_t stride_b, int m, int n, int k, int members, int splits, float* out_f32,
                                int64_t ldf, int64_t stride_f, int atomic, void* stream) {
  D3B_REQUIRE(m > 0 && n > 0 && k >= 0 && members > 0, "umma_gemm_tn: bad sizes");
  if (k == 0) return D3B_OK;
  D3B_REQUIRE(a && b && out_f32, "umma_gemm_tn: null pointer");
  D3B_REQUIRE(splits >= 1 && (splits == 1 || atomic), "umma_gemm_tn: split-K needs the RED epilogue");
  int BN = n > 128 ? 256 : (n > 64 ? 128 : (n > 32 ? 64 : 32));
  if (BN == 256 && (long long)ceil_div(m, BM) * ceil_div(n, 256) * members * splits < 2LL * kNumSM) BN = 128;
  Params p{};
  p.M = m; p.N = n; p.K = k; p.BN = BN; p.mn_major = 1;
  p.lg_bn = BN == 256 ? 8 : (BN == 128 ? 7 : (BN == 64 ? 6 : 5));
  int num_kb = ceil_div(k, BK);
  if (splits > num_kb) splits = num_kb;
  p.kb_per_split = ceil_div(num_kb, splits);
  p.splits = ceil_div(num_kb, p.kb_per_split);
  p.a_shared = stride_a == 0; p.b_shared = stride_b == 0;
  p.out_f32 = out_f32; p.ldf = ldf; p.sF = stride_f; p.atomic = atomic;
  p.dbg = g_umma_dbg;
  CUtensorMap tmA, tmB;
  int rc = make_map_mn(&tmA, a, m, k, p.a_shared ? 1 : members, lda, stride_a, "A");
  if (rc) return rc;
  rc = make_map_mn(&tmB, b, n, k, p.b_shared ? 1 : members, ldb, stride_b, "B");
  if (rc) return rc;
  size_t b_stage = (size_t)(BN < 64 ? 64 : BN) * BK * 2;
  size_t stage_bytes = (size_t)A_STAGE_BYTES + b_stage;
  size_t need_epi = (size_t)BM * (BN + 4) * 4;
  size_t tail = 256 + 1024;
  int stages = BN == 256 ? MAX_STAGES : 3;
  while (stages > 2 && 1024 + stages * stage_bytes + tail > 227 * 1024) --stages;
  while ((size_t)stages * stage_bytes < need_epi) ++stages;
  D3B_REQUIRE(stages <= MAX_STAGES && 1024 + stages * stage_bytes + tail <= 227 * 1024,
              "umma_gemm_tn: tile does not fit shared memory (BN=%d)", BN);
  p.stages = stages;
  size_t smem = 1024 + stages * stage_bytes + tail;
  static bool attr_set = false;
  if (!attr_set) {
    D3B_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(ceil_div(m, BM), ceil_div(n, BN), members * p.splits);
  launch_pdl(umma_gemm_kernel, grid, dim3(NTHREADS), smem, (cudaStream_t)stream, tmA, tmB, p);
  return check_launch("umma_gemm_tn");
}

// Batched weight-gradient form: problem i computes c_i[e] (m_i x n_i) += a_i[e]^T b_i[e] over k rows (same k and
// member count for all problems: the layers of one network).  All arrays are host arrays of length n_problems;
// stride_b_host[i] == 0 marks an operand shared by the members (the network input).
extern "C" int d3b_umma_gemm_tn_batched(int n_problems, const void* const* a_host, const int64_t* lda_host,
                                        const int64_t* stride_a_host, const void* const* b_host,
                                        const int64_t* ldb_host, const int64_t* stride_b_host, const int* m_host,
                                        const int* n_host, int k, int members, float* const* out_host,
                                        const int64_t* ldf_host, int64_t stride_f, void* stream) {
  D3B_REQUIRE(n_problems >= 1 && n_problems <= MAX_BATCH && k >= 0 && members >= 1, "umma_gemm_tn_batched: bad sizes");
  if (k == 0) return D3B_OK;
  D3B_REQUIRE(a_host && lda_host && stride_a_host && b_host && ldb_host && stride_b_host && m_host && n_host &&
                  out_host && ldf_host,
              "umma_gemm_tn_batched: null pointer");
  static BatchMaps maps;   // host staging (launches are serialised by the caller's stream-ordered use)
  static BatchParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.n = n_problems;
  int num_kb = ceil_div(k, BK);
  int gx = 1, gy = 1, z = 0;
  size_t smem = 0;
  // total CTA budget: ~2 CTAs per SM over all problems
  for (int i = 0; i < n_problems; ++i) {
    int m = m_host[i], n = n_host[i];
    D3B_REQUIRE(m > 0 && n > 0 && a_host[i] && b_host[i] && out_host[i], "umma_gemm_tn_batched: bad problem");
    int BN = n > 64 ? 128 : (n > 32 ? 64 : 32);
    Params& p = bp.p[i];
    p.M = m; p.N = n; p.K = k; p.BN = BN; p.mn_major = 1;
    p.lg_bn = BN == 128 ? 7 : (BN == 64 ? 6 : 5);
    int tiles = ceil_div(m, BM) * ceil_div(n, BN) * members;
    int splits = (2 * kNumSM) / (tiles * n_problems);
    if (splits < 1) splits = 1;
    if (splits > num_kb) splits = num_kb;
    p.kb_per_split = ceil_div(num_kb, splits);
    p.splits = ceil_div(num_kb, p.kb_per_split);
    p.a_shared = stride_a_host[i] == 0; p.b_shared = stride_b_host[i] == 0;
    p.out_f32 = out_host[i]; p.ldf = ldf_host[i]; p.sF = stride_f; p.atomic = 1;
    p.dbg = nullptr;
    int rc = make_map_mn(&maps.a[i], a_host[i], m, k, p.a_shared ? 1 : members, lda_host[i], stride_a_host[i], "A");
    if (rc) return rc;
    rc = make_map_mn(&maps.b[i], b_host[i], n, k, p.b_shared ? 1 : members, ldb_host[i], stride_b_host[i], "B");
    if (rc) return rc;
    size_t b_stage = (size_t)(BN < 64 ? 64 : BN) * BK * 2;
    size_t stage_bytes = (size_t)A_STAGE_BYTES + b_stage;
    size_t need_epi = (size_t)BM * (BN + 4) * 4;
    int stages = 3;
    while ((size_t)stages * stage_bytes < need_epi) ++stages;
    D3B_REQUIRE(stages <= MAX_STAGES, "umma_gemm_tn_batched: tile does not fit shared memory");
    p.stages = stages;
    size_t need = 1024 + stages * stage_bytes + 256 + 1024;
    if (need > smem) smem = need;
    bp.mtiles[i] = ceil_div(m, BM); bp.ntiles[i] = ceil_div(n, BN);
    if (bp.mtiles[i] > gx) gx = bp.mtiles[i];
    if (bp.ntiles[i] > gy) gy = bp.ntiles[i];
    bp.z_begin[i] = z;
    z += members * p.splits;
  }
  bp.z_begin[n_problems] = z;
  for (int i = n_problems; i < MAX_BATCH; ++i) { maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; }
  static bool attr_set = false;
  if (!attr_set) {
    D3B_CUDA(cudaFuncSetAttribute(umma_gemm_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  launch_pdl(umma_gemm_batched_kernel, dim3(gx, gy, z), dim3(NTHREADS), smem, (cudaStream_t)stream, maps, bp);
  return check_launch("umma_gemm_tn_batched");
}
