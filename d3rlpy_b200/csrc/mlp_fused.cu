// K2 (bf16 mode, fused): the whole ReLU-MLP trunk + narrow head of one network, for every ensemble member,
// in ONE persistent launch.
//
//   unit = (128-row tile, member).  Per unit the layers are chained on-chip:
//     layer 0 : A = x tile (TMA, zero-filled to 64-column K blocks)        B = W_0 K blocks (TMA ring)
//     layer l : A = relu(H_{l-1}) written by the epilogue warps straight into the SWIZZLE_128B K-major
//               operand buffer (never leaves shared memory)                 B = W_l K blocks (TMA ring)
//     tcgen05.mma M128 x N_l x K16, fp32 accumulators ping-ponging between two 256-column TMEM buffers.
//   Epilogue (8 warps, thread = accumulator row): tcgen05.ld -> +bias -> ReLU -> bf16 -> operand buffer;
//   the same buffer is TMA-stored to H_l in global memory when the backward pass will need it; on the last
//   layer the head (N <= 32 outputs: Q value, mu|logstd, action, VAE heads) is reduced per row from the
//   registers and written as fp32 (optionally through tanh).
// Replaces, per call: VectorEncoder[WithAction].forward (d3rlpy/models/torch/encoders.py:265-339), the Python
// loop over members of EnsembleContinuous/DiscreteQFunction (q_functions/ensemble_q_function.py:141-175) and the
// `_fc` / `_mu` / `_logstd` heads (mean_q_function.py:21,69; policies.py:55,92,153-158; imitators.py:45-54).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace d3b {
namespace fused {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int MAX_LAYERS = 4;
constexpr int A_KB_BYTES = BM * BK * 2;      // one 64-column K block of the operand buffer: 16 KB
constexpr int MAX_KB = 4;                    // layer widths <= 256
constexpr int MAXW = MAX_KB * BK;            // 256
constexpr int W_STAGE_BYTES = MAXW * BK * 2; // 32 KB
constexpr int W_STAGES_MAX = 4;  // forward weight ring: 4 stages when the head staging is small, else 3
constexpr int EPI_THREADS = 256, NTHREADS = EPI_THREADS + 64;

struct Maps {
  CUtensorMap x;
  CUtensorMap w[MAX_LAYERS];
  CUtensorMap h[MAX_LAYERS];
};

struct FwdParams {
  int rows, members, tiles, n_layers;
  int K[MAX_LAYERS], N[MAX_LAYERS];
  const float* bias[MAX_LAYERS];
  long long bias_stride;
  int x_shared, save_mask;
  int h_blocked;  // bit l: maps.h[l] is a column-blocked 4-D map (one store per tile), else one store per K block
  int save_rows;  // activations are stored only for tiles starting below this row (rows: all)
  const float* head_w;
  const float* head_b;
  long long head_stride;
  int n_head, head_tanh;
  float* head_out;  // [members][rows][n_head]
  long long* dbg;   // optional: 16 clock stamps per CTA (profiles/fused_phase_probe.py)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 4-D forms: dimension 2 enumerates the 64-column blocks of a row-major matrix (stride 128 B), so that ONE operation
// moves a whole [rows x 64*nblk] tile as [block][row][64] -- a TMA operation costs ~550 cycles of engine time whatever its
// size (profiles/r1_ubench.md), so the number of operations, not the bytes, is what the kernels below minimise
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// K-major SWIZZLE_128B operand: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
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
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// the wait names the destination registers as read-write operands so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// barrier slots
enum { B_WFULL = 0, B_WEMPTY = W_STAGES_MAX, B_XFULL = 2 * W_STAGES_MAX, B_AFREE, B_STDONE, B_ACCFULL, B_TEMPTY = B_ACCFULL + 2,
       B_ACTREADY = B_TEMPTY + 2, B_COUNT = B_ACTREADY + MAX_KB };

// EW = number of epilogue warps: 16 (four per TMEM lane quarter, one 16-column chunk of every K block each) hides the
// tcgen05.ld latency that bounds the epilogue when the per-thread state is small (NH <= 1), 8 otherwise.
template <int EW>
__device__ __forceinline__ void epi_sync_n() {
  asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
}

template <int NH, int EW>
__global__ void __launch_bounds__(EW * 32 + 96, 1) mlp_forward_kernel(const __grid_constant__ Maps maps, FwdParams p) {
  constexpr int EPI = EW * 32;          // epilogue threads
  constexpr int NSUB = EW / 4;          // column sub-ranges per K block
  constexpr int SUBW = BK / NSUB;       // columns of a K block owned by one sub-range
  constexpr int W_STAGES = NH <= 16 ? 4 : 3;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* smA = smem;
  uint8_t* smW = smem + MAX_KB * A_KB_BYTES;
  float* head_w_s = (float*)(smW + W_STAGES * W_STAGE_BYTES);  // [NH][MAXW]
  float* bias_s = head_w_s + NH * MAXW;                        // [MAX_LAYERS][MAXW]
  float* head_part = bias_s + MAX_LAYERS * MAXW;               // [NSUB-1][128][NH]
  uint64_t* bars = (uint64_t*)(head_part + (NSUB - 1) * BM * NH);
  uint32_t* tmem_slot = (uint32_t*)(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.n_layers;
  const int units = p.tiles * p.members;

  if (threadIdx.x == 0) {
    // TEMPTY / ACTREADY: one arrival per epilogue warp (no block-wide barrier in the layer loop)
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bars + i, (i >= B_TEMPTY && i < B_ACTREADY + MAX_KB) ? EW : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == EW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  long long* dbg = p.dbg ? p.dbg + 16 * blockIdx.x : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == EW) {
    // ================= TMA producer: x tile, then the weight K blocks of every layer through the ring
    if (lane == 0) {
      uint32_t wi = 0, it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int e = u / p.tiles, m0 = (u % p.tiles) * BM;
        mbar_wait(bars + B_AFREE, (it & 1) ^ 1);
        const int nkb0 = (p.K[0] + BK - 1) / BK;
        mbar_expect_tx(bars + B_XFULL, nkb0 * A_KB_BYTES);
        for (int kb = 0; kb < nkb0; ++kb)
          tma_load_3d(smA + kb * A_KB_BYTES, &maps.x, bars + B_XFULL, kb * BK, m0, p.x_shared ? 0 : e);
        for (int l = 0; l < L; ++l) {
          const int nkb = (p.K[l] + BK - 1) / BK;
          for (int kb = 0; kb < nkb; ++kb, ++wi) {
            const uint32_t s = wi % W_STAGES, ph = (wi / W_STAGES) & 1;
            mbar_wait(bars + B_WEMPTY + s, ph ^ 1);
            mbar_expect_tx(bars + B_WFULL + s, p.N[l] * BK * 2);
            tma_load_3d(smW + s * W_STAGE_BYTES, &maps.w[l], bars + B_WFULL + s, kb * BK, 0, e);
          }
        }
      }
    }
  } else if (warp == EW + 1) {
    // ================= MMA issuer
    if (lane == 0) {
      uint32_t g = 0, wi = 0, it = 0;
      uint32_t act_par = 0;  // phase parity bit per K-block barrier
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int m0 = (u % p.tiles) * BM;
        for (int l = 0; l < L; ++l, ++g) {
          const uint32_t buf = g & 1;
          mbar_wait(bars + B_TEMPTY + buf, ((g >> 1) & 1) ^ 1);
          if (l == 0) mbar_wait(bars + B_XFULL, it & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (dbg && g == 1) dbg[8] = clock64();                      // layer-1 operands ready
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N[l] >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
          const int nkb = (p.K[l] + BK - 1) / BK;
          const int ksteps = (p.K[l] + UMMA_K - 1) / UMMA_K;
          const uint32_t d_tmem = tmem_base + buf * 256;
          for (int kb = 0; kb < nkb; ++kb, ++wi) {
            const uint32_t s = wi % W_STAGES, ph = (wi / W_STAGES) & 1;
            if (l > 0) {  // K block kb of the previous layer's output has been written by the epilogue warps
              mbar_wait(bars + B_ACTREADY + kb, (act_par >> kb) & 1);
              act_par ^= 1u << kb;
            }
            mbar_wait(bars + B_WFULL + s, ph);
            if (dbg && g == 1 && kb < 4) dbg[9 + kb] = clock64();     // weight K block kb of layer 1 landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = make_desc(smem_u32(smA + kb * A_KB_BYTES));
            const uint64_t bdesc = make_desc(smem_u32(smW + s * W_STAGE_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (kb * (BK / UMMA_K) + k < ksteps)
                mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(bars + B_WEMPTY + s);
          }
          mma_commit(bars + B_ACCFULL + buf);
          if (dbg && g == 1) dbg[13] = clock64();                     // all MMAs of layer 1 issued
          if (l == L - 1 && ((p.save_mask >> l) & 1) && m0 < p.save_rows) {
            // the stored last layer completes one phase of the K-block barriers that only the store thread waits for
            const int nkb_out = (p.N[l] + BK - 1) / BK;
            for (int kb = 0; kb < nkb_out; ++kb) act_par ^= 1u << kb;
          }
        }
      }
    }
  } else if (warp == EW + 2) {
    // ================= activation stores: K blocks of H_l -> global memory as soon as they are complete in shared memory
    if (lane == 0) {
      uint32_t act_par = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int e = u / p.tiles, m0 = (u % p.tiles) * BM;
        for (int l = 0; l < L; ++l) {
          const bool last = (l == L - 1);
          const int nkb_out = (p.N[l] + BK - 1) / BK;
          const bool store = ((p.save_mask >> l) & 1) && m0 < p.save_rows;
          if (!store) {
            // The blocks are still produced as the next layer's operand (one barrier phase each).  The phases must be
            // WAITED for, not just counted: a parity wait only distinguishes adjacent phases, so a thread that runs two
            // phases ahead of the barrier would take an older completion for the one it wants.
            if (!last)
              for (int kb = 0; kb < nkb_out; ++kb) {
                mbar_wait(bars + B_ACTREADY + kb, (act_par >> kb) & 1);
                act_par ^= 1u << kb;
              }
            continue;
          }
          const bool blocked = (p.h_blocked >> l) & 1;
          for (int kb = 0; kb < nkb_out; ++kb) {
            mbar_wait(bars + B_ACTREADY + kb, (act_par >> kb) & 1);
            act_par ^= 1u << kb;
            if (!blocked) {
              tma_store_4d(&maps.h[l], smA + kb * A_KB_BYTES, kb * BK, m0, 0, e);
              tma_store_commit();
            }
          }
          if (blocked) {  // the whole tile in one operation
            tma_store_4d(&maps.h[l], smA, 0, m0, 0, e);
            tma_store_commit();
          }
          tma_store_wait_read();  // the bulk stores have read the operand buffer: it may be overwritten / reloaded
          mbar_arrive(bars + (last ? B_AFREE : B_STDONE));
        }
      }
    }
  } else {
    // ================= epilogue warps: thread = accumulator row 32*(warp%4)+lane, column sub-range warp/4 of each K block
    const int t = threadIdx.x;
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t g = 0, sd = 0;  // sd: completed STDONE phases this thread has consumed
    int cur_member = -1;
    bool stores_pending = false;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int e = u / p.tiles, m0 = (u % p.tiles) * BM;
      if (e != cur_member) {
        epi_sync_n<EW>();  // nobody is still reading the previous member's constants
        for (int l = 0; l < L; ++l) {
          const float* b = p.bias[l] + (long long)e * p.bias_stride;
          for (int j = t; j < p.N[l]; j += EPI) bias_s[l * MAXW + j] = __ldg(b + j);
        }
        if (NH > 0 && p.n_head > 0 && t < MAXW) {
          const int feat = p.N[L - 1];
          const float* hw = p.head_w + (long long)e * p.head_stride;
          // NH x 256 staging: all loads of a thread are issued before the stores (NH global loads in flight)
          float tmp[NH > 0 ? NH : 1];
#pragma unroll
          for (int j = 0; j < NH; ++j) tmp[j] = (j < p.n_head && t < feat) ? __ldg(hw + (long long)j * feat + t) : 0.f;
#pragma unroll
          for (int j = 0; j < NH; ++j) head_w_s[t * NH + j] = tmp[j];  // [column][head]: one row of NH weights per column
        }
        epi_sync_n<EW>();
        cur_member = e;
      }
      for (int l = 0; l < L; ++l, ++g) {
        const uint32_t buf = g & 1;
        const bool last = (l == L - 1);
        const bool store = ((p.save_mask >> l) & 1) && m0 < p.save_rows;
        const bool writeA = !last || store;
        const int N = p.N[l];
        mbar_wait(bars + B_ACCFULL + buf, (g >> 1) & 1);
        if (dbg && t == 0 && g < 6) dbg[2 + 2 * g] = clock64();      // accumulator of layer g ready
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (last && !store && t == 0) mbar_arrive(bars + B_AFREE);  // operand buffer no longer needed by this unit
        if (stores_pending) {  // the previous layer's TMA stores have finished reading the operand buffer
          mbar_wait(bars + B_STDONE, sd & 1);
          ++sd;
          stores_pending = false;
        }
        const uint32_t taddr = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
        float acc[NH > 0 ? NH : 1];
#pragma unroll
        for (int j = 0; j < (NH > 0 ? NH : 1); ++j) acc[j] = 0.f;
        // The 64-column K blocks of the operand buffer are produced IN ORDER by all epilogue warps (column
        // sub-range `sub` of each block), so that the next layer's MMAs on block kb start while blocks kb+1.. are still in
        // the epilogue, and the block's TMA store is issued as soon as it is complete.
        const int nkb_out = (N + BK - 1) / BK;
        for (int kb = 0; kb < nkb_out; ++kb) {
          const int cb = kb * BK;
          const int rem = (N - cb) < BK ? (N - cb) : BK;
          const int lo = sub * SUBW, hi = (sub + 1) * SUBW;
          const int c_begin = cb + (lo < rem ? lo : rem);
          const int c_end = cb + (hi < rem ? hi : rem);
          for (int c = c_begin; c < c_end; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)c, v);
            uint32_t pk[8];
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + l * MAXW + c);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bb = b4[i >> 2];
              float f0 = fmaxf(__uint_as_float(v[i]) + bb.x, 0.f);
              float f1 = fmaxf(__uint_as_float(v[i + 1]) + bb.y, 0.f);
              float f2 = fmaxf(__uint_as_float(v[i + 2]) + bb.z, 0.f);
              float f3 = fmaxf(__uint_as_float(v[i + 3]) + bb.w, 0.f);
              pk[i >> 1] = pack_bf16(f0, f1);
              pk[(i >> 1) + 1] = pack_bf16(f2, f3);
            }
            if (writeA) {
              const int j0 = (c & 63) >> 3;
              uint8_t* base = smA + kb * A_KB_BYTES + row * 128;
              *reinterpret_cast<uint4*>(base + ((j0 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(base + (((j0 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            if (NH > 0 && last && p.n_head > 0) {
              float hv[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                hv[2 * i] = bf16_lo(pk[i]);
                hv[2 * i + 1] = bf16_hi(pk[i]);
              }
              // column-outer: the NH accumulators are independent FMA chains; weights of one column are contiguous
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float* wrow = head_w_s + (c + i) * NH;
                if constexpr (NH % 4 == 0) {
#pragma unroll
                  for (int j4 = 0; j4 < NH / 4; ++j4) {
                    float4 w = reinterpret_cast<const float4*>(wrow)[j4];
                    acc[4 * j4] = fmaf(hv[i], w.x, acc[4 * j4]);
                    acc[4 * j4 + 1] = fmaf(hv[i], w.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(hv[i], w.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(hv[i], w.w, acc[4 * j4 + 3]);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < NH; ++j) acc[j] = fmaf(hv[i], wrow[j], acc[j]);
                }
              }
            }
          }
          if (writeA) {
            // this warp's piece of block kb is in shared memory: hand it to the MMA issuer / the store thread
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + B_ACTREADY + kb);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_TEMPTY + buf);  // this warp has drained its part of the accumulator buffer
        if (dbg && t == 0 && g < 6) dbg[3 + 2 * g] = clock64();      // epilogue of layer g done
        stores_pending = store && !last;
        const bool head_here = NH > 0 && last && p.n_head > 0;
        if (head_here) {
          if (sub > 0) {
#pragma unroll
            for (int j = 0; j < NH; ++j) head_part[((sub - 1) * BM + row) * NH + j] = acc[j];
          }
          epi_sync_n<EW>();
        }
        if (NH > 0 && last && p.n_head > 0 && sub == 0 && m0 + row < p.rows) {
          const float* hb = p.head_b + (long long)e * p.head_stride;
          float* o = p.head_out + ((long long)e * p.rows + m0 + row) * p.n_head;
#pragma unroll
          for (int j = 0; j < NH; ++j) {
            if (j < p.n_head) {
              float val = acc[j] + __ldg(hb + j);
#pragma unroll
              for (int s2 = 0; s2 < NSUB - 1; ++s2) val += head_part[(s2 * BM + row) * NH + j];
              o[j] = p.head_tanh ? tanhf(val) : val;
            }
          }
        }
      }
      if (NH > 0 && p.n_head > 0) epi_sync_n<EW>();  // head_part is rewritten by the next unit's last layer
    }
    if (dbg && t == 0) dbg[14] = clock64();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[15] = clock64();
  if (warp == EW + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ------------------------------------------------------------------------------------------ backward (data)
// One persistent launch for the whole data-gradient chain of a network (autograd of the forward above):
//   dZ_{L-1} = (d_head . W_head) * [H_{L-1} > 0]            computed by the epilogue warps (thread = row)
//   dZ_{l-1} = (dZ_l . W_l) * [H_{l-1} > 0]                 tcgen05.mma: A = dZ_l in the operand buffer,
//                                                           B = W_l row-major = MN-major tiles (no W^T copy)
//   dX[:, col0:col0+cols] = dZ_0 . W_0[:, col0:col0+cols]   optional (gradient w.r.t. the action input)
// and, when weight gradients are wanted: every dZ_l is TMA-stored for the weight-gradient GEMM, bias
// gradients (column sums of dZ_l) and the head's dW/db are reduced per tile and RED-added to the arena.
// H_l tiles (ReLU masks) are TMA-loaded into a second operand-layout buffer.
constexpr int BW_STAGES = 2;
enum { C_WFULL = 0, C_WEMPTY = BW_STAGES, C_MFULL = 2 * BW_STAGES, C_MEMPTY, C_ACCFULL, C_TEMPTY = C_ACCFULL + 2,
       C_ACTREADY = C_TEMPTY + 2, C_COUNT = C_ACTREADY + MAX_KB };

struct BwdMaps {
  CUtensorMap w[MAX_LAYERS];
  CUtensorMap h[MAX_LAYERS];
  CUtensorMap dz[MAX_LAYERS];
};

struct BwdParams {
  int rows, members, tiles, n_layers;
  int K[MAX_LAYERS], N[MAX_LAYERS];
  int w_blocked, h_blocked, dz_blocked;  // bit l: the layer's map is column-blocked (one TMA operation per tile / K block)
  const float* d_head;   // [members][rows][n_head]
  const float* head_w;   // [n_head][feat] per member
  long long head_stride;
  int n_head;
  int weight_grads;
  float* dbias[MAX_LAYERS];
  float* d_head_w;
  float* d_head_b;
  long long grad_stride;
  __nv_bfloat16* d_head_bf16;  // optional [members][rows][16] bf16 copy of d_head: the head's dW then comes from
                               // the batched weight-gradient GEMM instead of the in-kernel column pass
  float* dx;             // [members][rows][dx_cols] (lddx), may be null
  long long lddx, dx_stride;
  int dx_col0, dx_cols;
  long long* dbg;  // optional: 16 clock stamps per CTA (profiles/fused_phase_probe.py)
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((BK * 128) >> 4) << 16;  // next 64-wide MN group (one TMA box of 64 reduction rows)
  d |= (uint64_t)(1024 >> 4) << 32;        // next 8 reduction rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ const __nv_bfloat16* swz_ptr(const uint8_t* buf, int r, int c) {
  return reinterpret_cast<const __nv_bfloat16*>(buf + (c >> 6) * A_KB_BYTES + r * 128 +
                                                ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}

template <int NH>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_backward_kernel(const __grid_constant__ BwdMaps maps, BwdParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* smA = smem;                                   // dZ_l operand buffer (4 K blocks)
  uint8_t* smM = smem + MAX_KB * A_KB_BYTES;             // H_l tile (same layout): ReLU mask
  uint8_t* smW = smM + MAX_KB * A_KB_BYTES;              // weight ring
  float* head_w_s = (float*)(smW + BW_STAGES * W_STAGE_BYTES);  // [NH][MAXW]
  float* dhead_s = head_w_s + NH * MAXW;                        // [128][NH]
  float* colred_s = dhead_s + BM * NH;                          // [1 + NHW][4 row groups][MAXW] column-pass partials
  uint64_t* bars = (uint64_t*)(colred_s + (1 + (NH <= 2 ? NH : 0)) * 4 * MAXW);
  uint32_t* tmem_slot = (uint32_t*)(bars + C_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.n_layers;
  const int units = p.tiles * p.members;
  const bool has_dx = p.dx != nullptr;
  // TMA box starts must be 16-byte aligned: load from the 8-column boundary below dx_col0 and skip `dx_shift`
  // accumulator columns in the epilogue
  const int dx_shift = p.dx_col0 & 7;
  const int dx_w = has_dx ? ((dx_shift + p.dx_cols + 15) / 16) * 16 : 0;
  const int n_steps = (L - 1) + (has_dx ? 1 : 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < C_COUNT; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  long long* dbg = p.dbg ? p.dbg + 16 * blockIdx.x : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 8) {
    // ================= TMA producer: mask tiles H_l (l = L-1..0) interleaved with the weight blocks of each step
    if (lane == 0) {
      uint32_t wi = 0, mi = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int e = u / p.tiles, m0 = (u % p.tiles) * BM;
        for (int l = L - 1; l >= 0; --l) {
          // mask for dZ_l
          mbar_wait(bars + C_MEMPTY, (mi & 1) ^ 1);
          const int nkbm = (p.N[l] + BK - 1) / BK;
          mbar_expect_tx(bars + C_MFULL, nkbm * A_KB_BYTES);
          if ((p.h_blocked >> l) & 1) {
            tma_load_4d(smM, &maps.h[l], bars + C_MFULL, 0, m0, 0, e);
          } else {
            for (int kb = 0; kb < nkbm; ++kb)
              tma_load_4d(smM + kb * A_KB_BYTES, &maps.h[l], bars + C_MFULL, kb * BK, m0, 0, e);
          }
          ++mi;
          // weights of the step that consumes dZ_l: W_l (l >= 1), or the dx columns of W_0
          if (l == 0 && !has_dx) break;
          const int outw = l > 0 ? p.K[l] : dx_w;
          const int col0 = l > 0 ? 0 : (p.dx_col0 & ~7);
          const int nbox = (outw + 63) / 64;
          const int nkb = (p.N[l] + BK - 1) / BK;
          for (int kb = 0; kb < nkb; ++kb, ++wi) {
            const uint32_t s = wi % BW_STAGES, ph = (wi / BW_STAGES) & 1;
            mbar_wait(bars + C_WEMPTY + s, ph ^ 1);
            mbar_expect_tx(bars + C_WFULL + s, nbox * (BK * 128));
            if (l > 0 && ((p.w_blocked >> l) & 1)) {
              tma_load_4d(smW + s * W_STAGE_BYTES, &maps.w[l], bars + C_WFULL + s, 0, kb * BK, 0, e);
            } else {
              for (int j = 0; j < nbox; ++j)
                tma_load_4d(smW + s * W_STAGE_BYTES + j * (BK * 128), &maps.w[l], bars + C_WFULL + s, col0 + 64 * j,
                            kb * BK, 0, e);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer: step s consumes dZ_l (l = L-1-s; the last optional step is the dx step on dZ_0)
    if (lane == 0) {
      uint32_t g = 0, wi = 0;
      uint32_t act_cnt[MAX_KB] = {0, 0, 0, 0};
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        for (int s_ = 0; s_ < n_steps; ++s_, ++g) {
          const int l = L - 1 - s_;
          const int outw = l > 0 ? p.K[l] : dx_w;
          const uint32_t buf = g & 1;
          mbar_wait(bars + C_TEMPTY + buf, ((g >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(outw >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
          const int nkb = (p.N[l] + BK - 1) / BK;
          const int ksteps = (p.N[l] + UMMA_K - 1) / UMMA_K;
          const uint32_t d_tmem = tmem_base + buf * 256;
          for (int kb = 0; kb < nkb; ++kb, ++wi) {
            const uint32_t s = wi % BW_STAGES, ph = (wi / BW_STAGES) & 1;
            mbar_wait(bars + C_ACTREADY + kb, act_cnt[kb] & 1);  // K block kb of dZ_l is in the operand buffer
            ++act_cnt[kb];
            mbar_wait(bars + C_WFULL + s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = make_desc(smem_u32(smA + kb * A_KB_BYTES));
            const uint64_t bdesc = make_desc_mn(smem_u32(smW + s * W_STAGE_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (kb * (BK / UMMA_K) + k < ksteps)
                mma_bf16(d_tmem, adesc + 2 * k, bdesc + 128 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(bars + C_WEMPTY + s);
          }
          mma_commit(bars + C_ACCFULL + buf);
        }
      }
    }
  } else {
    // ================= epilogue warps
    const int t = threadIdx.x;
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    uint32_t g = 0, mi = 0;
    int cur_member = -1;
    bool stores_pending = false;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int e = u / p.tiles, m0 = (u % p.tiles) * BM;
      const int feat = p.N[L - 1];
      if (e != cur_member) {
        epi_sync();
        const float* hw = p.head_w + (long long)e * p.head_stride;
        float tmp[NH];
#pragma unroll
        for (int j = 0; j < NH; ++j) tmp[j] = (j < p.n_head && t < feat) ? __ldg(hw + (long long)j * feat + t) : 0.f;
#pragma unroll
        for (int j = 0; j < NH; ++j) head_w_s[j * MAXW + t] = tmp[j];
        cur_member = e;
      }
      // d_head row -> registers (+ shared copy for the column passes)
      float dh[NH];
      {
        const bool live = m0 + row < p.rows;
        const float* src = p.d_head + ((long long)e * p.rows + m0 + row) * p.n_head;
#pragma unroll
        for (int j = 0; j < NH; ++j) dh[j] = (live && j < p.n_head) ? __ldg(src + j) : 0.f;
        if (half == 0) {
#pragma unroll
          for (int j = 0; j < NH; ++j) dhead_s[row * NH + j] = dh[j];
          if constexpr (NH > 1) {
            if (p.d_head_bf16 && live) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(2 * j < NH ? dh[2 * j] : 0.f, 2 * j + 1 < NH ? dh[2 * j + 1] : 0.f);
              uint4* dst = reinterpret_cast<uint4*>(p.d_head_bf16 + ((long long)e * p.rows + m0 + row) * 16);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
      if (stores_pending) {
        if (t == 0) tma_store_wait_read();
        stores_pending = false;
      }
      epi_sync();  // head_w_s / dhead_s visible; previous unit's stores done reading the operand buffer
      if (dbg && t == 0 && u == (int)blockIdx.x) dbg[2] = clock64();  // head constants / d_head staged
      for (int l = L - 1; l >= 0; --l) {
        // ---- produce dZ_l into the operand buffer
        const int N = p.N[l];
        const bool top = (l == L - 1);
        uint32_t buf = 0;
        if (!top) {
          buf = g & 1;
          mbar_wait(bars + C_ACCFULL + buf, (g >> 1) & 1);  // (dZ_{l+1} W_{l+1}) ready; MMA finished reading the buffer
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (stores_pending) {
            if (t == 0) tma_store_wait_read();
            epi_sync();
            stores_pending = false;
          }
        }
        mbar_wait(bars + C_MFULL, mi & 1);
        ++mi;
        if (dbg && t == 0 && u == (int)blockIdx.x && L - 1 - l < 4) dbg[3 + 3 * (L - 1 - l)] = clock64();  // inputs of dZ_l ready
        const uint32_t taddr = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
        const bool mma_follows = (l > 0) || has_dx;
        const int nkb_out = (N + BK - 1) / BK;
        bool wide_top = false;
        if constexpr (NH > 2) {
          // Wide head (policy mu|logstd, VAE heads): dZ_top[r][c] = sum_j d_head[r][j] W_head[j][c] with thread = row is
          // bound by the broadcast reads of the 4*NH weights per 16 columns; with thread = 4 columns x 32 rows the NH x 4
          // weights stay in registers and each row costs NH/4 broadcast reads of d_head instead (FMA-bound).
          if (top) {
            wide_top = true;
            const int c4 = (t & 63) * 4, r0 = (t >> 6) * 32;
            if (c4 < N) {
              float w[NH][4];
#pragma unroll
              for (int j = 0; j < NH; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(head_w_s + j * MAXW + c4);
                w[j][0] = wv.x; w[j][1] = wv.y; w[j][2] = wv.z; w[j][3] = wv.w;
              }
              const int coff = (c4 >> 6) * A_KB_BYTES + (c4 & 7) * 2, chunk = (c4 & 63) >> 3;
#pragma unroll 2
              for (int r = r0; r < r0 + 32; ++r) {
                float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
                const float4* d4 = reinterpret_cast<const float4*>(dhead_s + r * NH);  // NH % 4 == 0 here, 16-B aligned
#pragma unroll
                for (int j4 = 0; j4 < NH / 4; ++j4) {
                  const float4 dv = d4[j4];
                  const float dj[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * j4 + jj;
                    f0 = fmaf(dj[jj], w[j][0], f0); f1 = fmaf(dj[jj], w[j][1], f1);
                    f2 = fmaf(dj[jj], w[j][2], f2); f3 = fmaf(dj[jj], w[j][3], f3);
                  }
                }
                const int off = coff + r * 128 + ((chunk ^ (r & 7)) << 4);
                const uint2 mv = *reinterpret_cast<const uint2*>(smM + off);
                const float a0 = bf16_lo(mv.x) > 0.f ? f0 : 0.f, a1 = bf16_hi(mv.x) > 0.f ? f1 : 0.f;
                const float a2 = bf16_lo(mv.y) > 0.f ? f2 : 0.f, a3 = bf16_hi(mv.y) > 0.f ? f3 : 0.f;
                *reinterpret_cast<uint2*>(smA + off) = make_uint2(pack_bf16(a0, a1), pack_bf16(a2, a3));
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            epi_sync();
            if (t == 0) {
              for (int kb = 0; kb < nkb_out; ++kb) {
                if (mma_follows) mbar_arrive(bars + C_ACTREADY + kb);
                if (p.weight_grads && !((p.dz_blocked >> l) & 1)) {
                  tma_store_4d(&maps.dz[l], smA + kb * A_KB_BYTES, kb * BK, m0, 0, e);
                  tma_store_commit();
                }
              }
            }
          }
        }
        // dZ_l is produced K block by K block (all 8 warps on one 64-column block at a time) so that the next
        // step's MMAs and the block's TMA store start while the remaining blocks are still being masked
        for (int kb = 0; kb < (wide_top ? 0 : nkb_out); ++kb) {
          const int cb = kb * BK;
          const int rem = (N - cb) < BK ? (N - cb) : BK;
          const int h0 = rem < 32 ? rem : 32;
          const int c_begin = half ? cb + h0 : cb;
          const int c_end = half ? cb + rem : cb + h0;
          for (int c = c_begin; c < c_end; c += 16) {
            float f[16];
            if (top) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = 0.f;
#pragma unroll
              for (int j = 0; j < NH; ++j) {
                const float4* w4 = reinterpret_cast<const float4*>(head_w_s + j * MAXW + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float4 w = w4[i];
                  f[4 * i] = fmaf(dh[j], w.x, f[4 * i]);
                  f[4 * i + 1] = fmaf(dh[j], w.y, f[4 * i + 1]);
                  f[4 * i + 2] = fmaf(dh[j], w.z, f[4 * i + 2]);
                  f[4 * i + 3] = fmaf(dh[j], w.w, f[4 * i + 3]);
                }
              }
            } else {
              uint32_t v[16];
              tmem_ld16(taddr + (uint32_t)c, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            }
            const int j0 = (c & 63) >> 3;
            const int off0 = kb * A_KB_BYTES + row * 128 + ((j0 ^ (row & 7)) << 4);
            const int off1 = kb * A_KB_BYTES + row * 128 + (((j0 + 1) ^ (row & 7)) << 4);
            uint4 m0v = *reinterpret_cast<const uint4*>(smM + off0);
            uint4 m1v = *reinterpret_cast<const uint4*>(smM + off1);
            const uint32_t mk[8] = {m0v.x, m0v.y, m0v.z, m0v.w, m1v.x, m1v.y, m1v.z, m1v.w};
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float a = bf16_lo(mk[i]) > 0.f ? f[2 * i] : 0.f;
              float b = bf16_hi(mk[i]) > 0.f ? f[2 * i + 1] : 0.f;
              pk[i] = pack_bf16(a, b);
            }
            *reinterpret_cast<uint4*>(smA + off0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(smA + off1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          epi_sync();
          if (t == 0) {
            if (mma_follows) mbar_arrive(bars + C_ACTREADY + kb);
            if (p.weight_grads && !((p.dz_blocked >> l) & 1)) {
              tma_store_4d(&maps.dz[l], smA + kb * A_KB_BYTES, cb, m0, 0, e);
              tma_store_commit();
            }
          }
        }
        if (p.weight_grads && ((p.dz_blocked >> l) & 1) && t == 0) {  // the whole dZ_l tile in one operation
          tma_store_4d(&maps.dz[l], smA, 0, m0, 0, e);
          tma_store_commit();
        }
        if (!top) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          epi_sync();
          if (t == 0) mbar_arrive(bars + C_TEMPTY + buf);  // every thread has drained this accumulator buffer
        }
        if (!top) ++g;
        if (dbg && t == 0 && u == (int)blockIdx.x && L - 1 - l < 4) dbg[4 + 3 * (L - 1 - l)] = clock64();  // dZ_l written
        stores_pending = p.weight_grads != 0;
        // ---- column passes over the tile, overlapping the MMA of the next step.  Thread = 4 consecutive columns x one
        // of 4 groups of 32 rows (a 4x shorter dependent chain than one thread per column); the four partial sums meet in
        // shared memory so that every column still costs ONE RED to the gradient arena per tile.
        if (p.weight_grads) {
          constexpr int NHW = NH <= 2 ? NH : 0;  // head dW in-kernel only for narrow heads (wide: GEMM path / fallback)
          const bool head_dw = top && !p.d_head_bf16;
          const int c4 = (t & 63) * 4, rg = t >> 6, r0 = rg * 32;
          if (c4 < N) {
            const int coff = (c4 >> 6) * A_KB_BYTES + (c4 & 7) * 2, chunk = (c4 & 63) >> 3;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
            for (int r = r0; r < r0 + 32; ++r) {
              const uint2 v = *reinterpret_cast<const uint2*>(smA + coff + r * 128 + ((chunk ^ (r & 7)) << 4));
              s0 += bf16_lo(v.x); s1 += bf16_hi(v.x); s2 += bf16_lo(v.y); s3 += bf16_hi(v.y);
            }
            *reinterpret_cast<float4*>(colred_s + rg * MAXW + c4) = make_float4(s0, s1, s2, s3);
            if constexpr (NHW > 0) {
              if (head_dw) {  // head dW[j][c] = sum_r d_head[r][j] * H[r][c]
                float a[NHW][4];
#pragma unroll
                for (int j = 0; j < NHW; ++j) a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
#pragma unroll 8
                for (int r = r0; r < r0 + 32; ++r) {
                  const uint2 v = *reinterpret_cast<const uint2*>(smM + coff + r * 128 + ((chunk ^ (r & 7)) << 4));
                  const float h0 = bf16_lo(v.x), h1 = bf16_hi(v.x), h2 = bf16_lo(v.y), h3 = bf16_hi(v.y);
#pragma unroll
                  for (int j = 0; j < NHW; ++j) {
                    const float d = dhead_s[r * NH + j];
                    a[j][0] = fmaf(d, h0, a[j][0]); a[j][1] = fmaf(d, h1, a[j][1]);
                    a[j][2] = fmaf(d, h2, a[j][2]); a[j][3] = fmaf(d, h3, a[j][3]);
                  }
                }
#pragma unroll
                for (int j = 0; j < NHW; ++j)
                  *reinterpret_cast<float4*>(colred_s + ((1 + j) * 4 + rg) * MAXW + c4) = make_float4(a[j][0], a[j][1], a[j][2], a[j][3]);
              }
            }
          }
          epi_sync();
          if (t < N) {
            atomicAdd(p.dbias[l] + (long long)e * p.grad_stride + t,
                      colred_s[t] + colred_s[MAXW + t] + colred_s[2 * MAXW + t] + colred_s[3 * MAXW + t]);
            if (head_dw) {
              if constexpr (NHW > 0) {
#pragma unroll
                for (int j = 0; j < NHW; ++j) {
                  const float* cr = colred_s + (1 + j) * 4 * MAXW + t;
                  if (j < p.n_head)
                    atomicAdd(p.d_head_w + (long long)e * p.grad_stride + (long long)j * feat + t,
                              cr[0] + cr[MAXW] + cr[2 * MAXW] + cr[3 * MAXW]);
                }
              } else {  // wide heads normally take the GEMM path (d_head_bf16); generic fallback, thread = column
                float a[NH];
#pragma unroll
                for (int j = 0; j < NH; ++j) a[j] = 0.f;
                for (int r = 0; r < BM; ++r) {
                  float hv = __bfloat162float(*swz_ptr(smM, r, t));
#pragma unroll
                  for (int j = 0; j < NH; ++j) a[j] = fmaf(dhead_s[r * NH + j], hv, a[j]);
                }
#pragma unroll
                for (int j = 0; j < NH; ++j)
                  if (j < p.n_head) atomicAdd(p.d_head_w + (long long)e * p.grad_stride + (long long)j * feat + t, a[j]);
              }
            }
          }
          if (top && t < p.n_head) {
            float s = 0.f;
            for (int r = 0; r < BM; ++r) s += dhead_s[r * NH + t];
            atomicAdd(p.d_head_b + (long long)e * p.grad_stride + t, s);
          }
        }
        // the mask tile has been consumed (by the element pass and the head column pass)
        epi_sync();
        if (dbg && t == 0 && u == (int)blockIdx.x && L - 1 - l < 4) dbg[5 + 3 * (L - 1 - l)] = clock64();  // column passes done
        if (t == 0) mbar_arrive(bars + C_MEMPTY);
      }
      if (has_dx) {
        const uint32_t buf = g & 1;
        mbar_wait(bars + C_ACCFULL + buf, (g >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
        if (half == 0) {
          for (int c = 0; c < dx_w; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)c, v);
            if (m0 + row < p.rows) {
              float* o = p.dx + (long long)e * p.dx_stride + (long long)(m0 + row) * p.lddx;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i >= dx_shift && c + i < dx_shift + p.dx_cols) o[c + i - dx_shift] = __uint_as_float(v[i]);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        epi_sync();
        if (t == 0) mbar_arrive(bars + C_TEMPTY + buf);
        ++g;
      }
    }
    if (t == 0) tma_store_wait_read();
    if (dbg && t == 0) dbg[15] = clock64();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace fused
}  // namespace d3b

using namespace d3b;
using namespace d3b::fused;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// bf16 tensor [members][rows][cols] (row-major, leading dim ld, member stride `stride`), box {64, box_rows, 1}
int make_map(CUtensorMap* map, const void* base, int cols, int rows, int members, long long ld, long long stride,
             int box_rows, const char* what) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (((uintptr_t)base & 15) || (ld & 7) || (members > 1 && (stride & 7)))
    return set_err(D3B_ERR_ARG, "mlp_fused: %s must be 16-byte aligned with ld/stride multiples of 8 bf16", what);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(members < 1 ? 1 : members)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)((members > 1 ? stride : ld * (long long)rows) * 2)};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r);
  return D3B_OK;
}

// 4-D map of a row-major bf16 matrix [members][rows][cols].  blocked (cols % 64 == 0): dims {64, rows, cols/64, members}
// with a 128-byte stride between column blocks and box {64, box_rows, cols/64, 1}: one operation moves `box_rows` full
// rows as [block][row][64].  Otherwise dims {cols, rows, 1, members}, box {64, box_rows, 1, 1}: the 3-D behaviour.
int make_map4(CUtensorMap* map, const void* base, int cols, int rows, int members, long long ld, long long stride,
              int box_rows, bool blocked, const char* what) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (((uintptr_t)base & 15) || (ld & 7) || (members > 1 && (stride & 7)))
    return set_err(D3B_ERR_ARG, "mlp_fused: %s must be 16-byte aligned with ld/stride multiples of 8 bf16", what);
  const cuuint64_t mstride = (cuuint64_t)((members > 1 ? stride : ld * (long long)rows) * 2);
  const int nblk = blocked ? cols / 64 : 1;
  cuuint64_t dims[4] = {(cuuint64_t)(blocked ? 64 : cols), (cuuint64_t)rows, (cuuint64_t)nblk,
                        (cuuint64_t)(members < 1 ? 1 : members)};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, blocked ? (cuuint64_t)128 : mstride, mstride};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)nblk, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(D3B_ERR_CUDA, "cuTensorMapEncodeTiled(%s, 4-D) failed: %d", what, (int)r);
  return D3B_OK;
}

template <int NH>
constexpr int fwd_epi_warps() { return NH <= 1 ? 16 : 8; }

template <int NH>
size_t fwd_smem() {
  constexpr int W_STAGES = NH <= 16 ? 4 : 3;
  return 1024 + (size_t)MAX_KB * A_KB_BYTES + (size_t)W_STAGES * W_STAGE_BYTES +
         sizeof(float) * ((size_t)NH * MAXW + MAX_LAYERS * MAXW + (size_t)(fwd_epi_warps<NH>() / 4 - 1) * BM * NH) +
         8 * B_COUNT + 64;
}

template <int NH>
int launch_fwd(const Maps& maps, const FwdParams& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    D3B_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<NH, fwd_epi_warps<NH>()>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)fwd_smem<NH>()));
    attr_set = true;
  }
  launch_pdl(mlp_forward_kernel<NH, fwd_epi_warps<NH>()>, dim3(grid), dim3(fwd_epi_warps<NH>() * 32 + 96), fwd_smem<NH>(), st,
             maps, p);
  return check_launch("mlp_forward_bf16");
}

}  // namespace

static long long* g_fused_dbg = nullptr;
// profiling hook: device buffer receiving 16 clock64() stamps per CTA of the next mlp_forward_bf16 launches
extern "C" int d3b_mlp_set_debug(void* device_buffer) {
  g_fused_dbg = (long long*)device_buffer;
  return D3B_OK;
}

// dims_host = {K_0, N_0, N_1, ..., N_{L-1}} (K_l = N_{l-1}); w_host[l] / bias_host[l] / acts_host[l] point at
// member 0 of layer l; acts_host may be NULL (nothing saved) and individual entries may be NULL.
extern "C" int d3b_mlp_forward_bf16(const void* x, int64_t ldx, int64_t stride_x, int rows, int members, int n_layers,
                                    const int* dims_host, const void* const* w_host, const int64_t* ldw_host,
                                    int64_t stride_w, const float* const* bias_host, int64_t stride_bias,
                                    void* const* acts_host, const int64_t* ld_act_host,
                                    const int64_t* stride_act_host, const float* head_w, const float* head_b,
                                    int64_t stride_head, int n_head, int head_tanh, float* head_out, int save_rows,
                                    void* stream) {
  D3B_REQUIRE(rows >= 0 && members >= 1 && n_layers >= 1 && n_layers <= MAX_LAYERS, "mlp_forward_bf16: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(x && dims_host && w_host && ldw_host && bias_host, "mlp_forward_bf16: null pointer");
  D3B_REQUIRE(n_head >= 0 && n_head <= 32, "mlp_forward_bf16: n_head must be <= 32");
  D3B_REQUIRE(n_head == 0 || (head_w && head_b && head_out), "mlp_forward_bf16: null head pointer");
  FwdParams p{};
  Maps maps;
  p.rows = rows; p.members = members; p.tiles = ceil_div(rows, BM); p.n_layers = n_layers;
  p.x_shared = stride_x == 0;
  p.bias_stride = stride_bias;
  p.head_w = head_w; p.head_b = head_b; p.head_stride = stride_head; p.n_head = n_head; p.head_tanh = head_tanh;
  p.head_out = head_out;
  p.save_rows = (save_rows > 0 && save_rows < rows) ? save_rows : rows;
  p.dbg = g_fused_dbg;
  int k = dims_host[0];
  D3B_REQUIRE(k >= 1 && k <= MAXW, "mlp_forward_bf16: input width must be in [1,256]");
  int rc = make_map(&maps.x, x, k, rows, p.x_shared ? 1 : members, ldx, stride_x, BM, "x");
  if (rc) return rc;
  for (int l = 0; l < n_layers; ++l) {
    int n = dims_host[l + 1];
    D3B_REQUIRE(n >= 16 && n <= MAXW && n % 16 == 0, "mlp_forward_bf16: layer widths must be multiples of 16 in [16,256]");
    D3B_REQUIRE(w_host[l] && bias_host[l], "mlp_forward_bf16: null layer pointer");
    p.K[l] = k; p.N[l] = n; p.bias[l] = bias_host[l];
    rc = make_map(&maps.w[l], w_host[l], k, n, members, ldw_host[l], stride_w, n, "W");
    if (rc) return rc;
    if (acts_host && acts_host[l]) {
      // per-block stores: a block leaves as soon as it is complete, which frees the operand buffer for the next
      // layer's epilogue earlier than one whole-tile store would (measured: +2 us per launch with the latter)
      const bool blocked = false;
      rc = make_map4(&maps.h[l], acts_host[l], n, rows, members, ld_act_host[l], stride_act_host[l], BM, blocked, "H");
      if (rc) return rc;
      p.save_mask |= 1 << l;
      if (blocked) p.h_blocked |= 1 << l;
    } else {
      maps.h[l] = maps.x;
    }
    k = n;
  }
  for (int l = n_layers; l < MAX_LAYERS; ++l) { maps.w[l] = maps.x; maps.h[l] = maps.x; }
  int units = p.tiles * members;
  int grid = units < kNumSM ? units : kNumSM;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_head == 0) return launch_fwd<0>(maps, p, grid, st);
  if (n_head == 1) return launch_fwd<1>(maps, p, grid, st);
  if (n_head <= 8) return launch_fwd<8>(maps, p, grid, st);
  if (n_head <= 12) return launch_fwd<12>(maps, p, grid, st);
  if (n_head <= 16) return launch_fwd<16>(maps, p, grid, st);
  return launch_fwd<32>(maps, p, grid, st);
}

namespace {

template <int NH>
size_t bwd_smem() {
  return 1024 + (size_t)2 * MAX_KB * A_KB_BYTES + (size_t)BW_STAGES * W_STAGE_BYTES +
         sizeof(float) * ((size_t)NH * MAXW + (size_t)BM * NH + (size_t)(1 + (NH <= 2 ? NH : 0)) * 4 * MAXW) +
         8 * C_COUNT + 64;
}

template <int NH>
int launch_bwd(const BwdMaps& maps, const BwdParams& p, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    D3B_CUDA(cudaFuncSetAttribute(mlp_backward_kernel<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)bwd_smem<NH>()));
    attr_set = true;
  }
  launch_pdl(mlp_backward_kernel<NH>, dim3(grid), dim3(NTHREADS), bwd_smem<NH>(), st, maps, p);
  return check_launch("mlp_backward_bf16");
}

}  // namespace

// Data-gradient chain of the network evaluated by d3b_mlp_forward_bf16 (same dims / weight shadows / saved
// activations).  d_head: fp32 [members][rows][n_head] gradient w.r.t. the head's pre-activation output.
// dbias_host != NULL selects the training form: dZ_l are stored to dz_host[l] (bf16, for the weight-gradient
// GEMMs d3b_umma_gemm_tn), bias and head gradients are RED-added into the gradient arena (member stride
// stride_grad).  dx != NULL additionally returns the gradient w.r.t. input columns [dx_col0, dx_col0+dx_cols).
extern "C" int d3b_mlp_backward_bf16(int rows, int members, int n_layers, const int* dims_host,
                                     const void* const* w_host, const int64_t* ldw_host, int64_t stride_w,
                                     const void* const* acts_host, const int64_t* ld_act_host,
                                     const int64_t* stride_act_host, void* const* dz_host, const int64_t* ld_dz_host,
                                     const int64_t* stride_dz_host, const float* d_head, const float* head_w,
                                     int64_t stride_head, int n_head, float* const* dbias_host, float* d_head_w,
                                     float* d_head_b, int64_t stride_grad, void* d_head_bf16, float* dx,
                                     int64_t lddx, int64_t stride_dx, int dx_col0, int dx_cols, void* stream) {
  D3B_REQUIRE(rows >= 0 && members >= 1 && n_layers >= 1 && n_layers <= MAX_LAYERS, "mlp_backward_bf16: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dims_host && w_host && ldw_host && acts_host && ld_act_host && stride_act_host && d_head && head_w,
              "mlp_backward_bf16: null pointer");
  D3B_REQUIRE(n_head >= 1 && n_head <= 16, "mlp_backward_bf16: n_head must be in [1,16]");
  const bool wg = dbias_host != nullptr;
  D3B_REQUIRE(!wg || (dz_host && ld_dz_host && stride_dz_host && d_head_w && d_head_b),
              "mlp_backward_bf16: weight-gradient form needs dz / head gradient pointers");
  D3B_REQUIRE(!dx || (dx_cols >= 1 && dx_col0 >= 0 && dx_col0 + dx_cols <= dims_host[0] && dx_cols + 7 <= MAXW),
              "mlp_backward_bf16: bad dx column range");
  BwdParams p{};
  BwdMaps maps;
  p.rows = rows; p.members = members; p.tiles = ceil_div(rows, BM); p.n_layers = n_layers;
  p.d_head = d_head; p.head_w = head_w; p.head_stride = stride_head; p.n_head = n_head;
  p.weight_grads = wg ? 1 : 0;
  p.d_head_w = d_head_w; p.d_head_b = d_head_b; p.grad_stride = stride_grad;
  p.d_head_bf16 = (wg && n_head > 1) ? (__nv_bfloat16*)d_head_bf16 : nullptr;
  p.dx = dx; p.lddx = lddx; p.dx_stride = stride_dx; p.dx_col0 = dx_col0; p.dx_cols = dx_cols;
  p.dbg = g_fused_dbg;
  int k = dims_host[0];
  D3B_REQUIRE(k >= 1 && k <= MAXW, "mlp_backward_bf16: input width must be in [1,256]");
  int rc;
  for (int l = 0; l < n_layers; ++l) {
    int n = dims_host[l + 1];
    D3B_REQUIRE(n >= 16 && n <= MAXW && n % 16 == 0, "mlp_backward_bf16: layer widths must be multiples of 16 in [16,256]");
    D3B_REQUIRE(w_host[l] && acts_host[l], "mlp_backward_bf16: null layer pointer");
    p.K[l] = k; p.N[l] = n;
    // W_l [n rows][k cols] row-major as the MN-major B operand: boxes of 64 reduction rows x (all | 64) columns
    const bool wb = l > 0 && k % 64 == 0, hb = n % 64 == 0;
    rc = make_map4(&maps.w[l], w_host[l], k, n, members, ldw_host[l], stride_w, BK, wb, "W");
    if (rc) return rc;
    if (wb) p.w_blocked |= 1 << l;
    rc = make_map4(&maps.h[l], acts_host[l], n, rows, members, ld_act_host[l], stride_act_host[l], BM, hb, "H");
    if (rc) return rc;
    if (hb) p.h_blocked |= 1 << l;
    if (wg) {
      D3B_REQUIRE(dz_host[l] && dbias_host[l], "mlp_backward_bf16: null dz / dbias pointer");
      rc = make_map4(&maps.dz[l], dz_host[l], n, rows, members, ld_dz_host[l], stride_dz_host[l], BM, false, "dZ");
      if (rc) return rc;  // per-block stores (see d3b_mlp_forward_bf16)
      p.dbias[l] = dbias_host[l];
    } else {
      maps.dz[l] = maps.h[l];
    }
    k = n;
  }
  for (int l = n_layers; l < MAX_LAYERS; ++l) { maps.w[l] = maps.w[0]; maps.h[l] = maps.h[0]; maps.dz[l] = maps.h[0]; }
  int units = p.tiles * members;
  int grid = units < kNumSM ? units : kNumSM;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_head == 1) return launch_bwd<1>(maps, p, grid, st);
  if (n_head <= 8) return launch_bwd<8>(maps, p, grid, st);
  if (n_head <= 12) return launch_bwd<12>(maps, p, grid, st);  // 2 x 6 actions: the default policy head
  return launch_bwd<16>(maps, p, grid, st);
}
