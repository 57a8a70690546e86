// K11: data-parallel gradient exchange.  The reference has no collective at all (single process, single
// device — SURVEY.md §2.1); sharding the minibatch over the GPUs of a box adds exactly one exchange per
// optimizer step: a sum all-reduce of the flat gradient arena (and of a few loss partial sums), issued on
// the update stream so that it is captured into the same CUDA graph as the kernels around it.
//
// NCCL is bound at run time with dlopen (the library ships inside the PyTorch wheel; libd3b.so has no
// link-time dependency on it): only ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* Comm;
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*CommDestroyFn)(Comm);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
} g_nccl;

constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclSum = 0;      // ncclSum

int nccl_fail(const char* what, int rc) {
  return d3b::set_err(D3B_ERR_CUDA, "%s: NCCL error %d (%s)", what, rc,
                      g_nccl.error_string ? g_nccl.error_string(rc) : "?");
}

}  // namespace

extern "C" int d3b_comm_load(const char* libnccl_path) {
  if (g_nccl.handle) return D3B_OK;
  const char* path = (libnccl_path && libnccl_path[0]) ? libnccl_path : "libnccl.so.2";
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return d3b::set_err(D3B_ERR_ARG, "comm_load: dlopen(%s) failed: %s", path, dlerror());
  g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  g_nccl.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
  g_nccl.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  g_nccl.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy)
    return d3b::set_err(D3B_ERR_ARG, "comm_load: %s does not export the NCCL entry points", path);
  g_nccl.handle = h;
  return D3B_OK;
}

extern "C" int d3b_comm_unique_id(void* id_out_128) {
  D3B_REQUIRE(g_nccl.handle, "comm_unique_id: call d3b_comm_load first");
  D3B_REQUIRE(id_out_128, "comm_unique_id: null pointer");
  UniqueId id;
  int rc = g_nccl.get_unique_id(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id_out_128, &id, sizeof(id));
  return D3B_OK;
}

extern "C" int d3b_comm_init(const void* id_128, int world_size, int rank, void** comm_out) {
  D3B_REQUIRE(g_nccl.handle, "comm_init: call d3b_comm_load first");
  D3B_REQUIRE(id_128 && comm_out && world_size >= 1 && rank >= 0 && rank < world_size, "comm_init: bad arguments");
  UniqueId id;
  memcpy(&id, id_128, sizeof(id));
  Comm c = nullptr;
  int rc = g_nccl.comm_init_rank(&c, world_size, id, rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  *comm_out = c;
  return D3B_OK;
}

extern "C" int d3b_allreduce_sum(void* comm, float* buf, int64_t n, void* stream) {
  D3B_REQUIRE(g_nccl.handle && comm, "allreduce_sum: communicator not initialised");
  D3B_REQUIRE(n >= 0 && (buf || n == 0), "allreduce_sum: bad arguments");
  if (n == 0) return D3B_OK;
  int rc = g_nccl.all_reduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, (Comm)comm, (cudaStream_t)stream);
  if (rc) return nccl_fail("ncclAllReduce", rc);
  d3b::count_launch();
  return D3B_OK;
}

extern "C" int d3b_comm_destroy(void* comm) {
  if (comm && g_nccl.handle) {
    int rc = g_nccl.comm_destroy((Comm)comm);
    if (rc) return nccl_fail("ncclCommDestroy", rc);
  }
  return D3B_OK;
}

// =====================================================================================================
// K10+K11 fused over NVLink peer memory: the data-parallel exchange without NCCL.
//
// Every rank maps the gradient arenas (and a small flag / exchange block) of all ranks of the box through CUDA
// IPC.  The optimizer kernel itself performs the all-reduce: after a flag handshake it reads every rank's
// gradient arena over NVLink in a fixed rank order (so all ranks compute bit-identical sums), applies Adam,
// the Polyak target sync and the bf16 shadow refresh in the same pass — one kernel instead of
// ncclAllReduce + Adam.  The few loss partial sums are exchanged by a single-warp kernel of the same kind.
// Flags are monotonically increasing update epochs: ready[r] >= e  <=>  rank r's gradients of update e are
// complete; done[r] >= e  <=>  rank r has finished reading everybody's gradients of update e (so they may
// be zeroed for update e+1; that wait + the zeroing is the first kernel of the next update).
// Flags are PUSHED: every rank's flag block holds one slot per (flag, writer rank); a writer stores its epoch into its
// slot of EVERY rank's block (W posted NVLink stores issued by W threads) and a waiter polls only its OWN block (local
// L2 hits, W threads in parallel) — a poll never crosses NVLink (round 1 polled the W-1 remote flags one after the
// other, ~2-3 us per remote poll).
// =====================================================================================================
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

namespace {

struct Imported { unsigned char handle[64]; void* base; };
std::vector<Imported> g_imported;

constexpr int kMaxRanks = 8;
struct PeerPtrs {
  const float* grads[kMaxRanks];
  int* flags[kMaxRanks];
  int world, rank;
};
// optional short vector (loss partial sums) that rides along with a gradient exchange: rank-local values in,
// all-rank sums out, through the per-rank exchange block (channel, epoch parity)
struct SmallVec {
  float* vec;
  int n, channel;
  float* xchg[kMaxRanks];
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin (a rank that never arrives becomes a trapped launch error instead of a hung box)
__device__ __forceinline__ void wait_flag_ge(const int* p, int v) {
  // ~1 minute of tolerated rank skew (e.g. a rank still staging its data) before the trap
  for (unsigned spins = 0; ld_acquire_sys(p) < v; ++spins) {
    if (spins > (1u << 25)) __trap();
    __nanosleep(spins < 1024 ? 32 : 512);
  }
}

// profiling hook (d3b_peer_set_trace): per traced kernel k and update e, slot (k * 64 + e % 64) * 4 of the buffer gets
// {globaltimer at entry, clock64 cycles spent in the rendezvous wait, clock64 cycles entry -> end of block 0, epoch}
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Trace {
  long long* slot; long long c0, c1; unsigned long long t0;
  __device__ __forceinline__ void begin(int kernel, int e) {
    slot = (g_trace && blockIdx.x == 0 && threadIdx.x == 0) ? g_trace + ((kernel * 64 + (e & 63)) * 4) : nullptr;
    if (slot) { t0 = gtime(); c0 = clock64(); c1 = c0; slot[3] = e; }
  }
  __device__ __forceinline__ void waited(long long since) { if (slot) c1 += clock64() - since; }
  __device__ __forceinline__ void end() {
    if (slot) { slot[0] = (long long)t0; slot[1] = c1 - c0; slot[2] = clock64() - c0; }
  }
};

// flag f of writer rank w lives at flags[any rank] + f * kMaxRanks + w
__device__ __forceinline__ void signal_all(const PeerPtrs& ps, int f, int e) {   // threads 0..world-1 of one block
  if ((int)threadIdx.x < ps.world) st_release_sys(ps.flags[threadIdx.x] + f * kMaxRanks + ps.rank, e);
}
__device__ __forceinline__ void wait_all(const PeerPtrs& ps, int f, int e) {     // threads 0..world-1, local polls
  if ((int)threadIdx.x < ps.world) wait_flag_ge(ps.flags[ps.rank] + f * kMaxRanks + threadIdx.x, e);
}

// Short vectors (loss partial sums) are PUSHED like the flags: writer w stores its n values into slot
// [channel][parity][w] of EVERY rank's exchange block (W independent remote stores), fences, raises its flag; after the
// flags arrived each rank sums its own block in rank order.  (Pulling them — one dependent NVLink round trip per peer,
// 2-3 us each, serialised by the volatile loads — cost ~20 us per rendezvous at 8 GPUs.)
__device__ __forceinline__ int xchg_slot(int channel, int e, int writer) {
  return ((channel * 2 + (e & 1)) * kMaxRanks + writer) * 16;
}
__device__ __forceinline__ void push_small(float* const* xchg, const PeerPtrs& ps, int channel, int e, const float* vec,
                                           int n) {   // threads 0..n-1 of one warp
  const int t = threadIdx.x;
  if (t < n) {
    const float v = vec[t];
    for (int r = 0; r < ps.world; ++r) xchg[r][xchg_slot(channel, e, ps.rank) + t] = v;
  }
  __threadfence_system();
}
__device__ __forceinline__ float sum_small(float* const* xchg, const PeerPtrs& ps, int channel, int e, int t) {
  const volatile float* mine = xchg[ps.rank];
  float s = 0.f;
  for (int r = 0; r < ps.world; ++r) s += mine[xchg_slot(channel, e, r) + t];
  return s;
}

// first kernel of an update: wait until every rank has finished reading this rank's gradients of the previous
// update, then zero them (the backward kernels accumulate with RED)
__global__ void __launch_bounds__(256) peer_wait_zero_kernel(PeerPtrs ps, int done_index, const int* epoch,
                                                             float* __restrict__ grads, long long n, int done_index2,
                                                             float* __restrict__ grads2, long long n2) {
  d3b::pdl_trigger();
  d3b::pdl_wait();
  Trace tr;
  tr.begin(0, *epoch);
  {
    const int prev = *epoch - 1;
    const long long w0 = clock64();
    wait_all(ps, done_index, prev);
    if (grads2) wait_all(ps, done_index2, prev);
    tr.waited(w0);
  }
  __syncthreads();
  tr.end();
  const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = i0; i < (n >> 2); i += stride) ((float4*)grads)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (grads2)
    for (long long i = i0; i < (n2 >> 2); i += stride) ((float4*)grads2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// sum of a short vector over the ranks through the exchange block: slot = epoch & 1
__global__ void peer_allreduce_small_kernel(float* __restrict__ vec, int n, PeerPtrs ps, int channel, const int* epoch) {
  d3b::pdl_trigger();
  d3b::pdl_wait();
  const int e = *epoch;
  const int t = threadIdx.x;
  float* const* xchg = (float* const*)ps.grads;   // the peer table carries the exchange blocks here
  push_small(xchg, ps, channel, e, vec, n);
  __syncwarp();
  signal_all(ps, 2 * channel, e);
  wait_all(ps, 2 * channel, e);
  __syncwarp();
  if (t < n) vec[t] = sum_small(xchg, ps, channel, e, t);
}

// Data-parallel CQL: the two scalar optimizer steps that need rank sums, in ONE launch (was: small all-reduce, metric
// copy, scalar Adam, finalize, scalar Adam = five launches on the critical path).
//   vec[0..2]  alpha-step partial sums {-, sum logsumexp, sum data value}      (cql_impl.py:119-141)
//   vec[3]     temperature-loss partial sum, which is also d loss / d log_temp  (sac_impl.py:128-146)
// all-reduce through the exchange block, then (temperature) metric + Adam, (alpha) loss + gradient + Adam.
// scalar blocks: {p, g, m, v} at float offsets 0, 4, 8, 12.
__device__ __forceinline__ float scalar_adam2(float* p, float G, float* m, float* v, int t, double lr) {
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
  float w1 = (float)(1.0 - b1), fb2 = (float)b2, w2 = (float)(1.0 - b2);
  float M = *m, V = *v;
  M = __fmaf_rn(w1, __fsub_rn(G, M), M);
  V = __fmul_rn(V, fb2);
  V = __fadd_rn(V, __fmul_rn(__fmul_rn(w2, G), G));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(V), (float)sqrt(bc2)), (float)eps);
  float P = __fadd_rn(*p, __fdiv_rn(__fmul_rn((float)(-(lr / bc1)), M), denom));
  *p = P; *m = M; *v = V;
  return P;
}

__global__ void dp_scalar_steps_kernel(float* __restrict__ vec, PeerPtrs ps, int channel, const int* epoch,
                                       float* temp, const int* step_temp, double lr_temp, float* metric_temp_loss,
                                       float* metric_temp, float* alpha, const int* step_alpha, double lr_alpha,
                                       float inv_eb, float cw, float threshold, float* metric_alpha_loss,
                                       float* metric_alpha) {
  d3b::pdl_trigger();
  d3b::pdl_wait();
  const int e = *epoch;
  const int t = threadIdx.x;
  const int n = temp ? 4 : 3;
  __shared__ float sum_s[4];
  float* const* xchg = (float* const*)ps.grads;   // the peer table carries the exchange blocks here
  Trace tr;
  tr.begin(1, e);
  push_small(xchg, ps, channel, e, vec, n);
  __syncwarp();
  signal_all(ps, 2 * channel, e);
  const long long w0 = clock64();
  wait_all(ps, 2 * channel, e);
  tr.waited(w0);
  tr.end();
  __syncwarp();
  if (t < n) {
    const float s = sum_small(xchg, ps, channel, e, t);
    vec[t] = s;
    sum_s[t] = s;
  }
  __syncwarp();
  if (t == 0 && temp) {  // update_temp: the loss equals its gradient
    const float G = sum_s[3];
    *metric_temp_loss = G;
    *metric_temp = expf(scalar_adam2(temp + 0, G, temp + 8, temp + 12, *step_temp, lr_temp));
    vec[3] = 0.f;
  }
  if (t == 1) {  // update_alpha: loss = -clip(exp(log_alpha)) (scaled - threshold), minimised
    const float ea = expf(alpha[0]);
    const float ca = fminf(fmaxf(ea, 0.f), 1e6f);
    const float scaled = cw * (sum_s[1] * inv_eb - sum_s[2] * inv_eb);
    *metric_alpha_loss = -(ca * (scaled - threshold));
    const float inside = (ea >= 0.f && ea <= 1e6f) ? 1.f : 0.f;
    const float G = -inside * ea * (scaled - threshold);
    *metric_alpha = expf(scalar_adam2(alpha + 0, G, alpha + 8, alpha + 12, *step_alpha, lr_alpha));
    alpha[4] = 0.f;
  }
}

struct AdamArgs {
  float* p; float* g; float* m; float* v; float* targ;
  long long n;
  const int* step;
  double lr, b1, b2, eps;
  float tau;
  __nv_bfloat16* sh_p; __nv_bfloat16* sh_t;
};
constexpr int MAX_SEG2 = 8;
struct Segs2 {
  long long param_off[MAX_SEG2], count[MAX_SEG2], shadow_off[MAX_SEG2];
  int cols[MAX_SEG2], ld[MAX_SEG2];
  int n;
  long long member_size, shadow_member;
};

__device__ __forceinline__ float adam_one2(float p, float g, float& m, float& v, float w1, float fb2, float w2,
                                           float feps, float neg_ss, float bc2_sqrt) {
  m = __fmaf_rn(w1, __fsub_rn(g, m), m);
  v = __fmul_rn(v, fb2);
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(w2, g), g));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), feps);
  return __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_ss, m), denom));
}

// sum over the ranks of float4 i of the gradient arenas, in rank order (every rank computes the same sum bit for
// bit).  All W loads are issued before the first add: W-1 of them are NVLink round trips of 2-3 us, which a
// load-add-load-add loop would serialise.
__device__ __forceinline__ float4 sum_ranks(const PeerPtrs& ps, long long i) {
  float4 x[kMaxRanks];
#pragma unroll
  for (int r = 0; r < kMaxRanks; ++r)
    if (r < ps.world) x[r] = ((const float4*)ps.grads[r])[i];
  float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < kMaxRanks; ++r)
    if (r < ps.world) { G.x += x[r].x; G.y += x[r].y; G.z += x[r].z; G.w += x[r].w; }
  return G;
}

// reduced-gradient buffers of every rank (two-shot exchange); gred[0] == nullptr selects the one-shot form
struct GredPtrs { float* gred[kMaxRanks]; };

// all-reduce (over NVLink peer memory) + Adam + Polyak + bf16 shadow refresh in one kernel.
//   one-shot (gred absent): every rank reads all W gradient arenas in rank order — W-1 remote arena reads per rank.
//   two-shot (gred present): reduce-scatter + all-gather of the reduced gradients inside the kernel.  Phase 1: rank r
//     sums its 1/W slice of the arena over all ranks (fixed rank order, so the sum exists once and is bit-identical
//     everywhere) and PUSHES it into every rank's `gred` buffer; the last block raises `done` (nobody's gradients are
//     needed any more) and `reduced`.  Phase 2: after all W `reduced` flags arrived (local polls) every rank runs Adam
//     over the whole arena from its local `gred`.  NVLink traffic per rank: 2 (W-1)/W arena sizes instead of W-1.
//     All blocks must be co-resident (they wait for each other's phase 1 through the flag): the host bounds the grid.
__global__ void __launch_bounds__(256) adam_allreduce_kernel(AdamArgs a, Segs2 segs, PeerPtrs ps, int flag_index,
                                                             const int* epoch, unsigned* block_counter, SmallVec sv,
                                                             GredPtrs gp, unsigned* block_counter2) {
  d3b::pdl_trigger();
  d3b::pdl_wait();
  const int e = *epoch;
  __shared__ bool last_block;
  const bool two_shot = gp.gred[0] != nullptr;
  Trace tr;
  tr.begin(flag_index <= 8 ? 2 : 3, e);
  if (blockIdx.x == 0 && sv.n > 0) {  // push my partial sums to every rank before announcing that my data is ready
    push_small(sv.xchg, ps, sv.channel, e, sv.vec, sv.n);
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    __threadfence_system();
    signal_all(ps, flag_index, e);  // my gradients of update e are complete
  }
  {
    const long long w0 = clock64();
    wait_all(ps, flag_index, e);
    tr.waited(w0);
  }
  __syncthreads();
  if (blockIdx.x == 0 && (int)threadIdx.x < sv.n) sv.vec[threadIdx.x] = sum_small(sv.xchg, ps, sv.channel, e, threadIdx.x);
  __shared__ float sc[2];
  if (threadIdx.x == 0) {  // double pow/sqrt once per block
    const int t = *a.step;
    const double bc1 = 1.0 - pow(a.b1, (double)t), bc2 = 1.0 - pow(a.b2, (double)t);
    sc[0] = (float)(-(a.lr / bc1));
    sc[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float w1 = (float)(1.0 - a.b1), fb2 = (float)a.b2, w2 = (float)(1.0 - a.b2), feps = (float)a.eps;
  const float neg_ss = sc[0], bc2s = sc[1];
  const float one_m_tau = (float)(1.0 - (double)a.tau);
  const long long n4 = a.n >> 2;
  if (two_shot) {
    // ---- phase 1: my slice of the arena, summed over the ranks, pushed to everybody
    const long long chunk = (n4 + ps.world - 1) / ps.world;
    const long long lo = chunk * ps.rank, hi = min(n4, lo + chunk);
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
         i += (long long)gridDim.x * blockDim.x) {
      const float4 G = sum_ranks(ps, i);
      for (int r = 0; r < ps.world; ++r) ((float4*)gp.gred[r])[i] = G;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned prev = atomicAdd(block_counter2, 1u);
      last_block = (prev == gridDim.x - 1);
      if (last_block) {
        *block_counter2 = 0u;
        __threadfence_system();
      }
    }
    __syncthreads();
    if (last_block) {
      signal_all(ps, flag_index + 1, e);   // done: I no longer read anybody's gradients of update e
      signal_all(ps, flag_index + 2, e);   // reduced: my slice has been pushed to every rank
    }
    // ---- phase 2 needs every rank's slice
    const long long w0 = clock64();
    wait_all(ps, flag_index + 2, e);
    tr.waited(w0);
    __syncthreads();
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
    if (two_shot) {
      G = __ldcg((const float4*)gp.gred[ps.rank] + i);   // written over NVLink by the slice owners: bypass L1
    } else {
      G = sum_ranks(ps, i);
    }
    float4 P = ((float4*)a.p)[i], M = ((float4*)a.m)[i], V = ((float4*)a.v)[i];
    P.x = adam_one2(P.x, G.x, M.x, V.x, w1, fb2, w2, feps, neg_ss, bc2s);
    P.y = adam_one2(P.y, G.y, M.y, V.y, w1, fb2, w2, feps, neg_ss, bc2s);
    P.z = adam_one2(P.z, G.z, M.z, V.z, w1, fb2, w2, feps, neg_ss, bc2s);
    P.w = adam_one2(P.w, G.w, M.w, V.w, w1, fb2, w2, feps, neg_ss, bc2s);
    ((float4*)a.p)[i] = P;
    ((float4*)a.m)[i] = M;
    ((float4*)a.v)[i] = V;
    float4 T = P;
    if (a.targ) {
      T = ((float4*)a.targ)[i];
      T.x = __fadd_rn(__fmul_rn(T.x, one_m_tau), __fmul_rn(a.tau, P.x));
      T.y = __fadd_rn(__fmul_rn(T.y, one_m_tau), __fmul_rn(a.tau, P.y));
      T.z = __fadd_rn(__fmul_rn(T.z, one_m_tau), __fmul_rn(a.tau, P.z));
      T.w = __fadd_rn(__fmul_rn(T.w, one_m_tau), __fmul_rn(a.tau, P.w));
      ((float4*)a.targ)[i] = T;
    }
    if (a.sh_p) {
      long long e0 = i << 2;
      long long member = e0 / segs.member_size, off = e0 - member * segs.member_size;
      int sidx = -1;
#pragma unroll
      for (int k = 0; k < MAX_SEG2; ++k)
        if (k < segs.n && off >= segs.param_off[k] && off < segs.param_off[k] + segs.count[k]) sidx = k;
      if (sidx >= 0) {
        long long rel = off - segs.param_off[sidx];
        int cols = segs.cols[sidx], ld = segs.ld[sidx];
        long long base = member * segs.shadow_member + segs.shadow_off[sidx];
        const float pv[4] = {P.x, P.y, P.z, P.w}, tv[4] = {T.x, T.y, T.z, T.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          long long r = rel + q;
          if (r < segs.count[sidx]) {
            long long row = r / cols;
            int c = (int)(r - row * cols);
            long long d = base + row * ld + c;
            a.sh_p[d] = __float2bfloat16_rn(pv[q]);
            if (a.sh_t && a.targ) a.sh_t[d] = __float2bfloat16_rn(tv[q]);
          }
        }
      }
    }
  }
  tr.end();
  // completion: the last block of this rank announces that it no longer needs anybody's gradients of update e
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned prev = atomicAdd(block_counter, 1u);
    last_block = (prev == gridDim.x - 1);
    if (last_block) {
      *block_counter = 0u;
      __threadfence_system();
    }
  }
  __syncthreads();
  if (last_block && !two_shot) signal_all(ps, flag_index + 1, e);
}

int fill_peers(PeerPtrs& ps, const void* const* grads_host, const void* const* flags_host, int world, int rank) {
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !grads_host || !flags_host) return -1;
  ps.world = world; ps.rank = rank;
  for (int r = 0; r < kMaxRanks; ++r) {
    ps.grads[r] = r < world ? (const float*)grads_host[r] : nullptr;
    ps.flags[r] = r < world ? (int*)flags_host[r] : nullptr;
    if (r < world && (!ps.grads[r] || !ps.flags[r])) return -1;
  }
  return 0;
}

}  // namespace

extern "C" int d3b_peer_set_trace(void* device_buffer) {
  long long* p = (long long*)device_buffer;
  D3B_CUDA(cudaMemcpyToSymbol(g_trace, &p, sizeof(p)));
  return D3B_OK;
}

extern "C" int d3b_peer_export(const void* ptr, void* handle_out_64, int64_t* offset_out) {
  D3B_REQUIRE(ptr && handle_out_64 && offset_out, "peer_export: null pointer");
  CUdeviceptr base = 0;
  size_t size = 0;
  // resolved through the runtime so that libd3b.so has no link-time dependency on libcuda.so (CPU-only build hosts)
  typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static GetRangeFn get_range = nullptr;
  if (!get_range) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return d3b::set_err(D3B_ERR_CUDA, "peer_export: cuMemGetAddressRange not available from the driver");
    get_range = (GetRangeFn)fp;
  }
  CUresult r = get_range(&base, &size, (CUdeviceptr)ptr);
  if (r != CUDA_SUCCESS) return d3b::set_err(D3B_ERR_CUDA, "peer_export: cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t h;
  D3B_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out_64, &h, 64);
  *offset_out = (int64_t)((CUdeviceptr)ptr - base);
  return D3B_OK;
}

extern "C" int d3b_peer_import(const void* handle_64, int64_t offset, void** ptr_out) {
  D3B_REQUIRE(handle_64 && ptr_out && offset >= 0, "peer_import: bad arguments");
  for (const Imported& im : g_imported) {
    if (memcmp(im.handle, handle_64, 64) == 0) {
      *ptr_out = (char*)im.base + offset;
      return D3B_OK;
    }
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, 64);
  void* base = nullptr;
  D3B_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  Imported im;
  memcpy(im.handle, handle_64, 64);
  im.base = base;
  g_imported.push_back(im);
  *ptr_out = (char*)base + offset;
  return D3B_OK;
}

extern "C" int d3b_peer_wait_zero(const void* const* flags_host, int world, int rank, int done_index,
                                  const int* epoch, float* grads, int64_t n, int done_index2, float* grads2,
                                  int64_t n2, void* stream) {
  D3B_REQUIRE(epoch && grads && n >= 0 && n % 4 == 0, "peer_wait_zero: bad arguments");
  D3B_REQUIRE(!grads2 || (n2 >= 0 && n2 % 4 == 0), "peer_wait_zero: bad second arena");
  PeerPtrs ps{};
  D3B_REQUIRE(fill_peers(ps, flags_host, flags_host, world, rank) == 0, "peer_wait_zero: bad peer table");
  long long blocks = ((n > n2 || !grads2 ? n : n2) / 4 + 255) / 256;
  if (blocks > 2 * d3b::kNumSM) blocks = 2 * d3b::kNumSM;
  if (blocks < 1) blocks = 1;
  d3b::launch_pdl(peer_wait_zero_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, ps, done_index, epoch, grads, (long long)n,
                                                                           done_index2, grads2, (long long)n2);
  return d3b::check_launch("peer_wait_zero");
}

extern "C" int d3b_dp_scalar_steps(float* vec, const void* const* xchg_host, const void* const* flags_host, int world,
                                   int rank, int channel, const int* epoch, float* temp_scalar, const int* step_temp,
                                   double lr_temp, float* metric_temp_loss, float* metric_temp, float* alpha_scalar,
                                   const int* step_alpha, double lr_alpha, float inv_members_batch,
                                   float conservative_weight, float alpha_threshold, float* metric_alpha_loss,
                                   float* metric_alpha, void* stream) {
  D3B_REQUIRE(vec && epoch && alpha_scalar && step_alpha && metric_alpha_loss && metric_alpha && channel >= 0 && channel < 4,
              "dp_scalar_steps: bad arguments");
  D3B_REQUIRE(!temp_scalar || (step_temp && metric_temp_loss && metric_temp), "dp_scalar_steps: null temperature pointers");
  PeerPtrs ps{};
  D3B_REQUIRE(fill_peers(ps, xchg_host, flags_host, world, rank) == 0, "dp_scalar_steps: bad peer table");
  d3b::launch_pdl(dp_scalar_steps_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, vec, ps, channel, epoch, temp_scalar, step_temp, lr_temp,
                                                              metric_temp_loss, metric_temp, alpha_scalar, step_alpha,
                                                              lr_alpha, inv_members_batch, conservative_weight,
                                                              alpha_threshold, metric_alpha_loss, metric_alpha);
  return d3b::check_launch("dp_scalar_steps");
}

extern "C" int d3b_peer_allreduce_small(float* vec, int n, const void* const* xchg_host, const void* const* flags_host,
                                        int world, int rank, int channel, const int* epoch, void* stream) {
  D3B_REQUIRE(vec && n >= 1 && n <= 16 && channel >= 0 && channel < 4 && epoch, "peer_allreduce_small: bad arguments");
  PeerPtrs ps{};
  D3B_REQUIRE(fill_peers(ps, xchg_host, flags_host, world, rank) == 0, "peer_allreduce_small: bad peer table");
  d3b::launch_pdl(peer_allreduce_small_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, vec, n, ps, channel, epoch);
  return d3b::check_launch("peer_allreduce_small");
}

// adam_step_shadow with the gradient all-reduce fused in (see the block comment above).  flag_index: position of
// this arena's {ready, done} pair in the flag block; block_counter: one zero-initialised uint32 per arena.
extern "C" int d3b_adam_step_peer(float* params, float* exp_avg, float* exp_avg_sq, float* target, int64_t n,
                                  const int* step, double lr, double beta1, double beta2, double eps, float tau,
                                  void* shadow_params, void* shadow_target, const int64_t* table_host, int n_segments,
                                  int64_t member_size, int64_t shadow_member, const void* const* grads_host,
                                  const void* const* flags_host, int world, int rank, int flag_index,
                                  const int* epoch, void* block_counter, float* small_vec, int small_n,
                                  const void* const* xchg_host, int small_channel, const void* const* gred_host,
                                  void* block_counter2, void* stream) {
  D3B_REQUIRE(n >= 0 && n % 4 == 0 && params && exp_avg && exp_avg_sq && step && epoch && block_counter,
              "adam_step_peer: bad arguments");
  D3B_REQUIRE(!gred_host || block_counter2, "adam_step_peer: the two-shot exchange needs its own block counter");
  if (n == 0) return D3B_OK;
  PeerPtrs ps{};
  D3B_REQUIRE(fill_peers(ps, grads_host, flags_host, world, rank) == 0, "adam_step_peer: bad peer table");
  AdamArgs a{};
  a.p = params; a.g = nullptr; a.m = exp_avg; a.v = exp_avg_sq; a.targ = target; a.n = n; a.step = step;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.tau = tau;
  a.sh_p = (__nv_bfloat16*)shadow_params; a.sh_t = (__nv_bfloat16*)shadow_target;
  Segs2 segs{};
  if (shadow_params) {
    D3B_REQUIRE(table_host && n_segments >= 1 && n_segments <= MAX_SEG2 && member_size > 0, "adam_step_peer: bad shadow table");
    segs.n = n_segments; segs.member_size = member_size; segs.shadow_member = shadow_member;
    for (int k = 0; k < n_segments; ++k) {
      const int64_t* t = table_host + 5 * k;
      segs.param_off[k] = t[0]; segs.count[k] = t[1] * t[2]; segs.cols[k] = (int)t[2];
      segs.shadow_off[k] = t[3]; segs.ld[k] = (int)t[4];
    }
  }
  SmallVec sv{};
  if (small_vec && small_n > 0) {
    D3B_REQUIRE(small_n <= 16 && small_channel >= 0 && small_channel < 4 && xchg_host, "adam_step_peer: bad small vector");
    sv.vec = small_vec; sv.n = small_n; sv.channel = small_channel;
    for (int r = 0; r < world; ++r) {
      sv.xchg[r] = (float*)xchg_host[r];
      D3B_REQUIRE(sv.xchg[r], "adam_step_peer: null exchange block");
    }
  }
  GredPtrs gp{};
  if (gred_host)
    for (int r = 0; r < world; ++r) {
      gp.gred[r] = (float*)gred_host[r];
      D3B_REQUIRE(gp.gred[r], "adam_step_peer: null reduced-gradient buffer");
    }
  // every block waits for the other blocks (and ranks) in the middle of the kernel: all must be co-resident
  static int per_sm = 0;
  if (!per_sm) {
    D3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adam_allreduce_kernel, 256, 0));
    if (per_sm > 2) per_sm = 2;
    D3B_REQUIRE(per_sm >= 1, "adam_step_peer: kernel does not fit an SM");
  }
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > (long long)per_sm * d3b::kNumSM) blocks = (long long)per_sm * d3b::kNumSM;
  d3b::launch_pdl(adam_allreduce_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a, segs, ps, flag_index, epoch,
                                                                           (unsigned*)block_counter, sv, gp,
                                                                           (unsigned*)block_counter2);
  return d3b::check_launch("adam_step_peer");
}
