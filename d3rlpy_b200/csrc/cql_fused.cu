// CQL/SAC update, fused "glue" kernels: at batch 256 the update is launch-latency bound, so everything that
// sits between the tensor-core launches is collapsed into a handful of kernels:
//   begin_step      : Adam/noise counters += mask, loss partial sums zeroed            (was tick + memset)
//   update_prologue : begin_step + the update's Philox noise + bf16 policy-input rows in one launch
//   cql_rows        : every row the critics see in one update — [data | pi(s_t) | pi(s_t+1) | random] for the
//                     critic step and for the alpha step, the target row tanh(mu(s')) (or a sampled one for
//                     soft_q_backup) and the actor row — written as bf16 (fp32 in fp32 mode) straight into the GEMM operand, with
//                     all tanh-Gaussian log-probs (cql_impl.py:143-204, policies.py:167-249,
//                     distributions.py:91-143).                                       (was 15 launches)
//   sac_temp_step   : update_temp loss + d/dlog_temp + Adam (sac_impl.py:123-146)      (was 2)
//   cql_loss_step   : IS-logsumexp conservative term + TD term + softmax gradient seed, then — by the last block
//                     to finish, which first adds the per-block partial sums in a fixed order (bit-reproducible) —
//                     the scalar tail: critic metric, or alpha loss + d/dlog_alpha + Adam
//                     (cql_impl.py:110-141,196-223)                                    (was 2-3)
//   sac_actor_step  : actor loss + arg-min routing of dQ + metric                      (was 2)
#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"

namespace d3b {

__device__ __forceinline__ float clampf3(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float softplusf3(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void begin_step_kernel(int* counters, int n, unsigned mask, float* slots, int n_slots) {
  pdl_trigger();
  pdl_wait();
  int i = threadIdx.x;
  if (i < n && ((mask >> i) & 1u)) counters[i] += 1;
  for (int j = i; j < n_slots; j += blockDim.x) slots[j] = 0.f;
}

// Update prologue in ONE launch (begin_step + noise_fill + to_bf16 of the policy input were three):
//   * loss partial sums / metric slots zeroed (block 0);
//   * the update's noise arena drawn with Philox at epoch = counters[draw] + 1 (read BEFORE any bump: the counters are
//     bumped by the last block to finish, so every block sees the same epoch without a grid barrier);
//   * rows x cols of fp32 `src` (policy input [obs; next_obs]) converted to a bf16 GEMM operand (optional).
struct PrologueParams {
  int* counters; int n_counters; unsigned mask; int draw;
  float* slots; int n_slots;
  float* noise; long long n_normal, n_uniform; unsigned long long seed; int noise_blocks;
  const float* src; long long lds; int rows, cols; __nv_bfloat16* dst; long long ldd;
  unsigned* done;
};

__global__ void __launch_bounds__(256) update_prologue_kernel(PrologueParams p) {
  pdl_trigger();
  pdl_wait();
  const int t = threadIdx.x;
  const uint32_t epoch = (uint32_t)(p.counters[p.draw] + (int)((p.mask >> p.draw) & 1u));
  if (blockIdx.x == 0)
    for (int j = t; j < p.n_slots; j += blockDim.x) p.slots[j] = 0.f;
  if ((int)blockIdx.x < p.noise_blocks) {
    const long long n = p.n_normal + p.n_uniform;
    long long q = (long long)blockIdx.x * blockDim.x + t;
    const long long stride = (long long)p.noise_blocks * blockDim.x;
    for (; (q << 2) < n; q += stride) noise_quad(p.noise, q, p.n_normal, n, p.seed, epoch);
  } else if (p.src) {
    // element (r, c) of the padded bf16 operand; padding columns stay zero (written once at allocation)
    const long long total = (long long)p.rows * p.cols;
    long long i = (long long)(blockIdx.x - p.noise_blocks) * blockDim.x + t;
    const long long stride = (long long)(gridDim.x - p.noise_blocks) * blockDim.x;
    for (; i < total; i += stride) {
      const int r = (int)(i / p.cols), c = (int)(i % p.cols);
      p.dst[(long long)r * p.ldd + c] = __float2bfloat16_rn(__ldg(p.src + (long long)r * p.lds + c));
    }
  }
  // every block has read the old counters by now; the last one to get here bumps them
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.done, 1u);
    if (prev == gridDim.x - 1) {
      for (int i = 0; i < p.n_counters; ++i)
        if ((p.mask >> i) & 1u) p.counters[i] += 1;
      *p.done = 0u;
    }
  }
}

struct RowsParams {
  const float* head;      // [2B][2A]: rows 0..B-1 = policy(obs_t), rows B..2B-1 = policy(obs_tp1)
  const float* obs;       // [B][O]
  const float* next_obs;  // [B][O]
  const float* act;       // [B][A]
  int B, N, O, A;
  float min_logstd, max_logstd;
  void* X;                // [rows_total][ldx], bf16 (tensor-core mode) or fp32 (fp32 mode) GEMM operand rows
  long long ldx;
  // per IS group g (0 = critic step, 1 = alpha step): noise and log-prob outputs
  const float* eps_t[2];      // [N][B][A]
  const float* eps_tp1[2];    // [N][B][A]
  const float* rand_act[2];   // [B*N][A] uniform(-1,1)
  float* logp_t[2];           // [B*N]
  float* logp_tp1[2];         // [B*N]
  int n_groups;
  long long group_row0[2];    // first row of each group in X
  long long target_row0;      // B rows [next_obs | a'] ; a' = tanh(mu) or a sample (soft backup)
  const float* eps_soft;      // null = deterministic backup
  float* logp_soft;
  long long actor_row0;       // B rows [obs | a]
  const float* eps_actor;
  float* logp_actor;
  const float* eps_temp;      // update_temp sample: log-prob only
  float* logp_temp;
};

// tanh-Gaussian sample + log-prob for the lanes of one warp (A <= 32 per pass); returns this lane's partial logp
__device__ __forceinline__ float sample_action(const float* __restrict__ head_row, int A, int j, float e,
                                               float min_ls, float max_ls, float& a_out) {
  float mu = __ldg(head_row + j);
  float ls = clampf3(__ldg(head_row + A + j), min_ls, max_ls);
  float sd = expf(ls);
  float u = mu + e * sd;
  a_out = tanhf(u);
  float d = u - mu;
  float nlp = -(d * d) / (2.f * sd * sd) - ls - 0.91893853320467267f;
  float jac = 2.f * (0.69314718055994529f - u - softplusf3(-2.f * u));
  return nlp - jac;
}

__device__ __forceinline__ void put(__nv_bfloat16* x, int j, float v) { x[j] = __float2bfloat16_rn(v); }
__device__ __forceinline__ void put(float* x, int j, float v) { x[j] = v; }

// one warp per work item; items: per group R = B(1+3N) rows, then B target rows, B actor rows, B temp items
template <typename T>
__global__ void __launch_bounds__(256) cql_rows_kernel(RowsParams p) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // 32-bit index arithmetic (the host checks that the row count fits): 64-bit divisions are slow on the GPU
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int B = p.B, N = p.N, O = p.O, A = p.A;
  const int BN = B * N;
  const int R = B + 3 * BN;
  const int n_is = R * p.n_groups;
  const float* obs_row;
  T* const X = (T*)p.X;
  T* x = nullptr;
  // decode
  int kind;         // 0 data, 1 pi(s_t), 2 pi(s_tp1), 3 random, 4 target, 5 actor, 6 temp
  int b, k = 0, g = 0;
  if (item < n_is) {
    g = item >= R ? 1 : 0;
    int r = item - g * R;
    if (r < B) { kind = 0; b = r; }
    else {
      r -= B;
      kind = 1 + r / BN;
      r -= (kind - 1) * BN;
      b = r / N;
      k = r - b * N;
    }
    x = X + (p.group_row0[g] + (long long)(item - g * R)) * p.ldx;
    obs_row = p.obs + (long long)b * O;
  } else {
    int r = item - n_is;
    if (r < B) { kind = 4; b = r; x = X + (p.target_row0 + b) * p.ldx; obs_row = p.next_obs + (long long)b * O; }
    else if (r < 2 * B) { kind = 5; b = r - B; x = X + (p.actor_row0 + b) * p.ldx; obs_row = p.obs + (long long)b * O; }
    else if (r < 3 * B && p.eps_temp) { kind = 6; b = r - 2 * B; obs_row = nullptr; }
    else return;
  }
  if (x) {
    for (int j = lane; j < O; j += 32) put(x, j, __ldg(obs_row + j));
  }
  float lp = 0.f;
  for (int j = lane; j < A; j += 32) {
    float a = 0.f;
    switch (kind) {
      case 0: a = __ldg(p.act + (long long)b * A + j); break;
      case 1: lp += sample_action(p.head + (long long)b * 2 * A, A, j, __ldg(p.eps_t[g] + ((long long)k * B + b) * A + j),
                                  p.min_logstd, p.max_logstd, a); break;
      case 2: lp += sample_action(p.head + (long long)(B + b) * 2 * A, A, j,
                                  __ldg(p.eps_tp1[g] + ((long long)k * B + b) * A + j), p.min_logstd, p.max_logstd, a);
              break;
      case 3: a = __ldg(p.rand_act[g] + ((long long)b * N + k) * A + j); break;
      case 4:
        if (p.eps_soft) lp += sample_action(p.head + (long long)(B + b) * 2 * A, A, j,
                                            __ldg(p.eps_soft + (long long)b * A + j), p.min_logstd, p.max_logstd, a);
        else a = tanhf(__ldg(p.head + (long long)(B + b) * 2 * A + j));
        break;
      case 5: lp += sample_action(p.head + (long long)b * 2 * A, A, j, __ldg(p.eps_actor + (long long)b * A + j),
                                  p.min_logstd, p.max_logstd, a); break;
      default: lp += sample_action(p.head + (long long)b * 2 * A, A, j, __ldg(p.eps_temp + (long long)b * A + j),
                                   p.min_logstd, p.max_logstd, a); break;
    }
    if (x) put(x, O + j, a);
  }
  if (kind == 1 || kind == 2 || kind >= 4) {
    lp = warp_sum(lp);
    if (lane == 0) {
      if (kind == 1) p.logp_t[g][(long long)b * N + k] = lp;
      else if (kind == 2) p.logp_tp1[g][(long long)b * N + k] = lp;
      else if (kind == 4) { if (p.logp_soft) p.logp_soft[b] = lp; }
      else if (kind == 5) p.logp_actor[b] = lp;
      else p.logp_temp[b] = lp;
    }
  }
}

// exact restatement of torch.optim.Adam's single-tensor step for a scalar.  The two bias
// corrections (double pow: a few hundred dependent FP64 instructions each) depend only on the step counter, so the
// callers evaluate them on two otherwise idle warps (`scalar_adam_bias`, threads 32 and 64) while the block reduction
// that produces the gradient is still running, instead of in front of the parameter write.
__device__ __forceinline__ void scalar_adam_bias(float* sc /*shared [2]*/, const int* step, double lr, double b1,
                                                 double b2) {
  if (threadIdx.x == 32) sc[0] = (float)(-(lr / (1.0 - pow(b1, (double)*step))));
  else if (threadIdx.x == 64) sc[1] = (float)sqrt(1.0 - pow(b2, (double)*step));
}

__device__ __forceinline__ float scalar_adam_update(float* p, float G, float* m, float* v, float neg_step_size,
                                                    float bc2_sqrt, double b1, double b2, double eps) {
  float w1 = (float)(1.0 - b1), fb2 = (float)b2, w2 = (float)(1.0 - b2);
  float M = *m, V = *v;
  M = __fmaf_rn(w1, __fsub_rn(G, M), M);
  V = __fmul_rn(V, fb2);
  V = __fadd_rn(V, __fmul_rn(__fmul_rn(w2, G), G));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(V), bc2_sqrt), (float)eps);
  float P = __fadd_rn(*p, __fdiv_rn(__fmul_rn(neg_step_size, M), denom));
  *p = P;
  *m = M;
  *v = V;
  return P;
}

// update_temp: loss = -(exp(log_temp) (logp - A)).mean(); grad == loss; Adam; metrics {loss, exp(new log_temp)}
__global__ void __launch_bounds__(1024) sac_temp_step_kernel(const float* __restrict__ logp, float* scalar /*p,g,m,v at stride 4*/,
                                                             const int* step, int B, int A, float inv_b, double lr,
                                                             float* metric_loss, float* metric_exp) {
  pdl_trigger();
  pdl_wait();
  __shared__ float adam_sc[2];
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += __ldg(logp + b) - (float)A;
  scalar_adam_bias(adam_sc, step, lr, 0.9, 0.999);
  s = block_sum(s);   // its barriers order adam_sc before thread 0's read
  if (threadIdx.x == 0) {
    float l = -expf(scalar[0]) * s * inv_b;
    *metric_loss = l;
    float P = scalar_adam_update(scalar + 0, l, scalar + 8, scalar + 12, adam_sc[0], adam_sc[1], 0.9, 0.999, 1e-8);
    scalar[4] = 0.f;
    *metric_exp = expf(P);
  }
}

struct LossParams {
  const float* q; long long sQ;            // [E][R] rows: data B | pi_t BN | pi_tp1 BN | random BN
  const float* q_targ; long long sQt; int Et;   // target critics on the target rows [Et][B] (min inside), or
  const float* q_tpn;                      // a ready target [B] (soft backup / injected), or both null (no TD)
  const float* rew; const float* term; const float* nsteps; float gamma;
  const float* logp_t; const float* logp_tp1; int N, A;
  float* scalar_alpha;                     // log_alpha {p,g,m,v} at float stride 4
  float cw, threshold;
  float* dq; long long sDq;                // gradient seed [E][R] (critic mode) or null
  float* sums;                             // 3 sums (written once, by the last block)
  unsigned* done;                          // workspace: word 0 = block completion counter (self-resetting), words
                                           // [4, 4 + 3 gridDim.x) = per-block partial sums (fixed-order final sum:
                                           // metrics and the alpha gradient are bit-reproducible run to run)
  int B, E; float inv_b, inv_eb;
  int mode;                                // 0 critic loss metric; 1 alpha loss + Adam on log_alpha
  const int* step_alpha; double lr_alpha;
  float* metric; float* metric_exp;
};

// One WARP per (member e, batch row b): the 3N importance-sampling values are spread over the lanes (coalesced
// loads), max / sum-exp by warp shuffles, softmax gradient written back by the same lanes.
__global__ void __launch_bounds__(256) cql_loss_step_kernel(LossParams p) {
  pdl_trigger();
  pdl_wait();
  const int B = p.B, E = p.E, N = p.N;
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float td = 0.f, lse_v = 0.f, data_v = 0.f;
  const bool td_enabled = p.mode == 0 && (p.q_targ || p.q_tpn);
  if (item < B * E) {
    const int e = item / B, b = item % B;
    const float* qe = p.q + (long long)e * p.sQ;
    float* dqe = p.dq ? p.dq + (long long)e * p.sDq : nullptr;
    const float qd = __ldg(qe + b);
    float g = 0.f;
    if (td_enabled) {
      float tq;
      if (p.q_tpn) tq = __ldg(p.q_tpn + b);
      else {
        tq = __ldg(p.q_targ + b);
        for (int i = 1; i < p.Et; ++i) tq = fminf(tq, __ldg(p.q_targ + (long long)i * p.sQt + b));
      }
      float n = __ldg(p.nsteps + b);
      float gp = n == 1.f ? p.gamma : powf(p.gamma, n);
      float y = __ldg(p.rew + b) + gp * tq * (1.f - __ldg(p.term + b));
      float d = qd - y;
      td = d * d;
      g = 2.f * p.inv_b * d;
    }
    const float ca = N > 0 ? clampf3(expf(p.scalar_alpha[0]), 0.f, 1e6f) : 0.f;
    const float c = ca * p.cw * p.inv_eb;
    const float rand_lp = (float)p.A * -0.69314718055994529f;
    const long long BN = (long long)B * N;
    // value index v in [0, 3N): group = v / N (pi(s_t), pi(s_tp1), random), sample k = v % N
    float mx = -INFINITY, s;
    if (3 * N <= 32) {
      // one value per lane: a single round of loads feeds the max, the sum and the gradient (default N = 10)
      const int v = lane;
      float x = -INFINITY;
      long long idx = 0;
      if (v < 3 * N) {
        int grp = v / N, k = v - grp * N;
        idx = B + grp * BN + (long long)b * N + k;
        float q = __ldg(qe + idx);
        float off = grp == 0 ? __ldg(p.logp_t + (long long)b * N + k)
                             : (grp == 1 ? __ldg(p.logp_tp1 + (long long)b * N + k) : rand_lp);
        x = q - off;
      }
      mx = warp_max(x);
      const float ex = v < 3 * N ? expf(x - mx) : 0.f;
      s = N > 0 ? warp_sum(ex) : 1.f;
      if (dqe) {
        if (v < 3 * N) dqe[idx] = ex * (c / s);
        if (lane == 0) dqe[b] = g - c;
      }
    } else {
      for (int v = lane; v < 3 * N; v += 32) {
        int grp = v / N, k = v - grp * N;
        float q = __ldg(qe + B + grp * BN + (long long)b * N + k);
        float off = grp == 0 ? __ldg(p.logp_t + (long long)b * N + k)
                             : (grp == 1 ? __ldg(p.logp_tp1 + (long long)b * N + k) : rand_lp);
        mx = fmaxf(mx, q - off);
      }
      mx = warp_max(mx);
      s = 0.f;
      for (int v = lane; v < 3 * N; v += 32) {
        int grp = v / N, k = v - grp * N;
        float q = __ldg(qe + B + grp * BN + (long long)b * N + k);
        float off = grp == 0 ? __ldg(p.logp_t + (long long)b * N + k)
                             : (grp == 1 ? __ldg(p.logp_tp1 + (long long)b * N + k) : rand_lp);
        s += expf(q - off - mx);
      }
      s = warp_sum(s);
      if (dqe) {
        const float inv_s = c / s;
        for (int v = lane; v < 3 * N; v += 32) {
          int grp = v / N, k = v - grp * N;
          long long idx = B + grp * BN + (long long)b * N + k;
          float q = __ldg(qe + idx);
          float off = grp == 0 ? __ldg(p.logp_t + (long long)b * N + k)
                               : (grp == 1 ? __ldg(p.logp_tp1 + (long long)b * N + k) : rand_lp);
          dqe[idx] = expf(q - off - mx) * inv_s;
        }
        if (lane == 0) dqe[b] = g - c;
      }
    }
    if (lane == 0) {
      lse_v = N > 0 ? mx + logf(s) : 0.f;
      data_v = N > 0 ? qd : 0.f;
    } else {
      td = 0.f;
    }
  }
  // the three partial sums of the block in one barrier round (lane 0 of each warp holds its row's values)
  __shared__ float red3[3][8];
  if (lane == 0) {
    const int w = threadIdx.x >> 5;
    red3[0][w] = td; red3[1][w] = lse_v; red3[2][w] = data_v;
  }
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a0 += red3[0][w]; a1 += red3[1][w]; a2 += red3[2][w]; }
    float* part = reinterpret_cast<float*>(p.done + 4);
    part[blockIdx.x] = a0;
    part[gridDim.x + blockIdx.x] = a1;
    part[2 * gridDim.x + blockIdx.x] = a2;
    __threadfence();
    unsigned prev = atomicAdd(p.done, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // ---- the last block to finish adds the per-block partial sums in a fixed order ...
  __threadfence();
  __shared__ float adam_sc[2];
  if (p.mode == 1 && N > 0) scalar_adam_bias(adam_sc, p.step_alpha, p.lr_alpha, 0.9, 0.999);
  const volatile float* part = reinterpret_cast<const volatile float*>(p.done + 4);
  float s3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {   // block_sum's barriers order adam_sc before thread 0's read below
    float acc = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) acc += part[k * gridDim.x + i];
    s3[k] = block_sum(acc);
  }
  if (threadIdx.x != 0) return;
  // ---- ... and runs the scalar tail (cql_impl.py:217-223, 119-141)
  float sv[3] = {td_enabled ? s3[0] : 0.f, s3[1], s3[2]};
  p.sums[0] = sv[0]; p.sums[1] = sv[1]; p.sums[2] = sv[2];
  float tdm = sv[0] * p.inv_b;
  if (N == 0) {  // plain TD loss (no conservative term): TD3+BC / BCQ critics
    *p.metric = tdm;
    *p.done = 0u;
    return;
  }
  float ea = expf(p.scalar_alpha[0]);
  float ca = clampf3(ea, 0.f, 1e6f);
  float scaled = p.cw * (sv[1] * p.inv_eb - sv[2] * p.inv_eb);
  float cons = ca * (scaled - p.threshold);
  if (p.mode == 0) {
    *p.metric = tdm + cons;
  } else {
    *p.metric = -cons;
    float inside = (ea >= 0.f && ea <= 1e6f) ? 1.f : 0.f;
    float G = -inside * ea * (scaled - p.threshold);
    float P = scalar_adam_update(p.scalar_alpha + 0, G, p.scalar_alpha + 8, p.scalar_alpha + 12, adam_sc[0],
                                 adam_sc[1], 0.9, 0.999, 1e-8);
    p.scalar_alpha[4] = 0.f;
    *p.metric_exp = expf(P);
  }
  *p.done = 0u;
}

// SAC actor loss (sac_impl.py:114-121): mean_b(exp(log_temp) logp_b - min_e Q_e); dQ to the arg-min member.
__global__ void __launch_bounds__(256) sac_actor_step_kernel(const float* __restrict__ q, long long sQ,
                                                             const float* __restrict__ logp,
                                                             const float* __restrict__ log_temp,
                                                             float* __restrict__ dq, long long sDq,
                                                             float* __restrict__ loss_sum, unsigned* done,
                                                             float* metric, int B, int E, float inv_b) {
  pdl_trigger();
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (b < B) {
    float qm = __ldg(q + b);
    int arg = 0;
    for (int e = 1; e < E; ++e) {
      float v = __ldg(q + (long long)e * sQ + b);
      if (v < qm) { qm = v; arg = e; }
    }
    for (int e = 0; e < E; ++e) dq[(long long)e * sDq + b] = (e == arg) ? -inv_b : 0.f;
    l = (expf(__ldg(log_temp)) * __ldg(logp + b) - qm) * inv_b;
  }
  l = block_sum(l);
  if (threadIdx.x == 0) {
    // done: word 0 = block counter (self-resetting), words [4, 4 + gridDim.x) = per-block partial sums, added in
    // block order by the last block to finish (bit-reproducible metric)
    volatile float* part = reinterpret_cast<volatile float*>(done + 4);
    part[blockIdx.x] = l;
    __threadfence();
    unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence();
      float acc = 0.f;
      for (unsigned i = 0; i < gridDim.x; ++i) acc += part[i];
      *loss_sum = acc;
      *metric = acc;
      *done = 0u;
    }
  }
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

extern "C" int d3b_begin_step(int* counters, int n, unsigned mask, float* slots, int n_slots, void* stream) {
  D3B_REQUIRE(counters && n >= 0 && n <= 32 && slots && n_slots >= 0, "begin_step: bad arguments");
  launch_pdl(begin_step_kernel, dim3(1), dim3(64), 0, ST, counters, n, mask, slots, n_slots);
  return check_launch("begin_step");
}

extern "C" int d3b_update_prologue(int* counters, int n_counters, unsigned mask, int draw_index, float* slots,
                                   int n_slots, float* noise, int64_t n_normal, int64_t n_uniform, uint64_t seed,
                                   const float* src, int64_t lds, int rows, int cols, void* dst_bf16, int64_t ldd,
                                   void* done_counter, void* stream) {
  D3B_REQUIRE(counters && n_counters >= 0 && n_counters <= 32 && draw_index >= 0 && draw_index < 32 && slots &&
                  n_slots >= 0 && done_counter,
              "update_prologue: bad arguments");
  D3B_REQUIRE(n_normal >= 0 && n_uniform >= 0 && (n_normal + n_uniform == 0 || noise), "update_prologue: noise arena");
  D3B_REQUIRE(!src || (dst_bf16 && rows >= 0 && cols >= 1), "update_prologue: conversion arguments");
  PrologueParams p{};
  p.counters = counters; p.n_counters = n_counters; p.mask = mask; p.draw = draw_index;
  p.slots = slots; p.n_slots = n_slots;
  p.noise = noise; p.n_normal = n_normal; p.n_uniform = n_uniform; p.seed = seed;
  const long long quads = (n_normal + n_uniform + 3) / 4;
  p.noise_blocks = (int)(quads == 0 ? 0 : (quads + 255) / 256 > 2 * kNumSM ? 2 * kNumSM : (quads + 255) / 256);
  p.src = src; p.lds = lds; p.rows = rows; p.cols = cols; p.dst = (__nv_bfloat16*)dst_bf16; p.ldd = ldd;
  p.done = (unsigned*)done_counter;
  int conv_blocks = 0;
  if (src && rows > 0) {
    const long long total = (long long)rows * cols;
    conv_blocks = (int)((total + 255) / 256 > kNumSM ? kNumSM : (total + 255) / 256);
  }
  int blocks = p.noise_blocks + conv_blocks;
  if (blocks < 1) blocks = 1;
  launch_pdl(update_prologue_kernel, dim3((unsigned)blocks), dim3(256), 0, ST, p);
  return check_launch("update_prologue");
}

// ptrs_host: 16 device pointers in the order
//   {eps_t[0], eps_tp1[0], rand[0], logp_t[0], logp_tp1[0], eps_t[1], eps_tp1[1], rand[1], logp_t[1], logp_tp1[1],
//    eps_soft, logp_soft, eps_actor, logp_actor, eps_temp, logp_temp}
// rows_host: {group_row0[0], group_row0[1], target_row0, actor_row0}
static int cql_rows_any(bool f32, const float* head, const float* obs, const float* next_obs, const float* act, int batch,
                            int n_action_samples, int obs_dim, int act_dim, float min_logstd, float max_logstd,
                            void* x_bf16, int64_t ldx, int n_groups, const void* const* ptrs_host,
                            const int64_t* rows_host, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_action_samples >= 0 && obs_dim >= 1 && act_dim >= 1, "cql_rows: bad sizes");
  D3B_REQUIRE(n_groups >= 1 && n_groups <= 2, "cql_rows: n_groups must be 1 or 2");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(head && obs && next_obs && act && x_bf16 && ptrs_host && rows_host, "cql_rows: null pointer");
  RowsParams p{};
  p.head = head; p.obs = obs; p.next_obs = next_obs; p.act = act;
  p.B = batch; p.N = n_action_samples; p.O = obs_dim; p.A = act_dim;
  p.min_logstd = min_logstd; p.max_logstd = max_logstd;
  p.X = x_bf16; p.ldx = ldx; p.n_groups = n_groups;
  for (int g = 0; g < 2; ++g) {
    p.eps_t[g] = (const float*)ptrs_host[5 * g + 0];
    p.eps_tp1[g] = (const float*)ptrs_host[5 * g + 1];
    p.rand_act[g] = (const float*)ptrs_host[5 * g + 2];
    p.logp_t[g] = (float*)ptrs_host[5 * g + 3];
    p.logp_tp1[g] = (float*)ptrs_host[5 * g + 4];
    p.group_row0[g] = rows_host[g];
    if (g < n_groups && n_action_samples > 0)  // N == 0 (plain SAC): data rows only, no importance-sampling groups
      D3B_REQUIRE(p.eps_t[g] && p.eps_tp1[g] && p.rand_act[g] && p.logp_t[g] && p.logp_tp1[g], "cql_rows: null group pointer");
  }
  p.eps_soft = (const float*)ptrs_host[10]; p.logp_soft = (float*)ptrs_host[11];
  p.eps_actor = (const float*)ptrs_host[12]; p.logp_actor = (float*)ptrs_host[13];
  p.eps_temp = (const float*)ptrs_host[14]; p.logp_temp = (float*)ptrs_host[15];
  D3B_REQUIRE(p.eps_actor && p.logp_actor, "cql_rows: null actor pointers");
  D3B_REQUIRE(!p.eps_temp || p.logp_temp, "cql_rows: null temp log-prob pointer");
  D3B_REQUIRE(!p.eps_soft || p.logp_soft, "cql_rows: null soft-backup log-prob pointer");
  p.target_row0 = rows_host[2]; p.actor_row0 = rows_host[3];
  long long items = ((long long)batch + 3LL * batch * n_action_samples) * n_groups + 3LL * batch;
  D3B_REQUIRE(items < (1LL << 30), "cql_rows: too many rows for one launch");
  if (f32) launch_pdl(cql_rows_kernel<float>, dim3((unsigned)ceil_div_ll(items, 8)), dim3(256), 0, ST, p);
  else launch_pdl(cql_rows_kernel<__nv_bfloat16>, dim3((unsigned)ceil_div_ll(items, 8)), dim3(256), 0, ST, p);
  return check_launch("cql_rows");
}

extern "C" int d3b_cql_rows(const float* head, const float* obs, const float* next_obs, const float* act, int batch,
                            int n_action_samples, int obs_dim, int act_dim, float min_logstd, float max_logstd,
                            void* x_bf16, int64_t ldx, int n_groups, const void* const* ptrs_host,
                            const int64_t* rows_host, void* stream) {
  return cql_rows_any(false, head, obs, next_obs, act, batch, n_action_samples, obs_dim, act_dim, min_logstd, max_logstd,
                      x_bf16, ldx, n_groups, ptrs_host, rows_host, stream);
}
// the same rows as fp32 GEMM operands (fp32 mode: 3xTF32 / SIMT dense layers)
extern "C" int d3b_cql_rows_f32(const float* head, const float* obs, const float* next_obs, const float* act, int batch,
                                int n_action_samples, int obs_dim, int act_dim, float min_logstd, float max_logstd,
                                float* x, int64_t ldx, int n_groups, const void* const* ptrs_host,
                                const int64_t* rows_host, void* stream) {
  return cql_rows_any(true, head, obs, next_obs, act, batch, n_action_samples, obs_dim, act_dim, min_logstd, max_logstd,
                      x, ldx, n_groups, ptrs_host, rows_host, stream);
}

extern "C" int d3b_sac_temp_step(const float* logp, float* scalar, const int* step, int batch, int act_dim,
                                 float inv_batch, double lr, float* metric_loss, float* metric_exp, void* stream) {
  D3B_REQUIRE(logp && scalar && step && metric_loss && metric_exp && batch >= 1, "sac_temp_step: bad arguments");
  launch_pdl(sac_temp_step_kernel, dim3(1), dim3(1024), 0, ST, logp, scalar, step, batch, act_dim, inv_batch, lr, metric_loss,
             metric_exp);
  return check_launch("sac_temp_step");
}

extern "C" int d3b_cql_loss_step(const float* q, int64_t stride_q, const float* q_targ, int64_t stride_qt,
                                 int targ_members, const float* q_tpn, const float* rewards, const float* terminals,
                                 const float* n_steps, float gamma, const float* logp_t, const float* logp_tp1,
                                 int n_action_samples, int act_dim, float* scalar_alpha, float conservative_weight,
                                 float alpha_threshold, float* dq, int64_t stride_dq, float* sums, void* done_counter,
                                 int batch, int members, float inv_batch, int mode, const int* step_alpha,
                                 double lr_alpha, float* metric, float* metric_exp, void* stream) {
  D3B_REQUIRE(batch >= 1 && members >= 1 && n_action_samples >= 0, "cql_loss_step: bad sizes");
  D3B_REQUIRE(q && sums && done_counter && metric, "cql_loss_step: null pointer");
  D3B_REQUIRE(n_action_samples == 0 || (logp_t && logp_tp1 && scalar_alpha), "cql_loss_step: null conservative-term pointer");
  D3B_REQUIRE(mode == 0 || (step_alpha && metric_exp), "cql_loss_step: alpha mode needs the step counter / metric slot");
  D3B_REQUIRE(!(q_targ || q_tpn) || (rewards && terminals && n_steps), "cql_loss_step: TD term needs the minibatch");
  LossParams p{};
  p.q = q; p.sQ = stride_q; p.q_targ = q_targ; p.sQt = stride_qt; p.Et = targ_members; p.q_tpn = q_tpn;
  p.rew = rewards; p.term = terminals; p.nsteps = n_steps; p.gamma = gamma;
  p.logp_t = logp_t; p.logp_tp1 = logp_tp1; p.N = n_action_samples; p.A = act_dim;
  p.scalar_alpha = scalar_alpha; p.cw = conservative_weight; p.threshold = alpha_threshold;
  p.dq = dq; p.sDq = stride_dq; p.sums = sums; p.done = (unsigned*)done_counter;
  p.B = batch; p.E = members; p.inv_b = inv_batch; p.inv_eb = inv_batch / (float)members;
  p.mode = mode; p.step_alpha = step_alpha; p.lr_alpha = lr_alpha; p.metric = metric; p.metric_exp = metric_exp;
  launch_pdl(cql_loss_step_kernel, dim3(ceil_div(batch * members, 8)), dim3(256), 0, ST, p);
  return check_launch("cql_loss_step");
}

extern "C" int d3b_sac_actor_step(const float* q, int64_t stride_q, const float* logp, const float* log_temp,
                                  float* dq, int64_t stride_dq, float* loss_sum, void* done_counter, float* metric,
                                  int batch, int members, float inv_batch, void* stream) {
  D3B_REQUIRE(batch >= 1 && members >= 1, "sac_actor_step: bad sizes");
  D3B_REQUIRE(q && logp && log_temp && dq && loss_sum && done_counter && metric, "sac_actor_step: null pointer");
  launch_pdl(sac_actor_step_kernel, dim3(ceil_div(batch, 256)), dim3(256), 0, ST, q, (long long)stride_q, logp, log_temp, dq,
             (long long)stride_dq, loss_sum, (unsigned*)done_counter, metric, batch, members, inv_batch);
  return check_launch("sac_actor_step");
}
