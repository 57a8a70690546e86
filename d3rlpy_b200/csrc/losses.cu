// K4-K7: row-assembly, sampling and loss kernels (forward value + hand-written gradient seed) for
// TD3+BC, SAC/CQL.  Warp-shuffle reductions, no autograd.  Reference semantics restated:
//   tanh-Gaussian sample/log-prob   d3rlpy/models/torch/distributions.py:91-143, policies.py:16-23
//   TD target / TD error            d3rlpy/algos/torch/ddpg_impl.py:154-165,
//                                   d3rlpy/models/torch/q_functions/mean_q_function.py:74-87,
//                                   ensemble_q_function.py:81-106 (sum over members of batch means)
//   TD3 target smoothing            d3rlpy/algos/torch/td3_impl.py:61-78
//   TD3+BC actor loss               d3rlpy/algos/torch/td3_plus_bc_impl.py:64-70
//   SAC actor / temperature         d3rlpy/algos/torch/sac_impl.py:114-146
//   CQL conservative term           d3rlpy/algos/torch/cql_impl.py:143-223
// Gradient formulas: SURVEY.md Appendix C.
#include <cuda_bf16.h>

#include "common.cuh"

namespace d3b {

__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float softplusf(float x) {  // F.softplus, threshold 20
  return x > 20.f ? x : log1pf(expf(x));
}

// X[(b*N+k)] = [ obs[b] | f(act[b*N+k]) ];  f = optional TD3 smoothing or symmetric clip.
// One warp per output row.
template <typename Out>
__global__ void __launch_bounds__(256) concat_rows_kernel(const float* __restrict__ obs, long long ldo,
                                                          const float* __restrict__ act, long long lda,
                                                          const float* __restrict__ noise, float sigma,
                                                          float noise_clip, float act_clip, Out* __restrict__ X,
                                                          long long ldx, int B, int N, int O, int A) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * N) return;
  int b = (int)(row / N);
  Out* x = X + row * ldx;
  const float* o = obs + (long long)b * ldo;
  for (int j = lane; j < O; j += 32) put(x + j, __ldg(o + j));
  if (act) {
    const float* a = act + row * lda;
    for (int j = lane; j < A; j += 32) {
      float v = __ldg(a + j);
      if (noise) {
        float nz = clampf(sigma * __ldg(noise + row * A + j), -noise_clip, noise_clip);
        v = clampf(v + nz, -1.f, 1.f);
      }
      if (act_clip > 0.f) v = clampf(v, -act_clip, act_clip);
      put(x + O + j, v);
    }
  }
}

// head: [B][2A] = mu | raw logstd.  eps: [N][B][A] (rsample((n,)) layout, policies.py:202-213).
// Output row b*N+k: X = [obs[b] | tanh(mu + exp(clamp(logstd)) eps)], logp[b*N+k].
__global__ void __launch_bounds__(256) policy_sample_rows_kernel(
    const float* __restrict__ head, long long ldh, const float* __restrict__ eps, const float* __restrict__ obs,
    long long ldo, float* __restrict__ X, long long ldx, float* __restrict__ act_out, float* __restrict__ logp,
    int B, int N, int O, int A, float min_logstd, float max_logstd, int deterministic) {
  int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)B * N) return;
  int b = (int)(row / N), k = (int)(row % N);
  if (obs && X) {
    const float* o = obs + (long long)b * ldo;
    float* x = X + row * ldx;
    for (int j = lane; j < O; j += 32) x[j] = __ldg(o + j);
  }
  float lp = 0.f;
  for (int j = lane; j < A; j += 32) {
    float mu = __ldg(head + (long long)b * ldh + j);
    float a;
    if (deterministic) {
      a = tanhf(mu);
    } else {
      float ls = clampf(__ldg(head + (long long)b * ldh + A + j), min_logstd, max_logstd);
      float sd = expf(ls);
      float e = __ldg(eps + ((long long)k * B + b) * A + j);
      float u = mu + e * sd;
      a = tanhf(u);
      // Normal.log_prob(u) = -(u-mu)^2/(2 sd^2) - log(sd) - log(sqrt(2 pi)); jacobian 2(log2 - u - softplus(-2u))
      float d = u - mu;
      float nlp = -(d * d) / (2.f * sd * sd) - ls - 0.91893853320467267f;
      float jac = 2.f * (0.69314718055994529f - u - softplusf(-2.f * u));
      lp += nlp - jac;
    }
    if (X) X[row * ldx + O + j] = a;
    if (act_out) act_out[row * A + j] = a;
  }
  if (logp && !deterministic) {
    lp = warp_sum(lp);
    if (lane == 0) logp[row] = lp;
  }
}

// gamma ** n_steps as torch does for a python-float base and float32 exponent tensor.
__device__ __forceinline__ float gamma_pow(float gamma, float n) { return n == 1.f ? gamma : powf(gamma, n); }

// One thread per (member e, batch row b).  Rows of q: [data B | pi(s_t) B*N | pi(s_t+1) B*N | random B*N].
// sums[0] += sum_e sum_b (q-y)^2 ; sums[1] += sum_e sum_b logsumexp ; sums[2] += sum_e sum_b q_data
__global__ void __launch_bounds__(256) critic_loss_kernel(
    const float* __restrict__ q, long long sQ, const float* __restrict__ q_targ, long long sQt, int Et,
    const float* __restrict__ q_tpn_in, const float* __restrict__ rew, const float* __restrict__ term,
    const float* __restrict__ nsteps, float gamma, const float* __restrict__ logp_t,
    const float* __restrict__ logp_tp1, int N, int A, const float* __restrict__ log_alpha, float cw,
    float* __restrict__ dq, long long sDq, float* __restrict__ sums, float* __restrict__ y_out, int B, int E,
    float inv_b, float inv_eb, int td_enabled) {
  pdl_trigger();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float td = 0.f, lse_v = 0.f, data_v = 0.f;
  if (idx < B * E) {
    int e = idx / B, b = idx % B;
    const float* qe = q + (long long)e * sQ;
    float* dqe = dq ? dq + (long long)e * sDq : nullptr;
    float qd = __ldg(qe + b);
    float g = 0.f;
    if (td_enabled) {
      float tq;
      if (q_tpn_in) {
        tq = __ldg(q_tpn_in + b);
      } else {
        tq = __ldg(q_targ + b);
        for (int i = 1; i < Et; ++i) tq = fminf(tq, __ldg(q_targ + (long long)i * sQt + b));
      }
      float y = __ldg(rew + b) + gamma_pow(gamma, __ldg(nsteps + b)) * tq * (1.f - __ldg(term + b));
      if (y_out && e == 0) y_out[b] = y;
      float d = qd - y;
      td = d * d;
      g = 2.f * inv_b * d;
    }
    if (N > 0) {
      float ca = clampf(expf(__ldg(log_alpha)), 0.f, 1e6f);
      float c = ca * cw * inv_eb;
      float rand_lp = (float)A * -0.69314718055994529f;  // log(0.5^A)
      const float* q1 = qe + B + (long long)b * N;
      const float* q2 = q1 + (long long)B * N;
      const float* q3 = q2 + (long long)B * N;
      const float* l1 = logp_t + (long long)b * N;
      const float* l2 = logp_tp1 + (long long)b * N;
      float mx = -INFINITY;
      for (int k = 0; k < N; ++k) {
        mx = fmaxf(mx, __ldg(q1 + k) - __ldg(l1 + k));
        mx = fmaxf(mx, __ldg(q2 + k) - __ldg(l2 + k));
        mx = fmaxf(mx, __ldg(q3 + k) - rand_lp);
      }
      float s = 0.f;
      for (int k = 0; k < N; ++k) {
        s += expf(__ldg(q1 + k) - __ldg(l1 + k) - mx);
        s += expf(__ldg(q2 + k) - __ldg(l2 + k) - mx);
        s += expf(__ldg(q3 + k) - rand_lp - mx);
      }
      lse_v = mx + logf(s);
      data_v = qd;
      g -= c;
      if (dqe) {
        float inv_s = c / s;
        float* d1 = dqe + B + (long long)b * N;
        float* d2 = d1 + (long long)B * N;
        float* d3 = d2 + (long long)B * N;
        for (int k = 0; k < N; ++k) {
          d1[k] = expf(__ldg(q1 + k) - __ldg(l1 + k) - mx) * inv_s;
          d2[k] = expf(__ldg(q2 + k) - __ldg(l2 + k) - mx) * inv_s;
          d3[k] = expf(__ldg(q3 + k) - rand_lp - mx) * inv_s;
        }
      }
    }
    if (dqe) dqe[b] = g;
  }
  td = block_sum(td);
  if (threadIdx.x == 0 && td_enabled) atomicAdd(sums + 0, td);
  if (N > 0) {
    lse_v = block_sum(lse_v);
    if (threadIdx.x == 0) atomicAdd(sums + 1, lse_v);
    data_v = block_sum(data_v);
    if (threadIdx.x == 0) atomicAdd(sums + 2, data_v);
  }
}

// mode 0: critic_loss metric.  mode 1: alpha loss (+ d/dlog_alpha), cql_impl.py:119-141.
__global__ void cql_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ log_alpha, float inv_b,
                                    float inv_eb, float cw, float threshold, int mode, int conservative,
                                    float* __restrict__ metric, float* __restrict__ grad_log_alpha) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float td = sums[0] * inv_b;
  if (!conservative) {
    *metric = td;
    return;
  }
  float ea = expf(*log_alpha);
  float ca = clampf(ea, 0.f, 1e6f);
  float scaled = cw * (sums[1] * inv_eb - sums[2] * inv_eb);
  float cons = ca * (scaled - threshold);
  if (mode == 0) {
    *metric = td + cons;
  } else {
    *metric = -cons;
    float inside = (ea >= 0.f && ea <= 1e6f) ? 1.f : 0.f;
    if (grad_log_alpha) *grad_log_alpha = -inside * ea * (scaled - threshold);
  }
}

// Adam on a single scalar parameter (log_temp / log_alpha) + exp(p) written to a metric slot.
__global__ void scalar_adam_kernel(float* p, float* g, float* m, float* v, const int* step, double lr, double b1,
                                   double b2, double eps, float* out_exp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int t = *step;
  double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
  float w1 = (float)(1.0 - b1), fb2 = (float)b2, w2 = (float)(1.0 - b2);
  float G = *g, M = *m, V = *v;
  M = __fmaf_rn(w1, __fsub_rn(G, M), M);
  V = __fmul_rn(V, fb2);
  V = __fadd_rn(V, __fmul_rn(__fmul_rn(w2, G), G));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(V), (float)sqrt(bc2)), (float)eps);
  float P = __fadd_rn(*p, __fdiv_rn(__fmul_rn((float)(-(lr / bc1)), M), denom));
  *p = P;
  *m = M;
  *v = V;
  *g = 0.f;
  if (out_exp) *out_exp = expf(P);
}

// SAC actor: loss = mean_b(exp(log_temp) logp_b - min_e Q_e), dQ routed to the arg-min member (first on ties).
__global__ void __launch_bounds__(256) sac_actor_loss_kernel(const float* __restrict__ q, long long sQ,
                                                             const float* __restrict__ logp,
                                                             const float* __restrict__ log_temp,
                                                             float* __restrict__ dq, long long sDq,
                                                             float* __restrict__ loss_sum, int B, int E, float inv_b) {
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
  if (threadIdx.x == 0) atomicAdd(loss_sum, l);
}

// d(head) from dL/da (summed over members) and the entropy term; SURVEY.md Appendix C.4.
__global__ void __launch_bounds__(256) sac_actor_backward_kernel(
    const float* __restrict__ head, long long ldh, const float* __restrict__ eps, const float* __restrict__ dX,
    long long lddx, long long sdX, int E, const float* __restrict__ log_temp, float* __restrict__ dhead,
    long long lddh, int B, int A, float min_logstd, float max_logstd, float inv_b) {
  pdl_trigger();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * A) return;
  int b = idx / A, j = idx % A;
  float mu = __ldg(head + (long long)b * ldh + j);
  float raw_ls = __ldg(head + (long long)b * ldh + A + j);
  float ls = clampf(raw_ls, min_logstd, max_logstd);
  float sd = expf(ls);
  float e = __ldg(eps + (long long)b * A + j);
  float u = mu + e * sd;
  float a = tanhf(u);
  float da = 0.f;
  for (int m = 0; m < E; ++m) da += __ldg(dX + (long long)m * sdX + (long long)b * lddx + j);
  float at = expf(__ldg(log_temp)) * inv_b;
  float du = da * (1.f - a * a) + at * 2.f * a;
  float inside = (raw_ls >= min_logstd && raw_ls <= max_logstd) ? 1.f : 0.f;
  dhead[(long long)b * lddh + j] = du;
  dhead[(long long)b * lddh + A + j] = inside * (du * sd * e - at);
}

// temperature loss: -(exp(log_temp) * (logp - A)).mean(); d/dlog_temp equals the loss value.
__global__ void __launch_bounds__(1024) sac_temp_loss_kernel(const float* __restrict__ logp,
                                                             const float* __restrict__ log_temp, int B, int A,
                                                             float inv_b, float* __restrict__ metric,
                                                             float* __restrict__ grad, int accumulate) {
  pdl_trigger();
  pdl_wait();
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += __ldg(logp + b) - (float)A;
  s = block_sum(s);
  if (threadIdx.x == 0) {
    float l = -expf(*log_temp) * s * inv_b;
    if (accumulate) { *metric += l; *grad += l; } else { *metric = l; *grad = l; }
  }
}

// SACImpl.compute_target (sac_impl.py:148-162): q_tpn[b] = min_e Q'_e(s', a') - exp(log_temp) * logp(a'|s')
__global__ void __launch_bounds__(256) sac_soft_backup_kernel(const float* __restrict__ q_targ, long long sQ, int E,
                                                              const float* __restrict__ logp,
                                                              const float* __restrict__ log_temp,
                                                              float* __restrict__ q_tpn, int B) {
  pdl_trigger();
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = __ldg(q_targ + b);
  for (int e = 1; e < E; ++e) m = fminf(m, __ldg(q_targ + (long long)e * sQ + b));
  q_tpn[b] = m - expf(__ldg(log_temp)) * __ldg(logp + b);
}

// TD3+BC actor, phase 1: sums[0] += sum|q0|, sums[1] += sum q0, sums[2] += sum (a_data - a)^2
__global__ void __launch_bounds__(256) td3bc_actor_stats_kernel(const float* __restrict__ q0,
                                                                const float* __restrict__ a, long long lda,
                                                                const float* __restrict__ a_data, long long ldd,
                                                                float* __restrict__ sums, int B, int A) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  if (b < B) {
    float v = __ldg(q0 + b);
    s0 = fabsf(v);
    s1 = v;
    for (int j = 0; j < A; ++j) {
      float d = __ldg(a_data + (long long)b * ldd + j) - __ldg(a + (long long)b * lda + j);
      s2 += d * d;
    }
  }
  s0 = block_sum(s0);
  if (threadIdx.x == 0) atomicAdd(sums + 0, s0);
  s1 = block_sum(s1);
  if (threadIdx.x == 0) atomicAdd(sums + 1, s1);
  s2 = block_sum(s2);
  if (threadIdx.x == 0) atomicAdd(sums + 2, s2);
}

// phase 2: lam = alpha / mean|q0| (detached); dq[0][b] = -lam/B, other members 0; metric = loss.
// `sums` must already hold the GLOBAL (all-rank) sums when the batch is sharded.
__global__ void __launch_bounds__(256) td3bc_actor_seed_kernel(const float* __restrict__ sums, float alpha,
                                                               float inv_b, float inv_ba, float* __restrict__ dq,
                                                               long long sDq, int B, int E,
                                                               float* __restrict__ metric) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float lam = alpha / (sums[0] * inv_b);
  if (b < B) {
    dq[b] = -lam * inv_b;
    for (int e = 1; e < E; ++e) dq[(long long)e * sDq + b] = 0.f;
  }
  if (b == 0) *metric = lam * -(sums[1] * inv_b) + sums[2] * inv_ba;
}

// phase 3: dz = (dL/da_from_Q + 2 (a - a_data)/(B A)) * (1 - a^2)   (a = tanh(z))
__global__ void __launch_bounds__(256) td3bc_actor_backward_kernel(const float* __restrict__ a, long long lda,
                                                                   const float* __restrict__ a_data, long long ldd,
                                                                   const float* __restrict__ dX, long long lddx,
                                                                   float* __restrict__ dz, long long lddz, int B,
                                                                   int A, float inv_ba) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * A) return;
  int b = idx / A, j = idx % A;
  float av = __ldg(a + (long long)b * lda + j);
  float da = __ldg(dX + (long long)b * lddx + j) + 2.f * inv_ba * (av - __ldg(a_data + (long long)b * ldd + j));
  dz[(long long)b * lddz + j] = da * (1.f - av * av);
}

// StandardScaler.transform on a host-staged batch: x = (x - mean) / (std + eps)
__global__ void __launch_bounds__(256) standardize_kernel(float* __restrict__ x, const float* __restrict__ mean,
                                                          const float* __restrict__ std, float eps, long long total,
                                                          int dim) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int j = (int)(i % dim);
  x[i] = __fdiv_rn(__fsub_rn(x[i], __ldg(mean + j)), __fadd_rn(__ldg(std + j), eps));
}

// MinMaxActionScaler.transform / reverse_transform, operator order of action_scalers.py:185-206
template <bool REVERSE>
__global__ void __launch_bounds__(256) action_scale_kernel(float* __restrict__ a, const float* __restrict__ mn,
                                                            const float* __restrict__ mx, long long total, int dim) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int j = (int)(i % dim);
  float lo = __ldg(mn + j), range = __fsub_rn(__ldg(mx + j), lo);
  if (REVERSE)
    a[i] = __fadd_rn(__fmul_rn(range, __fdiv_rn(__fadd_rn(a[i], 1.0f), 2.0f)), lo);
  else
    a[i] = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(a[i], lo), range), 2.0f), 1.0f);
}

// RewardScaler.transform family: (mul * (clamp(r, lo, hi) - sub)) / div; torch.clamp propagates NaN
__global__ void __launch_bounds__(256) reward_scale_kernel(float* __restrict__ r, int n, float lo, float hi, float sub,
                                                            float mul, float div) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = r[i];
  if (v == v) v = fminf(fmaxf(v, lo), hi);
  r[i] = __fdiv_rn(__fmul_rn(mul, __fsub_rn(v, sub)), div);
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

extern "C" int d3b_scale_actions(float* a, const float* minimum, const float* maximum, int rows, int dim,
                                 void* stream) {
  D3B_REQUIRE(rows >= 0 && dim >= 1, "scale_actions: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(a && minimum && maximum, "scale_actions: null pointer");
  long long total = (long long)rows * dim;
  action_scale_kernel<false><<<(unsigned)ceil_div_ll(total, 256), 256, 0, ST>>>(a, minimum, maximum, total, dim);
  return check_launch("scale_actions");
}

extern "C" int d3b_unscale_actions(float* a, const float* minimum, const float* maximum, int rows, int dim,
                                   void* stream) {
  D3B_REQUIRE(rows >= 0 && dim >= 1, "unscale_actions: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(a && minimum && maximum, "unscale_actions: null pointer");
  long long total = (long long)rows * dim;
  action_scale_kernel<true><<<(unsigned)ceil_div_ll(total, 256), 256, 0, ST>>>(a, minimum, maximum, total, dim);
  return check_launch("unscale_actions");
}

extern "C" int d3b_scale_rewards(float* r, int n, float lo, float hi, float sub, float mul, float div, void* stream) {
  D3B_REQUIRE(n >= 0, "scale_rewards: bad size");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(r, "scale_rewards: null pointer");
  D3B_REQUIRE(lo <= hi && div != 0.0f, "scale_rewards: empty clip interval or zero divisor");
  reward_scale_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ST>>>(r, n, lo, hi, sub, mul, div);
  return check_launch("scale_rewards");
}

extern "C" int d3b_standardize(float* x, const float* mean, const float* std, float eps, int rows, int dim,
                               void* stream) {
  D3B_REQUIRE(rows >= 0 && dim >= 1, "standardize: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(x && mean && std, "standardize: null pointer");
  long long total = (long long)rows * dim;
  standardize_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, ST>>>(x, mean, std, eps, total, dim);
  return check_launch("standardize");
}

extern "C" int d3b_concat_rows(const float* obs, int64_t ldo, const float* act, int64_t lda, const float* noise,
                               float sigma, float noise_clip, float act_clip, float* x, int64_t ldx, int batch,
                               int n_repeat, int obs_dim, int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_repeat >= 1 && obs_dim >= 0 && act_dim >= 0, "concat_rows: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(obs && x, "concat_rows: null pointer");
  D3B_REQUIRE(ldx >= obs_dim + (act ? act_dim : 0), "concat_rows: ldx too small");
  long long rows = (long long)batch * n_repeat;
  concat_rows_kernel<float><<<(unsigned)ceil_div_ll(rows, 8), 256, 0, ST>>>(obs, ldo, act, lda, noise, sigma,
                                                                            noise_clip, act_clip, x, ldx, batch,
                                                                            n_repeat, obs_dim, act_dim);
  return check_launch("concat_rows");
}

// same rows written as bf16 (GEMM operands of the tensor-core path; no separate conversion launch)
extern "C" int d3b_concat_rows_bf16(const float* obs, int64_t ldo, const float* act, int64_t lda, const float* noise,
                                    float sigma, float noise_clip, float act_clip, void* x_bf16, int64_t ldx,
                                    int batch, int n_repeat, int obs_dim, int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_repeat >= 1 && obs_dim >= 0 && act_dim >= 0, "concat_rows_bf16: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(obs && x_bf16, "concat_rows_bf16: null pointer");
  D3B_REQUIRE(ldx >= obs_dim + (act ? act_dim : 0), "concat_rows_bf16: ldx too small");
  long long rows = (long long)batch * n_repeat;
  launch_pdl(concat_rows_kernel<__nv_bfloat16>, dim3((unsigned)ceil_div_ll(rows, 8)), dim3(256), 0, ST, obs,
             (long long)ldo, act, (long long)lda, noise, sigma, noise_clip, act_clip, (__nv_bfloat16*)x_bf16,
             (long long)ldx, batch, n_repeat, obs_dim, act_dim);
  return check_launch("concat_rows_bf16");
}

extern "C" int d3b_policy_sample_rows(const float* head, int64_t ld_head, const float* eps, const float* obs,
                                      int64_t ldo, float* x, int64_t ldx, float* act_out, float* logp, int batch,
                                      int n_samples, int obs_dim, int act_dim, float min_logstd, float max_logstd,
                                      int deterministic, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_samples >= 1 && act_dim >= 1, "policy_sample_rows: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(head && (deterministic || eps), "policy_sample_rows: null pointer");
  D3B_REQUIRE(ld_head >= 2 * act_dim || deterministic, "policy_sample_rows: head needs mu|logstd columns");
  long long rows = (long long)batch * n_samples;
  policy_sample_rows_kernel<<<(unsigned)ceil_div_ll(rows, 8), 256, 0, ST>>>(
      head, ld_head, eps, obs, ldo, x, ldx, act_out, logp, batch, n_samples, obs_dim, act_dim, min_logstd, max_logstd,
      deterministic);
  return check_launch("policy_sample_rows");
}

extern "C" int d3b_critic_loss(const float* q, int64_t stride_q, const float* q_targ, int64_t stride_qt,
                               int targ_members, const float* q_tpn, const float* rewards, const float* terminals,
                               const float* n_steps, float gamma, const float* logp_t, const float* logp_tp1,
                               int n_action_samples, int act_dim, const float* log_alpha, float conservative_weight,
                               float* dq, int64_t stride_dq, float* sums, float* y_out, int batch, int members,
                               float inv_batch, int td_enabled, void* stream) {
  D3B_REQUIRE(batch >= 0 && members >= 1 && n_action_samples >= 0, "critic_loss: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q && sums, "critic_loss: null pointer");
  D3B_REQUIRE(!td_enabled || ((q_tpn || q_targ) && rewards && terminals && n_steps),
              "critic_loss: TD term needs target, rewards, terminals, n_steps");
  D3B_REQUIRE(n_action_samples == 0 || (logp_t && logp_tp1 && log_alpha),
              "critic_loss: conservative term needs log-probs and log_alpha");
  int total = batch * members;
  launch_pdl(critic_loss_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, ST, 
      q, stride_q, q_targ, stride_qt, targ_members, q_tpn, rewards, terminals, n_steps, gamma, logp_t, logp_tp1,
      n_action_samples, act_dim, log_alpha, conservative_weight, dq, stride_dq, sums, y_out, batch, members, inv_batch,
      inv_batch / (float)members, td_enabled);
  return check_launch("critic_loss");
}

extern "C" int d3b_cql_finalize(const float* sums, const float* log_alpha, float inv_batch, int members,
                                float conservative_weight, float alpha_threshold, int mode, int conservative,
                                float* metric, float* grad_log_alpha, void* stream) {
  D3B_REQUIRE(sums && metric && (!conservative || log_alpha), "cql_finalize: null pointer");
  launch_pdl(cql_finalize_kernel, dim3(1), dim3(32), 0, ST, sums, log_alpha, inv_batch, inv_batch / (float)members, conservative_weight,
                                        alpha_threshold, mode, conservative, metric, grad_log_alpha);
  return check_launch("cql_finalize");
}

extern "C" int d3b_scalar_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const int* step,
                               double lr, double beta1, double beta2, double eps, float* out_exp, void* stream) {
  D3B_REQUIRE(param && grad && exp_avg && exp_avg_sq && step, "scalar_adam: null pointer");
  scalar_adam_kernel<<<1, 32, 0, ST>>>(param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, out_exp);
  return check_launch("scalar_adam");
}

extern "C" int d3b_sac_actor_loss(const float* q, int64_t stride_q, const float* logp, const float* log_temp,
                                  float* dq, int64_t stride_dq, float* loss_sum, int batch, int members,
                                  float inv_batch, void* stream) {
  D3B_REQUIRE(batch >= 0 && members >= 1, "sac_actor_loss: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q && logp && log_temp && dq && loss_sum, "sac_actor_loss: null pointer");
  launch_pdl(sac_actor_loss_kernel, dim3(ceil_div(batch, 256)), dim3(256), 0, ST, q, stride_q, logp, log_temp, dq, stride_dq, loss_sum,
                                                              batch, members, inv_batch);
  return check_launch("sac_actor_loss");
}

extern "C" int d3b_sac_actor_backward(const float* head, int64_t ld_head, const float* eps, const float* dx_action,
                                      int64_t lddx, int64_t stride_dx, int members, const float* log_temp,
                                      float* dhead, int64_t ld_dhead, int batch, int act_dim, float min_logstd,
                                      float max_logstd, float inv_batch, void* stream) {
  D3B_REQUIRE(batch >= 0 && act_dim >= 1 && members >= 1, "sac_actor_backward: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(head && eps && dx_action && log_temp && dhead, "sac_actor_backward: null pointer");
  launch_pdl(sac_actor_backward_kernel, dim3(ceil_div(batch * act_dim, 256)), dim3(256), 0, ST, head, ld_head, eps, dx_action,
             lddx, stride_dx, members, log_temp, dhead, ld_dhead, batch, act_dim, min_logstd, max_logstd, inv_batch);
  return check_launch("sac_actor_backward");
}

extern "C" int d3b_sac_temp_loss(const float* logp, const float* log_temp, int batch, int act_dim, float inv_batch,
                                 float* metric, float* grad, int accumulate, void* stream) {
  D3B_REQUIRE(logp && log_temp && metric && grad && batch >= 0, "sac_temp_loss: bad arguments");
  launch_pdl(sac_temp_loss_kernel, dim3(1), dim3(1024), 0, ST, logp, log_temp, batch, act_dim, inv_batch, metric, grad, accumulate);
  return check_launch("sac_temp_loss");
}

extern "C" int d3b_td3bc_actor_stats(const float* q0, const float* a, int64_t lda, const float* a_data, int64_t ldd,
                                     float* sums, int batch, int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 0 && act_dim >= 1, "td3bc_actor_stats: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q0 && a && a_data && sums, "td3bc_actor_stats: null pointer");
  td3bc_actor_stats_kernel<<<ceil_div(batch, 256), 256, 0, ST>>>(q0, a, lda, a_data, ldd, sums, batch, act_dim);
  return check_launch("td3bc_actor_stats");
}

extern "C" int d3b_td3bc_actor_seed(const float* sums, float alpha, float inv_batch, int act_dim, float* dq,
                                    int64_t stride_dq, int batch, int members, float* metric, void* stream) {
  D3B_REQUIRE(batch >= 1 && members >= 1 && act_dim >= 1, "td3bc_actor_seed: bad sizes");
  D3B_REQUIRE(sums && dq && metric, "td3bc_actor_seed: null pointer");
  td3bc_actor_seed_kernel<<<ceil_div(batch, 256), 256, 0, ST>>>(sums, alpha, inv_batch, inv_batch / (float)act_dim, dq,
                                                                stride_dq, batch, members, metric);
  return check_launch("td3bc_actor_seed");
}

extern "C" int d3b_td3bc_actor_backward(const float* a, int64_t lda, const float* a_data, int64_t ldd,
                                        const float* dx_action, int64_t lddx, float* dz, int64_t lddz, int batch,
                                        int act_dim, float inv_batch, void* stream) {
  D3B_REQUIRE(batch >= 0 && act_dim >= 1, "td3bc_actor_backward: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(a && a_data && dx_action && dz, "td3bc_actor_backward: null pointer");
  td3bc_actor_backward_kernel<<<ceil_div(batch * act_dim, 256), 256, 0, ST>>>(
      a, lda, a_data, ldd, dx_action, lddx, dz, lddz, batch, act_dim, inv_batch / (float)act_dim);
  return check_launch("td3bc_actor_backward");
}

extern "C" int d3b_sac_soft_backup(const float* q_targ, int64_t stride_q, int members, const float* logp,
                                   const float* log_temp, float* q_tpn, int batch, void* stream) {
  D3B_REQUIRE(batch >= 0 && members >= 1, "sac_soft_backup: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q_targ && logp && log_temp && q_tpn, "sac_soft_backup: null pointer");
  launch_pdl(sac_soft_backup_kernel, dim3(ceil_div(batch, 256)), dim3(256), 0, ST, q_targ, stride_q, members, logp, log_temp,
             q_tpn, batch);
  return check_launch("sac_soft_backup");
}
