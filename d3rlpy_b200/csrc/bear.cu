// BEAR loss kernels (sibling algorithm on the update path's building blocks).  Replace
//   d3rlpy/algos/torch/bear_impl.py:233-281  _compute_mmd: MMD between n raw (pre-tanh) policy samples and n raw decoder
//     samples per observation, Laplacian or Gaussian kernel (bear_impl.py:27-38), sqrt(mmd + 1e-6)
//   bear_impl.py:186-190,215-231  _compute_mmd_loss / update_alpha: (exp(log_alpha) (mmd - threshold)).mean(), Lagrange
//     step on log_alpha with the clamp to [-5, 10]
//   bear_impl.py:283-303  compute_target: per sample the lam-mix of the target members at the best of n policy actions
//     (q_functions/__init__.py:8-63) minus exp(log_temp) times that action's log-probability
// The MMD kernel runs one thread per observation (n <= 16 samples of <= 32 action dimensions: a few thousand flops) and
// reduces its sums in a fixed order.
#include "common.cuh"

namespace d3b {

constexpr int BEAR_MAX_N = 16;
constexpr int BEAR_MAX_A = 32;

// x[k*B + b] = [obs_b | clamp(latent[k*B + b], +-clip)]   (rows in sample-major order like x.expand(n, ...) in
// ConditionalVAE.sample_n_without_squash, imitators.py:94-118)
__global__ void bear_latent_rows_kernel(const float* __restrict__ obs, long long ldo, const float* __restrict__ latent,
                                        float clip, float* __restrict__ x, long long ldx, int B, int n, int O, int Z) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = O + Z;
  if (i >= (long long)B * n * W) return;
  const long long r = i / W;
  const int j = (int)(i % W);
  const int b = (int)(r % B);
  float v;
  if (j < O) v = __ldg(obs + (long long)b * ldo + j);
  else v = fminf(fmaxf(__ldg(latent + r * Z + (j - O)), -clip), clip);
  x[r * ldx + j] = v;
}

__device__ __forceinline__ float mmd_k(const float* x, const float* y, int A, int gaussian, float sigma) {
  float s = 0.f;
  for (int a = 0; a < A; ++a) {
    const float d = x[a] - y[a];
    s += gaussian ? d * d : fabsf(d);
  }
  return expf(-s / (2.f * sigma));
}

// one thread per observation; sum_out[0] += sum_b (mmd_b - threshold) (block-ordered), optional gradient into d_head
__global__ void __launch_bounds__(128) bear_mmd_kernel(const float* __restrict__ head, long long ldh,
                                                       const float* __restrict__ eps, const float* __restrict__ beh,
                                                       long long ldb, int gaussian, float sigma, float min_ls,
                                                       float max_ls, const float* __restrict__ log_alpha,
                                                       float threshold, float inv_b, float* __restrict__ d_head,
                                                       long long lddh, float* __restrict__ sum_out, int B, int n,
                                                       int A) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (b < B) {
    float p[BEAR_MAX_N][BEAR_MAX_A];   // policy raw samples (local memory: a few KB per thread, one pass)
    float sd[BEAR_MAX_A], pass[BEAR_MAX_A];
    for (int a = 0; a < A; ++a) {
      const float raw = __ldg(head + (long long)b * ldh + A + a);
      sd[a] = expf(fminf(fmaxf(raw, min_ls), max_ls));
      pass[a] = (raw >= min_ls && raw <= max_ls) ? 1.f : 0.f;
      const float mu = __ldg(head + (long long)b * ldh + a);
      for (int i = 0; i < n; ++i) p[i][a] = mu + sd[a] * __ldg(eps + ((long long)i * B + b) * A + a);
    }
    auto behav = [&](int j) { return beh + ((long long)j * B + b) * ldb; };
    float bj[BEAR_MAX_A], bi[BEAR_MAX_A];
    float pp = 0.f, bb = 0.f, pb = 0.f;
    for (int i = 0; i < n; ++i) {
      for (int a = 0; a < A; ++a) bi[a] = __ldg(behav(i) + a);
      for (int j = 0; j < n; ++j) {
        for (int a = 0; a < A; ++a) bj[a] = __ldg(behav(j) + a);
        pp += mmd_k(p[i], p[j], A, gaussian, sigma);
        bb += mmd_k(bi, bj, A, gaussian, sigma);
        pb += mmd_k(p[i], bj, A, gaussian, sigma);
      }
    }
    const float inv_n2 = 1.f / (float)(n * n);
    const float mmd = sqrtf(pp * inv_n2 + bb * inv_n2 - 2.f * pb * inv_n2 + 1e-6f);
    contrib = mmd - threshold;
    if (d_head) {
      // d loss / d p_i = (alpha / B) / (2 mmd) * (2 / n^2) * sum_j [dk(p_i, p_j)/dp_i - dk(p_i, b_j)/dp_i]
      const float c = expf(__ldg(log_alpha)) * inv_b / (2.f * mmd) * 2.f * inv_n2;
      float dmu[BEAR_MAX_A], dls[BEAR_MAX_A];
      for (int a = 0; a < A; ++a) dmu[a] = dls[a] = 0.f;
      for (int i = 0; i < n; ++i) {
        float g[BEAR_MAX_A];
        for (int a = 0; a < A; ++a) g[a] = 0.f;
        for (int j = 0; j < n; ++j) {
          for (int a = 0; a < A; ++a) bj[a] = __ldg(behav(j) + a);
          const float kpp = mmd_k(p[i], p[j], A, gaussian, sigma), kpb = mmd_k(p[i], bj, A, gaussian, sigma);
          for (int a = 0; a < A; ++a) {
            const float dpp = p[i][a] - p[j][a], dpb = p[i][a] - bj[a];
            if (gaussian) {
              g[a] += -kpp * dpp / sigma + kpb * dpb / sigma;
            } else {
              const float spp = dpp > 0.f ? 1.f : (dpp < 0.f ? -1.f : 0.f), spb = dpb > 0.f ? 1.f : (dpb < 0.f ? -1.f : 0.f);
              g[a] += (-kpp * spp + kpb * spb) / (2.f * sigma);
            }
          }
        }
        for (int a = 0; a < A; ++a) {
          const float e = __ldg(eps + ((long long)i * B + b) * A + a);
          dmu[a] += c * g[a];
          dls[a] += c * g[a] * e * sd[a] * pass[a];
        }
      }
      for (int a = 0; a < A; ++a) {
        d_head[(long long)b * lddh + a] += dmu[a];
        d_head[(long long)b * lddh + A + a] += dls[a];
      }
    }
  }
  // fixed-order block sum, one atomic per block (B <= a few blocks; the order of the few atomics does not matter at
  // the 1e-5 tolerance and is deterministic for B <= 128)
  contrib = block_sum(contrib);
  if (threadIdx.x == 0) atomicAdd(sum_out, contrib);
}

__device__ __forceinline__ float bear_scalar_adam(float* p, float G, float* m, float* v, int t, double lr) {
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

// update_alpha: loss = -exp(log_alpha) * mean(mmd - threshold) (== its gradient w.r.t. log_alpha), Adam, clamp
// scalar block: {p, g, m, v} at float offsets 0, 4, 8, 12
__global__ void bear_alpha_step_kernel(const float* __restrict__ sum, float* scalar, const int* step, double lr,
                                       float inv_b, float* metric_loss, float* metric_alpha) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    const float loss = -expf(scalar[0]) * (*sum) * inv_b;
    *metric_loss = loss;
    float P = bear_scalar_adam(scalar + 0, loss, scalar + 8, scalar + 12, *step, lr);
    P = fminf(fmaxf(P, -5.f), 10.f);
    scalar[0] = P;
    *metric_alpha = expf(P);
  }
}

// actor_loss = [SAC actor loss] + exp(log_alpha) * mean(mmd - threshold)
__global__ void bear_actor_metric_kernel(const float* __restrict__ sac_loss, const float* __restrict__ mmd_sum,
                                         const float* __restrict__ log_alpha, float inv_b, float* metric) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) *metric = (sac_loss ? *sac_loss : 0.f) + expf(*log_alpha) * (*mmd_sum) * inv_b;
}

// q[E][B*n] (row b*n + k), logp[B*n] -> q_tpn[b] = mix(b, k*) - exp(log_temp) * logp[b*n + k*],
// mix = (1 - lam) max_e + lam min_e, k* = argmax_k mix (first maximum, like torch.argmax)
__global__ void bear_target_kernel(const float* __restrict__ q, long long sQ, const float* __restrict__ logp,
                                   const float* __restrict__ log_temp, float lam, float* __restrict__ q_tpn, int B,
                                   int n, int E) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float best = -INFINITY;
  int best_k = 0;
  for (int k = 0; k < n; ++k) {
    float mx = -INFINITY, mn = INFINITY;
    for (int e = 0; e < E; ++e) {
      const float v = __ldg(q + (long long)e * sQ + (long long)b * n + k);
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    const float mix = __fadd_rn(__fmul_rn(1.0f - lam, mx), __fmul_rn(lam, mn));
    if (mix > best) { best = mix; best_k = k; }
  }
  q_tpn[b] = best - expf(__ldg(log_temp)) * __ldg(logp + (long long)b * n + best_k);
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

extern "C" int d3b_bear_latent_rows(const float* obs, int64_t ldo, const float* latent, float clip, float* x, int64_t ldx,
                                    int batch, int n, int obs_dim, int latent_dim, void* stream) {
  D3B_REQUIRE(batch >= 1 && n >= 1 && obs_dim >= 0 && latent_dim >= 1, "bear_latent_rows: bad sizes");
  D3B_REQUIRE(latent && x && (obs || obs_dim == 0), "bear_latent_rows: null pointer");
  const long long total = (long long)batch * n * (obs_dim + latent_dim);
  launch_pdl(bear_latent_rows_kernel, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, ST, obs, (long long)ldo,
             latent, clip, x, (long long)ldx, batch, n, obs_dim, latent_dim);
  return check_launch("bear_latent_rows");
}

extern "C" int d3b_bear_mmd(const float* head, int64_t ld_head, const float* eps, const float* behavior_raw,
                            int64_t ld_behavior, int gaussian_kernel, float sigma, float min_logstd, float max_logstd,
                            const float* log_alpha, float threshold, float inv_batch, float* d_head, int64_t ld_dhead,
                            float* sum_out, int batch, int n, int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 1 && n >= 1 && n <= BEAR_MAX_N && act_dim >= 1 && act_dim <= BEAR_MAX_A && sigma > 0.f,
              "bear_mmd: bad sizes (n <= %d, act_dim <= %d)", BEAR_MAX_N, BEAR_MAX_A);
  D3B_REQUIRE(head && eps && behavior_raw && sum_out && (!d_head || log_alpha), "bear_mmd: null pointer");
  launch_pdl(bear_mmd_kernel, dim3(ceil_div(batch, 128)), dim3(128), 0, ST, head, (long long)ld_head, eps, behavior_raw,
             (long long)ld_behavior, gaussian_kernel, sigma, min_logstd, max_logstd, log_alpha, threshold, inv_batch,
             d_head, (long long)ld_dhead, sum_out, batch, n, act_dim);
  return check_launch("bear_mmd");
}

extern "C" int d3b_bear_alpha_step(const float* mmd_sum, float* alpha_scalar, const int* step, double lr, float inv_batch,
                                   float* metric_loss, float* metric_alpha, void* stream) {
  D3B_REQUIRE(mmd_sum && alpha_scalar && step && metric_loss && metric_alpha, "bear_alpha_step: null pointer");
  launch_pdl(bear_alpha_step_kernel, dim3(1), dim3(32), 0, ST, mmd_sum, alpha_scalar, step, lr, inv_batch, metric_loss,
             metric_alpha);
  return check_launch("bear_alpha_step");
}

extern "C" int d3b_bear_actor_metric(const float* sac_loss, const float* mmd_sum, const float* log_alpha, float inv_batch,
                                     float* metric, void* stream) {
  D3B_REQUIRE(mmd_sum && log_alpha && metric, "bear_actor_metric: null pointer");
  launch_pdl(bear_actor_metric_kernel, dim3(1), dim3(32), 0, ST, sac_loss, mmd_sum, log_alpha, inv_batch, metric);
  return check_launch("bear_actor_metric");
}

extern "C" int d3b_bear_target(const float* q, int64_t stride_q, const float* logp, const float* log_temp, float lam,
                               float* q_tpn, int batch, int n, int members, void* stream) {
  D3B_REQUIRE(batch >= 1 && n >= 1 && members >= 1, "bear_target: bad sizes");
  D3B_REQUIRE(q && logp && log_temp && q_tpn, "bear_target: null pointer");
  launch_pdl(bear_target_kernel, dim3(ceil_div(batch, 128)), dim3(128), 0, ST, q, (long long)stride_q, logp, log_temp, lam,
             q_tpn, batch, n, members);
  return check_launch("bear_target");
}
