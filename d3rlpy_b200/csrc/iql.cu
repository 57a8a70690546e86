// Implicit Q-Learning loss kernels (sibling algorithm on the update path's building blocks).  Replace
//   d3rlpy/algos/torch/iql_impl.py:130-141  compute_value_loss  (expectile regression of V(s) towards min_e Q'_e(s, a))
//   d3rlpy/algos/torch/iql_impl.py:109-128  compute_actor_loss / _compute_weight  (advantage-weighted log-likelihood)
//   d3rlpy/models/torch/policies.py:168-181,248-253 + distributions.py:33-88  (Normal(tanh(mu), exp(logstd)) with the
//     learnable logstd parameter squashed into [min_logstd, max_logstd] by a sigmoid)
// Both run as ONE block with a fixed summation order: the batch is at most a few thousand samples and the sums feed
// parameters (logstd gradient) and metrics that must be bit-reproducible across runs.
#include "common.cuh"

namespace d3b {

constexpr int IQL_MAX_A = 32;

__global__ void __launch_bounds__(256) iql_value_loss_kernel(const float* __restrict__ q_targ, long long sQ, int E,
                                                             const float* __restrict__ v, float expectile, float inv_b,
                                                             float* __restrict__ dv, float* __restrict__ metric,
                                                             int B) {
  float loss = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float q = INFINITY;
    for (int e = 0; e < E; ++e) q = fminf(q, __ldg(q_targ + (long long)e * sQ + b));
    float diff = q - __ldg(v + b);
    float w = fabsf(expectile - (diff < 0.f ? 1.f : 0.f));
    loss += w * diff * diff;
    dv[b] = -2.f * w * diff * inv_b;
  }
  loss = block_sum(loss);
  if (threadIdx.x == 0) *metric = loss * inv_b;
}

__global__ void __launch_bounds__(256) iql_actor_loss_kernel(
    const float* __restrict__ mu, long long ldmu, const float* __restrict__ logstd_param,
    const float* __restrict__ actions, long long lda, const float* __restrict__ q_targ, long long sQ, int E,
    const float* __restrict__ v, float weight_temp, float max_weight, float min_logstd, float max_logstd, float inv_b,
    float* __restrict__ dmu, long long lddmu, float* __restrict__ dlogstd, float* __restrict__ metric, int B, int A) {
  __shared__ float s_logstd[IQL_MAX_A], s_inv_var[IQL_MAX_A], s_dlog[IQL_MAX_A];
  const float range = max_logstd - min_logstd;
  if (threadIdx.x < A) {
    float s = 1.f / (1.f + expf(-__ldg(logstd_param + threadIdx.x)));
    float ls = min_logstd + s * range;       // get_logstd_parameter (policies.py:248-253)
    float sd = expf(ls);
    s_logstd[threadIdx.x] = logf(sd);        // Normal.log_prob uses scale.log()
    s_inv_var[threadIdx.x] = 1.f / (sd * sd);
    s_dlog[threadIdx.x] = range * s * (1.f - s);
  }
  __syncthreads();
  float loss = 0.f;
  float acc[IQL_MAX_A];
#pragma unroll
  for (int j = 0; j < IQL_MAX_A; ++j) acc[j] = 0.f;
  const float half_log_2pi = 0.91893853320467274178f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float q = INFINITY;
    for (int e = 0; e < E; ++e) q = fminf(q, __ldg(q_targ + (long long)e * sQ + b));
    float w = fminf(expf(weight_temp * (q - __ldg(v + b))), max_weight);
    float logp = 0.f;
    float c = -w * inv_b;                    // d(loss)/d(logp_b)
#pragma unroll
    for (int j = 0; j < IQL_MAX_A; ++j) {
      if (j < A) {
        float m = tanhf(__ldg(mu + (long long)b * ldmu + j));
        float d = __ldg(actions + (long long)b * lda + j) - m;
        float z = d * d * s_inv_var[j];
        logp += -0.5f * z - s_logstd[j] - half_log_2pi;
        dmu[(long long)b * lddmu + j] = c * (d * s_inv_var[j]) * (1.f - m * m);
        acc[j] += c * (z - 1.f);
      }
    }
    loss -= w * logp;
  }
  loss = block_sum(loss);
  if (threadIdx.x == 0) *metric = loss * inv_b;
#pragma unroll
  for (int j = 0; j < IQL_MAX_A; ++j) {
    if (j < A) {
      float g = block_sum(acc[j]);
      if (threadIdx.x == 0) dlogstd[j] += g * s_dlog[j];
    }
  }
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

extern "C" int d3b_iql_value_loss(const float* q_targ, int64_t stride_q, int members, const float* v, float expectile,
                                  float inv_batch, float* dv, float* metric, int batch, void* stream) {
  D3B_REQUIRE(batch >= 1 && members >= 1, "iql_value_loss: bad sizes");
  D3B_REQUIRE(q_targ && v && dv && metric, "iql_value_loss: null pointer");
  iql_value_loss_kernel<<<1, 256, 0, ST>>>(q_targ, stride_q, members, v, expectile, inv_batch, dv, metric, batch);
  return check_launch("iql_value_loss");
}

extern "C" int d3b_iql_actor_loss(const float* mu, int64_t ld_mu, const float* logstd_param, const float* actions,
                                  int64_t ld_act, const float* q_targ, int64_t stride_q, int members, const float* v,
                                  float weight_temp, float max_weight, float min_logstd, float max_logstd,
                                  float inv_batch, float* dmu, int64_t ld_dmu, float* dlogstd, float* metric, int batch,
                                  int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 1 && members >= 1, "iql_actor_loss: bad sizes");
  D3B_REQUIRE(act_dim >= 1 && act_dim <= IQL_MAX_A, "iql_actor_loss: act_dim %d not in 1..%d", act_dim, IQL_MAX_A);
  D3B_REQUIRE(mu && logstd_param && actions && q_targ && v && dmu && dlogstd && metric, "iql_actor_loss: null pointer");
  iql_actor_loss_kernel<<<1, 256, 0, ST>>>(mu, ld_mu, logstd_param, actions, ld_act, q_targ, stride_q, members, v,
                                           weight_temp, max_weight, min_logstd, max_logstd, inv_batch, dmu, ld_dmu,
                                           dlogstd, metric, batch, act_dim);
  return check_launch("iql_actor_loss");
}
