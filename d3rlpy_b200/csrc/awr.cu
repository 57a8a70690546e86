// Advantage-weighted actor steps over a non-squashed Gaussian policy (siblings of the update path: AWAC, CRR).
//   d3rlpy/models/torch/policies.py:160-181,248-253 + distributions.py:33-88: dist = Normal(tanh(mu), exp(logstd)) with
//     logstd either a learnable parameter squashed by a sigmoid into [min, max] (AWAC) or a clamped head (CRR);
//     sample = clamp(loc + scale * eps, -1, 1)
//   d3rlpy/algos/torch/awac_impl.py:103-154: weights = softmax_batch((min_e Q(s,a) - mean_n min_e Q(s,a_n)) / lam) * B,
//     loss = -sum(log pi(a|s) * w)
//   d3rlpy/algos/torch/crr_impl.py:82-141: advantage with the member MEAN, state value = mean or max over n samples,
//     weights = clamp(exp(adv / beta), 0, max_weight) or [adv > 0], loss = -mean(log pi(a|s) * w)
// The weight / loss kernels run as ONE block with a fixed summation order (the batch is at most a few thousand rows;
// the sums feed parameters and metrics that must be bit-reproducible).
#include "common.cuh"

namespace d3b {

constexpr int AWR_MAX_A = 32;

__device__ __forceinline__ float block_max(float v) {
  __shared__ float redm[32];
  v = warp_max(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) redm[w] = v;
  __syncthreads();
  v = (lane < (int)((blockDim.x + 31) >> 5)) ? redm[lane] : -INFINITY;
  v = warp_max(v);
  return __shfl_sync(0xffffffffu, v, 0);   // every thread gets the maximum
}

// x[b*n + k] = [obs_b | clamp(tanh(mu_b) + exp(logstd) * eps[k][b], -1, 1)]   (eps laid out (n, B, A) like rsample((n,)))
__global__ void gauss_policy_rows_kernel(const float* __restrict__ head, long long ldh,
                                         const float* __restrict__ logstd_param, float min_ls, float max_ls,
                                         const float* __restrict__ eps, const float* __restrict__ obs, long long ldo,
                                         float* __restrict__ x, long long ldx, int B, int n, int O, int A) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * n) return;
  const int b = warp / n, k = warp % n;
  float* xr = x + (long long)warp * ldx;
  for (int j = lane; j < O; j += 32) xr[j] = __ldg(obs + (long long)b * ldo + j);
  for (int j = lane; j < A; j += 32) {
    const float loc = tanhf(__ldg(head + (long long)b * ldh + j));
    float ls;
    if (logstd_param) {
      const float s = 1.f / (1.f + expf(-__ldg(logstd_param + j)));
      ls = min_ls + s * (max_ls - min_ls);
    } else {
      ls = fminf(fmaxf(__ldg(head + (long long)b * ldh + A + j), min_ls), max_ls);
    }
    const float a = loc + expf(ls) * __ldg(eps + ((long long)k * B + b) * A + j);
    xr[O + j] = fminf(fmaxf(a, -1.f), 1.f);
  }
}

// per-row advantage weights (see the header comment)
__global__ void __launch_bounds__(256) awr_weights_kernel(const float* __restrict__ q_data, long long sQd,
                                                          const float* __restrict__ q_samp, long long sQs, int E, int B,
                                                          int n, int member_reduce, int value_reduce, int weight_mode,
                                                          float temperature, float max_weight,
                                                          float* __restrict__ weights) {
  pdl_trigger();
  pdl_wait();
  const float inv_e = 1.f / (float)E;
  auto reduce_members = [&](const float* q, long long sQ, long long i) {
    float v = __ldg(q + i);
    for (int e = 1; e < E; ++e) {
      float u = __ldg(q + (long long)e * sQ + i);
      v = member_reduce == 0 ? fminf(v, u) : v + u;
    }
    return member_reduce == 0 ? v : v * inv_e;
  };
  float local_max = -INFINITY;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float val = value_reduce == 0 ? 0.f : -INFINITY;
    for (int k = 0; k < n; ++k) {
      float q = reduce_members(q_samp, sQs, (long long)b * n + k);
      val = value_reduce == 0 ? val + q : fmaxf(val, q);
    }
    if (value_reduce == 0) val /= (float)n;
    const float adv = reduce_members(q_data, sQd, b) - val;
    float w;
    if (weight_mode == 0) { w = adv / temperature; local_max = fmaxf(local_max, w); }
    else if (weight_mode == 1) w = fminf(fmaxf(expf(adv / temperature), 0.f), max_weight);
    else w = adv > 0.f ? 1.f : 0.f;
    weights[b] = w;
  }
  if (weight_mode != 0) return;
  // softmax over the batch, times the batch size
  __syncthreads();
  const float mx = block_max(local_max);
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float ex = expf(weights[b] - mx);
    weights[b] = ex;
    s += ex;
  }
  __shared__ float total;
  s = block_sum(s);
  if (threadIdx.x == 0) total = s;
  __syncthreads();
  const float inv = 1.f / total;
  for (int b = threadIdx.x; b < B; b += blockDim.x) weights[b] = weights[b] * inv * (float)B;
}

// loss = -scale * sum_b w_b log N(a_b; tanh(mu_b), exp(logstd)^2), gradient seeds for the policy head(s)
__global__ void __launch_bounds__(256) gauss_wll_loss_kernel(
    const float* __restrict__ head, long long ldh, const float* __restrict__ logstd_param,
    const float* __restrict__ actions, long long lda, const float* __restrict__ weights, float min_ls, float max_ls,
    float scale, float* __restrict__ d_head, long long lddh, float* __restrict__ dlogstd_param,
    float* __restrict__ metric_loss, float* __restrict__ metric_mean_std, int B, int A) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_ls[AWR_MAX_A], s_dls[AWR_MAX_A];
  const float range = max_ls - min_ls;
  if (logstd_param && threadIdx.x < A) {
    const float s = 1.f / (1.f + expf(-__ldg(logstd_param + threadIdx.x)));
    s_ls[threadIdx.x] = min_ls + s * range;          // get_logstd_parameter (policies.py:248-253)
    s_dls[threadIdx.x] = range * s * (1.f - s);
  }
  __syncthreads();
  float loss = 0.f;
  float acc[AWR_MAX_A];
#pragma unroll
  for (int j = 0; j < AWR_MAX_A; ++j) acc[j] = 0.f;
  const float half_log_2pi = 0.91893853320467274178f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float w = __ldg(weights + b);
    const float c = -w * scale;                      // d(loss) / d(logp_b)
    float logp = 0.f;
#pragma unroll
    for (int j = 0; j < AWR_MAX_A; ++j) {
      if (j < A) {
        const float m = tanhf(__ldg(head + (long long)b * ldh + j));
        float ls, pass = 1.f;
        if (logstd_param) {
          ls = s_ls[j];
        } else {
          const float raw = __ldg(head + (long long)b * ldh + A + j);
          ls = fminf(fmaxf(raw, min_ls), max_ls);
          pass = (raw >= min_ls && raw <= max_ls) ? 1.f : 0.f;   // clamp backward
        }
        const float sd = expf(ls);
        const float inv_var = 1.f / (sd * sd);
        const float d = __ldg(actions + (long long)b * lda + j) - m;
        const float z = d * d * inv_var;
        logp += -0.5f * z - logf(sd) - half_log_2pi;
        d_head[(long long)b * lddh + j] = c * (d * inv_var) * (1.f - m * m);
        if (logstd_param) acc[j] += c * (z - 1.f);
        else d_head[(long long)b * lddh + A + j] = c * (z - 1.f) * pass;
      }
    }
    loss -= w * logp;
  }
  loss = block_sum(loss);
  if (threadIdx.x == 0) *metric_loss = loss * scale;
  if (logstd_param) {
#pragma unroll
    for (int j = 0; j < AWR_MAX_A; ++j) {
      if (j < A) {
        const float g = block_sum(acc[j]);
        if (threadIdx.x == 0) dlogstd_param[j] += g * s_dls[j];
      }
    }
    if (metric_mean_std && threadIdx.x == 0) {
      float s = 0.f;
      for (int j = 0; j < A; ++j) s += expf(s_ls[j]);
      *metric_mean_std = s / (float)A;
    }
  }
}

// AWACImpl.update_actor reports exp(logstd parameter).mean() AFTER the optimizer step (awac_impl.py:97-99)
__global__ void gauss_mean_std_kernel(const float* __restrict__ logstd_param, float min_ls, float max_ls, int A,
                                      float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < A; ++j) {
      const float sg = 1.f / (1.f + expf(-logstd_param[j]));
      s += expf(min_ls + sg * (max_ls - min_ls));
    }
    *out = s / (float)A;
  }
}

}  // namespace d3b

using namespace d3b;

extern "C" int d3b_gauss_mean_std(const float* logstd_param, float min_logstd, float max_logstd, int act_dim, float* out,
                                  void* stream) {
  D3B_REQUIRE(logstd_param && out && act_dim >= 1, "gauss_mean_std: bad arguments");
  launch_pdl(gauss_mean_std_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, logstd_param, min_logstd, max_logstd,
             act_dim, out);
  return check_launch("gauss_mean_std");
}

extern "C" int d3b_gauss_policy_rows(const float* head, int64_t ld_head, const float* logstd_param, float min_logstd,
                                     float max_logstd, const float* eps, const float* obs, int64_t ld_obs, float* x,
                                     int64_t ldx, int batch, int n, int obs_dim, int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 0 && n >= 1 && obs_dim >= 0 && act_dim >= 1, "gauss_policy_rows: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(head && eps && x && (obs || obs_dim == 0), "gauss_policy_rows: null pointer");
  const long long rows = (long long)batch * n;
  launch_pdl(gauss_policy_rows_kernel, dim3((unsigned)ceil_div_ll(rows * 32, 256)), dim3(256), 0, (cudaStream_t)stream,
             head, (long long)ld_head, logstd_param, min_logstd, max_logstd, eps, obs, (long long)ld_obs, x,
             (long long)ldx, batch, n, obs_dim, act_dim);
  return check_launch("gauss_policy_rows");
}

extern "C" int d3b_awr_weights(const float* q_data, int64_t stride_q_data, const float* q_samples,
                               int64_t stride_q_samples, int members, int batch, int n, int member_reduce,
                               int value_reduce, int weight_mode, float temperature, float max_weight, float* weights,
                               void* stream) {
  D3B_REQUIRE(batch >= 1 && n >= 1 && members >= 1, "awr_weights: bad sizes");
  D3B_REQUIRE(member_reduce >= 0 && member_reduce <= 1 && value_reduce >= 0 && value_reduce <= 1 && weight_mode >= 0 &&
                  weight_mode <= 2 && temperature > 0.f,
              "awr_weights: bad mode");
  D3B_REQUIRE(q_data && q_samples && weights, "awr_weights: null pointer");
  launch_pdl(awr_weights_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, q_data, (long long)stride_q_data, q_samples,
             (long long)stride_q_samples, members, batch, n, member_reduce, value_reduce, weight_mode, temperature,
             max_weight, weights);
  return check_launch("awr_weights");
}

extern "C" int d3b_gauss_wll_loss(const float* head, int64_t ld_head, const float* logstd_param, const float* actions,
                                  int64_t ld_act, const float* weights, float min_logstd, float max_logstd, float scale,
                                  float* d_head, int64_t ld_dhead, float* dlogstd_param, float* metric_loss,
                                  float* metric_mean_std, int batch, int act_dim, void* stream) {
  D3B_REQUIRE(batch >= 1 && act_dim >= 1 && act_dim <= AWR_MAX_A, "gauss_wll_loss: bad sizes (act_dim <= %d)", AWR_MAX_A);
  D3B_REQUIRE(head && actions && weights && d_head && metric_loss && (!logstd_param || dlogstd_param),
              "gauss_wll_loss: null pointer");
  launch_pdl(gauss_wll_loss_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, head, (long long)ld_head, logstd_param,
             actions, (long long)ld_act, weights, min_logstd, max_logstd, scale, d_head, (long long)ld_dhead,
             dlogstd_param, metric_loss, metric_mean_std, batch, act_dim);
  return check_launch("gauss_wll_loss");
}

// ---- PLAS glue (d3rlpy/algos/torch/plas_impl.py:138-168): the deterministic policy acts in the VAE's latent space,
// action = decode(s, 2 * tanh(fc(encoder(s)))).
namespace d3b {

// x[b] = [obs_b | scale * z_b]
__global__ void scaled_concat_rows_kernel(const float* __restrict__ obs, long long ldo, const float* __restrict__ z,
                                          long long ldz, float scale, float* __restrict__ x, long long ldx, int B, int O,
                                          int Z) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = O + Z;
  if (i >= (long long)B * W) return;
  const int b = (int)(i / W), j = (int)(i % W);
  x[(long long)b * ldx + j] = j < O ? __ldg(obs + (long long)b * ldo + j) : scale * __ldg(z + (long long)b * ldz + (j - O));
}

// out = scale * dy * (1 - y^2): gradient through y = tanh(pre) (and through the scale of the latent)
__global__ void tanh_backward_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ y,
                                     long long ldy, float scale, float* __restrict__ out, long long ldo, int rows,
                                     int cols) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const float v = __ldg(y + (long long)r * ldy + c);
  out[(long long)r * ldo + c] = scale * __ldg(dy + (long long)r * lddy + c) * (1.f - v * v);
}

}  // namespace d3b

extern "C" int d3b_scaled_concat_rows(const float* obs, int64_t ldo, const float* z, int64_t ldz, float scale, float* x,
                                      int64_t ldx, int batch, int obs_dim, int z_dim, void* stream) {
  D3B_REQUIRE(batch >= 0 && obs_dim >= 0 && z_dim >= 1, "scaled_concat_rows: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(z && x && (obs || obs_dim == 0), "scaled_concat_rows: null pointer");
  const long long n = (long long)batch * (obs_dim + z_dim);
  launch_pdl(scaled_concat_rows_kernel, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, (cudaStream_t)stream, obs,
             (long long)ldo, z, (long long)ldz, scale, x, (long long)ldx, batch, obs_dim, z_dim);
  return check_launch("scaled_concat_rows");
}

extern "C" int d3b_tanh_backward(const float* dy, int64_t lddy, const float* y, int64_t ldy, float scale, float* out,
                                 int64_t ldo, int rows, int cols, void* stream) {
  D3B_REQUIRE(rows >= 0 && cols >= 1, "tanh_backward: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && y && out, "tanh_backward: null pointer");
  const long long n = (long long)rows * cols;
  launch_pdl(tanh_backward_kernel, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, (cudaStream_t)stream, dy,
             (long long)lddy, y, (long long)ldy, scale, out, (long long)ldo, rows, cols);
  return check_launch("tanh_backward");
}
