// K8/K9 loss-side kernels: BCQ (conditional VAE, perturbation policy, mix-max target) and DQN /
// DoubleDQN / DiscreteCQL (Huber TD + discrete conservative term).  Reference semantics restated:
//   ConditionalVAE.encode/decode/compute_error      d3rlpy/models/torch/imitators.py:63-86
//   BCQImpl (latent clamp, residual action, losses)   d3rlpy/algos/torch/bcq_impl.py:115-226
//   DeterministicResidualPolicy.forward               d3rlpy/models/torch/policies.py:94-97
//   compute_max_with_n_actions_and_indices            d3rlpy/models/torch/q_functions/__init__.py:8-63
//   DQNImpl / DoubleDQNImpl                           d3rlpy/algos/torch/dqn_impl.py:97-171
//   DiscreteMeanQFunction.compute_error, Huber        q_functions/mean_q_function.py:26-42, utility.py:27-32
//   DiscreteCQLImpl.compute_loss                      d3rlpy/algos/torch/cql_impl.py:279-302
// Gradient formulas: SURVEY.md Appendix C.6-C.7.
#include "common.cuh"

namespace d3b {

__device__ __forceinline__ float clampf2(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// z = mu + exp(clamp(logstd)) * eps ; rows [obs | z] ; sums[0] += sum KL(N(mu,sd) || N(0,1)) over B*L
__global__ void __launch_bounds__(256) vae_sample_rows_kernel(const float* __restrict__ head, long long ldh,
                                                              const float* __restrict__ eps,
                                                              const float* __restrict__ obs, long long ldo,
                                                              float* __restrict__ X, long long ldx,
                                                              float* __restrict__ kl_sum, int B, int O, int Lz,
                                                              float min_logstd, float max_logstd) {
  int lane = threadIdx.x & 31;
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float kl = 0.f;
  if (b < B) {
    const float* o = obs + (long long)b * ldo;
    float* x = X + (long long)b * ldx;
    for (int j = lane; j < O; j += 32) x[j] = __ldg(o + j);
    for (int j = lane; j < Lz; j += 32) {
      float mu = __ldg(head + (long long)b * ldh + j);
      float ls = clampf2(__ldg(head + (long long)b * ldh + Lz + j), min_logstd, max_logstd);
      float sd = expf(ls);
      x[O + j] = mu + sd * __ldg(eps + (long long)b * Lz + j);
      float var = sd * sd;
      kl += 0.5f * (var + mu * mu - 1.f - logf(var));  // _kl_normal_normal with N(0,1)
    }
  }
  kl = block_sum(kl);
  if (threadIdx.x == 0) atomicAdd(kl_sum, kl);
}

// y = decoder output (tanh already applied).  sums[1] += sum (y - a)^2 ; dpre = 2 (y-a)/(B*A) * (1 - y^2)
__global__ void __launch_bounds__(256) vae_recon_kernel(const float* __restrict__ y, const float* __restrict__ act,
                                                        long long lda, float* __restrict__ dpre,
                                                        float* __restrict__ sq_sum, int B, int A, float inv_ba) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  if (idx < B * A) {
    int b = idx / A, j = idx % A;
    float yv = __ldg(y + idx);
    float d = yv - __ldg(act + (long long)b * lda + j);
    s = d * d;
    dpre[idx] = 2.f * inv_ba * d * (1.f - yv * yv);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(sq_sum, s);
}

// d(head) of the VAE encoder from dz (gradient w.r.t. the latent columns of the decoder input) + KL term.
__global__ void __launch_bounds__(256) vae_backward_kernel(const float* __restrict__ head, long long ldh,
                                                           const float* __restrict__ eps,
                                                           const float* __restrict__ dz, long long lddz,
                                                           float* __restrict__ dhead, long long lddh, int B, int Lz,
                                                           float min_logstd, float max_logstd, float beta_inv_bl) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Lz) return;
  int b = idx / Lz, j = idx % Lz;
  float mu = __ldg(head + (long long)b * ldh + j);
  float raw = __ldg(head + (long long)b * ldh + Lz + j);
  float ls = clampf2(raw, min_logstd, max_logstd);
  float sd = expf(ls);
  float g = __ldg(dz + (long long)b * lddz + j);
  float inside = (raw >= min_logstd && raw <= max_logstd) ? 1.f : 0.f;
  dhead[(long long)b * lddh + j] = g + beta_inv_bl * mu;
  dhead[(long long)b * lddh + Lz + j] = inside * (g * sd * __ldg(eps + (long long)b * Lz + j) +
                                                  beta_inv_bl * (sd * sd - 1.f));
}

// metric = sums[1]/(B*A) + beta * sums[0]/(B*L)
__global__ void vae_finalize_kernel(const float* sums, float inv_ba, float beta_inv_bl, float* metric) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *metric = sums[1] * inv_ba + beta_inv_bl * sums[0];
}

// a = clamp(sampled + scale * tanh(z), -1, 1) ; rows [obs[b] | a]   (row r belongs to batch row r / n_repeat)
__global__ void __launch_bounds__(256) residual_rows_kernel(const float* __restrict__ z, long long ldz,
                                                            const float* __restrict__ sampled, long long lds,
                                                            const float* __restrict__ obs, long long ldo,
                                                            float* __restrict__ X, long long ldx, float scale,
                                                            int rows, int n_repeat, int O, int A) {
  int lane = threadIdx.x & 31;
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* o = obs + (r / n_repeat) * ldo;
  float* x = X + r * ldx;
  for (int j = lane; j < O; j += 32) x[j] = __ldg(o + j);
  for (int j = lane; j < A; j += 32)
    x[O + j] = clampf2(__ldg(sampled + r * lds + j) + scale * tanhf(__ldg(z + r * ldz + j)), -1.f, 1.f);
}

// dz = da * [pre in [-1,1]] * scale * (1 - tanh(z)^2)
__global__ void __launch_bounds__(256) residual_backward_kernel(const float* __restrict__ z, long long ldz,
                                                                const float* __restrict__ sampled, long long lds,
                                                                const float* __restrict__ da, long long ldda,
                                                                float* __restrict__ dz, long long lddz, float scale,
                                                                int B, int A) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * A) return;
  int b = idx / A, j = idx % A;
  float t = tanhf(__ldg(z + (long long)b * ldz + j));
  float pre = __ldg(sampled + (long long)b * lds + j) + scale * t;
  float gate = (pre >= -1.f && pre <= 1.f) ? 1.f : 0.f;
  dz[(long long)b * lddz + j] = __ldg(da + (long long)b * ldda + j) * gate * scale * (1.f - t * t);
}

// q: [E][B*N]; out[b] = max_k ( (1-lam) max_e q + lam min_e q )   (one warp per batch row)
__global__ void __launch_bounds__(256) bcq_target_reduce_kernel(const float* __restrict__ q, long long sQ,
                                                                float* __restrict__ out, int B, int N, int E,
                                                                float lam) {
  int lane = threadIdx.x & 31;
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float best = -INFINITY;
  for (int k = lane; k < N; k += 32) {
    float mx = -INFINITY, mn = INFINITY;
    for (int e = 0; e < E; ++e) {
      float v = __ldg(q + (long long)e * sQ + (long long)b * N + k);
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    best = fmaxf(best, (1.f - lam) * mx + lam * mn);
  }
  best = warp_max(best);
  if (lane == 0) out[b] = best;
}

// loss = -mean(q0) ; dq[b] = -1/B
__global__ void __launch_bounds__(256) neg_mean_seed_kernel(const float* __restrict__ q0, float* __restrict__ dq,
                                                            float* __restrict__ loss_sum, int B, float inv_b) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  if (b < B) {
    s = -__ldg(q0 + b) * inv_b;
    dq[b] = -inv_b;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(loss_sum, s);
}

// ---------------------------------------------------------------------------------------- discrete Q
// DoubleDQN target: a* = argmax_a mean_e Q_online(s')[e][b][a]; q_tpn[b] = min_e Q_targ(s')[e][b][a*]
__global__ void __launch_bounds__(256) dqn_target_kernel(const float* __restrict__ q_online, long long sQo,
                                                         const float* __restrict__ q_targ, long long sQt,
                                                         float* __restrict__ q_tpn, int B, int A, int E) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int best = 0;
  float bv = -INFINITY;
  for (int a = 0; a < A; ++a) {
    float m = 0.f;
    for (int e = 0; e < E; ++e) m += __ldg(q_online + (long long)e * sQo + (long long)b * A + a);
    m /= (float)E;
    if (m > bv) { bv = m; best = a; }
  }
  float mn = INFINITY;
  for (int e = 0; e < E; ++e) mn = fminf(mn, __ldg(q_targ + (long long)e * sQt + (long long)b * A + best));
  q_tpn[b] = mn;
}

// Huber TD (sum over members of batch means) + alpha * mean_b(logsumexp_a Qbar - Qbar[a_data]) and dQ.
// sums[0] += sum_e sum_b huber ; sums[1] += sum_b (lse - data)
__global__ void __launch_bounds__(256) dcql_loss_kernel(const float* __restrict__ q, long long sQ,
                                                        const float* __restrict__ q_tpn,
                                                        const float* __restrict__ actions,
                                                        const float* __restrict__ rew, const float* __restrict__ term,
                                                        const float* __restrict__ nsteps, float gamma, float alpha,
                                                        float* __restrict__ dq, long long sDq,
                                                        float* __restrict__ sums, int B, int A, int E, float inv_b,
                                                        int conservative) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  float hub = 0.f, cons = 0.f;
  if (b < B) {
    int ad = (int)__ldg(actions + b);
    float n = __ldg(nsteps + b);
    float g = n == 1.f ? gamma : powf(gamma, n);
    float y = __ldg(rew + b) + g * __ldg(q_tpn + b) * (1.f - __ldg(term + b));
    // mean over members, logsumexp over actions
    float mx = -INFINITY;
    for (int a = 0; a < A; ++a) {
      float m = 0.f;
      for (int e = 0; e < E; ++e) m += __ldg(q + (long long)e * sQ + (long long)b * A + a);
      mx = fmaxf(mx, m / (float)E);
    }
    float se = 0.f, data = 0.f;
    for (int a = 0; a < A; ++a) {
      float m = 0.f;
      for (int e = 0; e < E; ++e) m += __ldg(q + (long long)e * sQ + (long long)b * A + a);
      m /= (float)E;
      se += expf(m - mx);
      if (a == ad) data = m;
    }
    cons = mx + logf(se) - data;
    float cscale = conservative ? alpha * inv_b / (float)E : 0.f;
    for (int e = 0; e < E; ++e) {
      const float* qe = q + (long long)e * sQ + (long long)b * A;
      float* de = dq + (long long)e * sDq + (long long)b * A;
      float diff = y - __ldg(qe + ad);
      hub += fabsf(diff) < 1.f ? 0.5f * diff * diff : (fabsf(diff) - 0.5f);
      float dtd = -clampf2(diff, -1.f, 1.f) * inv_b;
      for (int a = 0; a < A; ++a) {
        float m = 0.f;
        for (int e2 = 0; e2 < E; ++e2) m += __ldg(q + (long long)e2 * sQ + (long long)b * A + a);
        m /= (float)E;
        float gcons = cscale * (expf(m - mx) / se - (a == ad ? 1.f : 0.f));
        de[a] = gcons + (a == ad ? dtd : 0.f);
      }
    }
  }
  hub = block_sum(hub);
  if (threadIdx.x == 0) atomicAdd(sums + 0, hub);
  cons = block_sum(cons);
  if (threadIdx.x == 0) atomicAdd(sums + 1, cons);
}

__global__ void dcql_finalize_kernel(const float* sums, float inv_b, float alpha, int conservative, float* metric) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *metric = sums[0] * inv_b + (conservative ? alpha * sums[1] * inv_b : 0.f);
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

extern "C" int d3b_vae_sample_rows(const float* head, int64_t ld_head, const float* eps, const float* obs,
                                   int64_t ldo, float* x, int64_t ldx, float* kl_sum, int batch, int obs_dim,
                                   int latent, float min_logstd, float max_logstd, void* stream) {
  D3B_REQUIRE(batch >= 0 && latent >= 1 && obs_dim >= 0, "vae_sample_rows: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(head && eps && obs && x && kl_sum, "vae_sample_rows: null pointer");
  vae_sample_rows_kernel<<<ceil_div(batch, 8), 256, 0, ST>>>(head, ld_head, eps, obs, ldo, x, ldx, kl_sum, batch,
                                                             obs_dim, latent, min_logstd, max_logstd);
  return check_launch("vae_sample_rows");
}

extern "C" int d3b_vae_recon(const float* y, const float* actions, int64_t lda, float* dpre, float* sq_sum, int batch,
                             int act_dim, float inv_batch, void* stream) {
  D3B_REQUIRE(batch >= 0 && act_dim >= 1, "vae_recon: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(y && actions && dpre && sq_sum, "vae_recon: null pointer");
  vae_recon_kernel<<<ceil_div(batch * act_dim, 256), 256, 0, ST>>>(y, actions, lda, dpre, sq_sum, batch, act_dim,
                                                                   inv_batch / (float)act_dim);
  return check_launch("vae_recon");
}

extern "C" int d3b_vae_backward(const float* head, int64_t ld_head, const float* eps, const float* dz, int64_t lddz,
                                float* dhead, int64_t ld_dhead, int batch, int latent, float min_logstd,
                                float max_logstd, float beta, float inv_batch, void* stream) {
  D3B_REQUIRE(batch >= 0 && latent >= 1, "vae_backward: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(head && eps && dz && dhead, "vae_backward: null pointer");
  vae_backward_kernel<<<ceil_div(batch * latent, 256), 256, 0, ST>>>(head, ld_head, eps, dz, lddz, dhead, ld_dhead,
                                                                     batch, latent, min_logstd, max_logstd,
                                                                     beta * inv_batch / (float)latent);
  return check_launch("vae_backward");
}

extern "C" int d3b_vae_finalize(const float* sums, int act_dim, int latent, float beta, float inv_batch,
                                float* metric, void* stream) {
  D3B_REQUIRE(sums && metric && act_dim >= 1 && latent >= 1, "vae_finalize: bad arguments");
  vae_finalize_kernel<<<1, 32, 0, ST>>>(sums, inv_batch / (float)act_dim, beta * inv_batch / (float)latent, metric);
  return check_launch("vae_finalize");
}

extern "C" int d3b_residual_rows(const float* z, int64_t ldz, const float* sampled, int64_t lds, const float* obs,
                                 int64_t ldo, float* x, int64_t ldx, float scale, int rows, int n_repeat,
                                 int obs_dim, int act_dim, void* stream) {
  D3B_REQUIRE(rows >= 0 && n_repeat >= 1 && act_dim >= 1, "residual_rows: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(z && sampled && obs && x, "residual_rows: null pointer");
  residual_rows_kernel<<<ceil_div(rows, 8), 256, 0, ST>>>(z, ldz, sampled, lds, obs, ldo, x, ldx, scale, rows,
                                                          n_repeat, obs_dim, act_dim);
  return check_launch("residual_rows");
}

extern "C" int d3b_residual_backward(const float* z, int64_t ldz, const float* sampled, int64_t lds, const float* da,
                                     int64_t ldda, float* dz, int64_t lddz, float scale, int batch, int act_dim,
                                     void* stream) {
  D3B_REQUIRE(batch >= 0 && act_dim >= 1, "residual_backward: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(z && sampled && da && dz, "residual_backward: null pointer");
  residual_backward_kernel<<<ceil_div(batch * act_dim, 256), 256, 0, ST>>>(z, ldz, sampled, lds, da, ldda, dz, lddz,
                                                                           scale, batch, act_dim);
  return check_launch("residual_backward");
}

extern "C" int d3b_bcq_target_reduce(const float* q, int64_t stride_q, float* q_tpn, int batch, int n_actions,
                                     int members, float lam, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_actions >= 1 && members >= 1, "bcq_target_reduce: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q && q_tpn, "bcq_target_reduce: null pointer");
  bcq_target_reduce_kernel<<<ceil_div(batch, 8), 256, 0, ST>>>(q, stride_q, q_tpn, batch, n_actions, members, lam);
  return check_launch("bcq_target_reduce");
}

extern "C" int d3b_neg_mean_seed(const float* q0, float* dq, float* loss_sum, int batch, float inv_batch,
                                 void* stream) {
  D3B_REQUIRE(batch >= 0, "neg_mean_seed: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q0 && dq && loss_sum, "neg_mean_seed: null pointer");
  neg_mean_seed_kernel<<<ceil_div(batch, 256), 256, 0, ST>>>(q0, dq, loss_sum, batch, inv_batch);
  return check_launch("neg_mean_seed");
}

extern "C" int d3b_dqn_target(const float* q_online, int64_t stride_qo, const float* q_targ, int64_t stride_qt,
                              float* q_tpn, int batch, int n_actions, int members, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_actions >= 1 && members >= 1, "dqn_target: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q_online && q_targ && q_tpn, "dqn_target: null pointer");
  dqn_target_kernel<<<ceil_div(batch, 256), 256, 0, ST>>>(q_online, stride_qo, q_targ, stride_qt, q_tpn, batch,
                                                          n_actions, members);
  return check_launch("dqn_target");
}

extern "C" int d3b_dcql_loss(const float* q, int64_t stride_q, const float* q_tpn, const float* actions,
                             const float* rewards, const float* terminals, const float* n_steps, float gamma,
                             float alpha, float* dq, int64_t stride_dq, float* sums, int batch, int n_actions,
                             int members, float inv_batch, int conservative, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_actions >= 1 && members >= 1, "dcql_loss: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(q && q_tpn && actions && rewards && terminals && n_steps && dq && sums, "dcql_loss: null pointer");
  dcql_loss_kernel<<<ceil_div(batch, 256), 256, 0, ST>>>(q, stride_q, q_tpn, actions, rewards, terminals, n_steps,
                                                         gamma, alpha, dq, stride_dq, sums, batch, n_actions, members,
                                                         inv_batch, conservative);
  return check_launch("dcql_loss");
}

extern "C" int d3b_dcql_finalize(const float* sums, float inv_batch, float alpha, int conservative, float* metric,
                                 void* stream) {
  D3B_REQUIRE(sums && metric, "dcql_finalize: null pointer");
  dcql_finalize_kernel<<<1, 32, 0, ST>>>(sums, inv_batch, alpha, conservative, metric);
  return check_launch("dcql_finalize");
}
