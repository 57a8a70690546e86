// Quantile-regression Q head for the discrete algorithms (DQN / DoubleDQN / DiscreteCQL with
// QRQFunctionFactory): target selection and the quantile-Huber loss with its hand-written gradient.
// The head itself (feature -> A * n_quantiles) is a dense layer and runs on the GEMM kernels.  Replaces
//   d3rlpy/models/torch/q_functions/qr_q_function.py:15-88      (taus, forward = mean over quantiles, compute_error,
//                                                                compute_target)
//   d3rlpy/models/torch/q_functions/utility.py:17-61            (pick_quantile_value_by_action, quantile Huber)
//   d3rlpy/models/torch/q_functions/ensemble_q_function.py:27-66,108-134  (_reduce_quantile_ensemble "min")
//   d3rlpy/algos/torch/dqn_impl.py:113-141,162-171, cql_impl.py:279-302   (the callers)
#include "common.cuh"

namespace d3b {

// theta layout: [E][B][A][n] (row b of member e = the head's A*n outputs, action-major like `.view(-1, A, n)`).
//
// One warp per sample: a* = argmax_a mean_e(mean_i theta_sel[e][b][a][i]);  e* = argmin_e mean_i theta_targ[e][b][a*][i]
// (first index on ties, like torch.min);  q_tpn[b][:] = theta_targ[e*][b][a*][:].
__global__ void __launch_bounds__(256) qr_target_kernel(const float* __restrict__ th_sel, long long sSel,
                                                        const float* __restrict__ th_targ, long long sTarg,
                                                        float* __restrict__ q_tpn, int B, int A, int n, int E) {
  int lane = threadIdx.x & 31;
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  long long row = (long long)b * A * n;
  int best = 0;
  float bv = -INFINITY;
  for (int a = 0; a < A; ++a) {
    float m = 0.f;
    for (int e = 0; e < E; ++e) {
      const float* t = th_sel + (long long)e * sSel + row + (long long)a * n;
      float s = 0.f;
      for (int i = lane; i < n; i += 32) s += __ldg(t + i);
      m += warp_sum(s) / (float)n;
    }
    m /= (float)E;
    if (m > bv) { bv = m; best = a; }
  }
  int emin = 0;
  float mv = INFINITY;
  for (int e = 0; e < E; ++e) {
    const float* t = th_targ + (long long)e * sTarg + row + (long long)best * n;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += __ldg(t + i);
    s = warp_sum(s) / (float)n;
    if (s < mv) { mv = s; emin = e; }
  }
  const float* t = th_targ + (long long)emin * sTarg + row + (long long)best * n;
  for (int i = lane; i < n; i += 32) q_tpn[(long long)b * n + i] = __ldg(t + i);
}

// One block per sample.  Quantile Huber (sum over members of batch means):
//   L[e][b] = (1/n) sum_j sum_i |tau_i - 1[y_j - th_i < 0]| * huber(y_j - th_i),   th = theta[e][b][a_data][:],
//   y_j = r + gamma^n_steps * q_tpn[b][j] * (1 - terminal),   tau_i = ((i+1)/n + i/n) / 2
// plus (conservative) alpha * (logsumexp_a Vbar[a] - Vbar[a_data]), Vbar[a] = mean_e mean_i theta[e][b][a][i].
// Writes d(loss)/d(theta) for every (e, a, i) of the row and the sample's two loss terms to partials[b], partials[B+b];
// qr_loss_reduce_kernel adds them up in a fixed order (bit-reproducible metric, unlike one atomicAdd per block).
__global__ void __launch_bounds__(128) qr_loss_kernel(const float* __restrict__ theta, long long sTh,
                                                      const float* __restrict__ q_tpn,
                                                      const float* __restrict__ actions,
                                                      const float* __restrict__ rew, const float* __restrict__ term,
                                                      const float* __restrict__ nsteps, float gamma, float alpha,
                                                      float* __restrict__ dtheta, long long sD,
                                                      float* __restrict__ partials, int B, int A, int n, int E,
                                                      float inv_b, int conservative) {
  extern __shared__ float sm[];
  float* y = sm;           // [n]
  float* v = sm + n;       // [E*A] member values, then [A] ensemble means
  float* vbar = v + E * A;
  int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  int ad = (int)__ldg(actions + b);
  float ns = __ldg(nsteps + b);
  float g = ns == 1.f ? gamma : powf(gamma, ns);
  float r = __ldg(rew + b), nt = 1.f - __ldg(term + b);
  for (int j = tid; j < n; j += blockDim.x) y[j] = r + g * __ldg(q_tpn + (long long)b * n + j) * nt;
  long long row = (long long)b * A * n;
  float cons = 0.f, mx = 0.f, se = 1.f;
  if (conservative) {
    for (int p = warp; p < E * A; p += nwarp) {
      const float* t = theta + (long long)(p / A) * sTh + row + (long long)(p % A) * n;
      float s = 0.f;
      for (int i = lane; i < n; i += 32) s += __ldg(t + i);
      s = warp_sum(s);
      if (lane == 0) v[p] = s / (float)n;
    }
    __syncthreads();
    if (tid < A) {
      float m = 0.f;
      for (int e = 0; e < E; ++e) m += v[e * A + tid];
      vbar[tid] = m / (float)E;
    }
    __syncthreads();
    mx = -INFINITY;
    for (int a = 0; a < A; ++a) mx = fmaxf(mx, vbar[a]);
    se = 0.f;
    for (int a = 0; a < A; ++a) se += expf(vbar[a] - mx);
    cons = mx + logf(se) - vbar[ad];
  } else {
    __syncthreads();
  }
  float cscale = conservative ? alpha * inv_b / ((float)E * (float)n) : 0.f;
  float inv_n = 1.f / (float)n;
  float loss = 0.f;
  for (int e = 0; e < E; ++e) {
    const float* th = theta + (long long)e * sTh + row;
    float* d = dtheta + (long long)e * sD + row;
    // conservative gradient (every action), then the TD gradient on the taken action's quantiles
    for (int p = tid; p < A * n; p += blockDim.x) {
      int a = p / n;
      float gc = conservative ? cscale * (expf(vbar[a] - mx) / se - (a == ad ? 1.f : 0.f)) : 0.f;
      if (a != ad) d[p] = gc;
    }
    for (int i = tid; i < n; i += blockDim.x) {
      float ti = __ldg(th + (long long)ad * n + i);
      float tau = ((float)(i + 1) / (float)n + (float)i / (float)n) / 2.f;
      float al = 0.f, ag = 0.f;
      for (int j = 0; j < n; ++j) {
        float diff = y[j] - ti;
        float ab = fabsf(diff);
        float h = ab < 1.f ? 0.5f * diff * diff : ab - 0.5f;
        float w = fabsf(tau - (diff < 0.f ? 1.f : 0.f));
        al = fmaf(w, h, al);
        ag = fmaf(w, -fminf(fmaxf(diff, -1.f), 1.f), ag);
      }
      loss += al * inv_n;
      float gc = conservative ? cscale * (expf(vbar[ad] - mx) / se - 1.f) : 0.f;
      d[(long long)ad * n + i] = gc + ag * inv_n * inv_b;
    }
  }
  loss = block_sum(loss);
  if (tid == 0) {
    partials[b] = loss;
    partials[B + b] = cons;
  }
}

// sums[0] += sum_b partials[b], sums[1] += sum_b partials[B + b]: one block, fixed summation tree.
__global__ void __launch_bounds__(256) qr_loss_reduce_kernel(const float* __restrict__ partials, float* __restrict__ sums,
                                                             int B) {
  float a = 0.f, c = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    a += partials[b];
    c += partials[B + b];
  }
  a = block_sum(a);
  c = block_sum(c);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 0, a);
    atomicAdd(sums + 1, c);
  }
}

// values[e][b][a] = mean_i theta[e][b][a][i]  (DiscreteQRQFunction.forward): one warp per (e, b, a).
__global__ void __launch_bounds__(256) qr_values_kernel(const float* __restrict__ theta, long long sTh,
                                                        float* __restrict__ values, long long sV, int B, int A, int n,
                                                        int E) {
  int lane = threadIdx.x & 31;
  long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)E * B * A) return;
  int e = (int)(w / ((long long)B * A));
  long long ba = w % ((long long)B * A);
  const float* t = theta + (long long)e * sTh + ba * n;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += __ldg(t + i);
  s = warp_sum(s);
  if (lane == 0) values[(long long)e * sV + ba] = s / (float)n;
}

// d(theta)[v][i] = d(values)[v] / n: backward of `quantiles.mean(dim=-1)` (qr_q_function.py:44-48,118-122).
__global__ void __launch_bounds__(256) qr_values_backward_kernel(const float* __restrict__ dvalues,
                                                                 float* __restrict__ dtheta, long long total, int n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) dtheta[i] = __ldg(dvalues + i / n) / (float)n;
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

extern "C" int d3b_qr_target(const float* theta_select, int64_t stride_select, const float* theta_targ,
                             int64_t stride_targ, float* q_tpn, int batch, int n_actions, int n_quantiles,
                             int members, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_actions >= 1 && n_quantiles >= 1 && members >= 1, "qr_target: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(theta_select && theta_targ && q_tpn, "qr_target: null pointer");
  qr_target_kernel<<<ceil_div(batch, 8), 256, 0, ST>>>(theta_select, stride_select, theta_targ, stride_targ, q_tpn,
                                                       batch, n_actions, n_quantiles, members);
  return check_launch("qr_target");
}

extern "C" int d3b_qr_loss(const float* theta, int64_t stride_theta, const float* q_tpn, const float* actions,
                           const float* rewards, const float* terminals, const float* n_steps, float gamma,
                           float alpha, float* dtheta, int64_t stride_dtheta, float* partials, float* sums, int batch,
                           int n_actions, int n_quantiles, int members, float inv_batch, int conservative,
                           void* stream) {
  D3B_REQUIRE(batch >= 0 && n_actions >= 1 && n_quantiles >= 1 && members >= 1, "qr_loss: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(theta && q_tpn && actions && rewards && terminals && n_steps && dtheta && partials && sums,
              "qr_loss: null pointer");
  size_t smem = ((size_t)n_quantiles + (size_t)members * n_actions + n_actions) * sizeof(float);
  D3B_REQUIRE(smem <= 48 * 1024, "qr_loss: n_quantiles + members * n_actions too large (%zu bytes of shared memory)",
              smem);
  qr_loss_kernel<<<batch, 128, smem, ST>>>(theta, stride_theta, q_tpn, actions, rewards, terminals, n_steps, gamma,
                                           alpha, dtheta, stride_dtheta, partials, batch, n_actions, n_quantiles,
                                           members, inv_batch, conservative);
  int rc = check_launch("qr_loss");
  if (rc) return rc;
  qr_loss_reduce_kernel<<<1, 256, 0, ST>>>(partials, sums, batch);
  return check_launch("qr_loss_reduce");
}

extern "C" int d3b_qr_values(const float* theta, int64_t stride_theta, float* values, int64_t stride_values, int batch,
                             int n_actions, int n_quantiles, int members, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_actions >= 1 && n_quantiles >= 1 && members >= 1, "qr_values: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(theta && values, "qr_values: null pointer");
  long long warps = (long long)members * batch * n_actions;
  qr_values_kernel<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, ST>>>(theta, stride_theta, values, stride_values, batch,
                                                                    n_actions, n_quantiles, members);
  return check_launch("qr_values");
}

extern "C" int d3b_qr_values_backward(const float* dvalues, float* dtheta, int64_t n_values, int n_quantiles,
                                      void* stream) {
  D3B_REQUIRE(n_values >= 0 && n_quantiles >= 1, "qr_values_backward: bad sizes");
  if (n_values == 0) return D3B_OK;
  D3B_REQUIRE(dvalues && dtheta, "qr_values_backward: null pointer");
  long long total = (long long)n_values * n_quantiles;
  qr_values_backward_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, ST>>>(dvalues, dtheta, total, n_quantiles);
  return check_launch("qr_values_backward");
}
