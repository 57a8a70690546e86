// Narrow output heads (out_features <= 32): Q head (256->1), policy mu|logstd (256->2A), VAE heads,
// discrete Q head (512->A).  These are GEMV-shaped, so they run as warp-shuffle reductions instead of
// wasting GEMM tiles.  Replaces the `_fc` / `_mu` / `_logstd` nn.Linear calls of
//   d3rlpy/models/torch/q_functions/mean_q_function.py:21,24,69,72
//   d3rlpy/models/torch/policies.py:55-59,92-97,153-181
//   d3rlpy/models/torch/imitators.py:45-54,63-72
#include "common.cuh"

namespace d3b {

constexpr int HEAD_MAX_N = 32;

// One warp per (member,row).  Y[e][m][n] = act(X[e][m][:] . W[e][n][:] + b[e][n]).  The row stays in registers
// (K <= 1024) and four outputs are reduced at a time, so the shuffle chains of different outputs overlap.
__global__ void __launch_bounds__(256) head_forward_kernel(const float* __restrict__ X, long long ldx, long long sX,
                                                           const float* __restrict__ W, long long ldw, long long sW,
                                                           const float* __restrict__ bias, long long sB,
                                                           float* __restrict__ Y, long long ldy, long long sY, int M,
                                                           int N, int K, int E, int act_tanh) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31;
  long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)M * E) return;
  int e = (int)(wid / M), m = (int)(wid % M);
  const float* x = X + (long long)e * sX + (long long)m * ldx;
  const float* w = W + (long long)e * sW;
  float mine = 0.f;
  if (K <= 1024) {
    float xr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) xr[i] = (lane + 32 * i < K) ? __ldg(x + lane + 32 * i) : 0.f;
    for (int n0 = 0; n0 < N; n0 += 4) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + j < N) {
          const float* wr = w + (long long)(n0 + j) * ldw;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (32 * i < K) { const int k = lane + 32 * i; if (k < K) s[j] = fmaf(xr[i], __ldg(wr + k), s[j]); }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane == n0 + j) mine = s[j];
    }
  } else {
    for (int n = 0; n < N; ++n) {
      const float* wr = w + (long long)n * ldw;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s = fmaf(__ldg(x + k), __ldg(wr + k), s);
      s = warp_sum(s);
      if (lane == n) mine = s;
    }
  }
  if (lane < N) {
    float v = mine + (bias ? __ldg(bias + (long long)e * sB + lane) : 0.f);
    if (act_tanh) v = tanhf(v);
    Y[(long long)e * sY + (long long)m * ldy + lane] = v;
  }
}

// dX[e][m][k] = (sum_n dY[e][m][n] W[e][n][k]) * [src[e][m][k] > 0]; VEC = 4: one thread per four consecutive k
// (16-byte loads / stores; K, the leading dimensions and the member strides multiples of 4, pointers 16-byte aligned)
template <int VEC>
__global__ void __launch_bounds__(256) head_backward_data_kernel(
    const float* __restrict__ dY, long long lddy, long long sdY, const float* __restrict__ W, long long ldw,
    long long sW, float* __restrict__ dX, long long lddx, long long sdX, const float* __restrict__ src,
    long long ldsrc, long long sSrc, int M, int N, int K, int E) {
  pdl_trigger();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int KV = K / VEC;
  long long total = (long long)E * M * KV;
  if (idx >= total) return;
  int k = (int)(idx % KV) * VEC;
  long long t = idx / KV;
  int m = (int)(t % M), e = (int)(t / M);
  const float* dy = dY + (long long)e * sdY + (long long)m * lddy;
  const float* w = W + (long long)e * sW + k;
  if (VEC == 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = 0; n < N; ++n) {
      const float d = __ldg(dy + n);
      const float4 wv = __ldg((const float4*)(w + (long long)n * ldw));
      s.x = fmaf(d, wv.x, s.x); s.y = fmaf(d, wv.y, s.y); s.z = fmaf(d, wv.z, s.z); s.w = fmaf(d, wv.w, s.w);
    }
    if (src) {
      const float4 h = __ldg((const float4*)(src + (long long)e * sSrc + (long long)m * ldsrc + k));
      if (!(h.x > 0.f)) s.x = 0.f;
      if (!(h.y > 0.f)) s.y = 0.f;
      if (!(h.z > 0.f)) s.z = 0.f;
      if (!(h.w > 0.f)) s.w = 0.f;
    }
    *reinterpret_cast<float4*>(dX + (long long)e * sdX + (long long)m * lddx + k) = s;
  } else {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(__ldg(dy + n), __ldg(w + (long long)n * ldw), s);
    if (src && !(__ldg(src + (long long)e * sSrc + (long long)m * ldsrc + k) > 0.f)) s = 0.f;
    dX[(long long)e * sdX + (long long)m * lddx + k] = s;
  }
}

// dW[e][n][k] += sum_m dY[e][m][n] X[e][m][k];  db[e][n] += sum_m dY[e][m][n]
// grid: (ceil(K/256), row chunks, E); dY chunk staged in shared memory.
template <int NMAX>
__global__ void __launch_bounds__(256) head_backward_weight_kernel(
    const float* __restrict__ dY, long long lddy, long long sdY, const float* __restrict__ X, long long ldx,
    long long sX, float* __restrict__ dW, long long lddw, long long sdW, float* __restrict__ db, long long sdb, int M,
    int N, int K, int rows_per_block) {
  extern __shared__ float sdy[];  // [rows_per_block][N]
  pdl_trigger();
  pdl_wait();
  int e = blockIdx.z;
  int m0 = blockIdx.y * rows_per_block;
  int rows = min(rows_per_block, M - m0);
  const float* dy = dY + (long long)e * sdY + (long long)m0 * lddy;
  for (int i = threadIdx.x; i < rows * N; i += blockDim.x) {
    int r = i / N, n = i % N;
    sdy[i] = __ldg(dy + (long long)r * lddy + n);
  }
  __syncthreads();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) {
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    const float* x = X + (long long)e * sX + (long long)m0 * ldx + k;
    for (int r = 0; r < rows; ++r) {
      float xv = __ldg(x + (long long)r * ldx);
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) acc[n] = fmaf(sdy[r * N + n], xv, acc[n]);
    }
    float* dw = dW + (long long)e * sdW + k;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) atomicAdd(dw + (long long)n * lddw, acc[n]);
  }
  if (db && blockIdx.x == 0 && threadIdx.x < N) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += sdy[r * N + threadIdx.x];
    atomicAdd(db + (long long)e * sdb + threadIdx.x, s);
  }
}

}  // namespace d3b

using namespace d3b;

extern "C" int d3b_head_forward(const float* x, int64_t ldx, int64_t stride_x, const float* w, int64_t ldw,
                                int64_t stride_w, const float* bias, int64_t stride_b, float* y, int64_t ldy,
                                int64_t stride_y, int rows, int out_features, int in_features, int members,
                                int act_tanh, void* stream) {
  D3B_REQUIRE(rows >= 0 && in_features > 0 && members > 0, "head_forward: bad sizes");
  D3B_REQUIRE(out_features >= 1 && out_features <= HEAD_MAX_N, "head_forward: out_features %d not in 1..%d",
              out_features, HEAD_MAX_N);
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(x && w && y, "head_forward: null pointer");
  long long warps = (long long)rows * members;
  launch_pdl(head_forward_kernel, dim3((unsigned)ceil_div_ll(warps, 8)), dim3(256), 0, (cudaStream_t)stream, x,
             (long long)ldx, (long long)stride_x, w, (long long)ldw, (long long)stride_w, bias, (long long)stride_b, y,
             (long long)ldy, (long long)stride_y, rows, out_features, in_features, members, act_tanh);
  return check_launch("head_forward");
}

extern "C" int d3b_head_backward_data(const float* dy, int64_t lddy, int64_t stride_dy, const float* w, int64_t ldw,
                                      int64_t stride_w, float* dx, int64_t lddx, int64_t stride_dx,
                                      const float* relu_src, int64_t ld_src, int64_t stride_src, int rows,
                                      int out_features, int in_features, int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && in_features > 0 && members > 0 && out_features >= 1, "head_backward_data: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && w && dx, "head_backward_data: null pointer");
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = in_features % 4 == 0 && ldw % 4 == 0 && stride_w % 4 == 0 && lddx % 4 == 0 && stride_dx % 4 == 0 &&
                   al16(w) && al16(dx) && (!relu_src || (al16(relu_src) && ld_src % 4 == 0 && stride_src % 4 == 0));
  long long total = (long long)rows * (in_features / (vec ? 4 : 1)) * members;
  if (vec)
    launch_pdl(head_backward_data_kernel<4>, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, (cudaStream_t)stream,
               dy, (long long)lddy, (long long)stride_dy, w, (long long)ldw, (long long)stride_w, dx, (long long)lddx,
               (long long)stride_dx, relu_src, (long long)ld_src, (long long)stride_src, rows, out_features,
               in_features, members);
  else
    launch_pdl(head_backward_data_kernel<1>, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, (cudaStream_t)stream,
               dy, (long long)lddy, (long long)stride_dy, w, (long long)ldw, (long long)stride_w, dx, (long long)lddx,
               (long long)stride_dx, relu_src, (long long)ld_src, (long long)stride_src, rows, out_features,
               in_features, members);
  return check_launch("head_backward_data");
}

extern "C" int d3b_head_backward_weight(const float* dy, int64_t lddy, int64_t stride_dy, const float* x,
                                        int64_t ldx, int64_t stride_x, float* dw, int64_t lddw, int64_t stride_dw,
                                        float* dbias, int64_t stride_db, int rows, int out_features, int in_features,
                                        int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && in_features > 0 && members > 0, "head_backward_weight: bad sizes");
  D3B_REQUIRE(out_features >= 1 && out_features <= HEAD_MAX_N, "head_backward_weight: out_features %d not in 1..%d",
              out_features, HEAD_MAX_N);
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && x && dw, "head_backward_weight: null pointer");
  int kblocks = ceil_div(in_features, 256);
  int rpb = rows / (2 * kNumSM / (kblocks * members) + 1);
  if (rpb < 4) rpb = 4;
  if (rpb > 64) rpb = 64;
  dim3 grid(kblocks, ceil_div(rows, rpb), members);
  size_t smem = (size_t)rpb * out_features * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_features <= 1)
    launch_pdl(head_backward_weight_kernel<1>, grid, dim3(256), smem, st, dy, (long long)lddy, (long long)stride_dy, x,
               (long long)ldx, (long long)stride_x, dw, (long long)lddw, (long long)stride_dw, dbias, (long long)stride_db, rows,
               out_features, in_features, rpb);
  else if (out_features <= 8)
    launch_pdl(head_backward_weight_kernel<8>, grid, dim3(256), smem, st, dy, (long long)lddy, (long long)stride_dy, x,
               (long long)ldx, (long long)stride_x, dw, (long long)lddw, (long long)stride_dw, dbias, (long long)stride_db, rows,
               out_features, in_features, rpb);
  else if (out_features <= 16)
    launch_pdl(head_backward_weight_kernel<16>, grid, dim3(256), smem, st, dy, (long long)lddy, (long long)stride_dy, x,
               (long long)ldx, (long long)stride_x, dw, (long long)lddw, (long long)stride_dw, dbias, (long long)stride_db, rows,
               out_features, in_features, rpb);
  else
    launch_pdl(head_backward_weight_kernel<32>, grid, dim3(256), smem, st, dy, (long long)lddy, (long long)stride_dy, x,
               (long long)ldx, (long long)stride_x, dw, (long long)lddw, (long long)stride_dw, dbias, (long long)stride_db, rows,
               out_features, in_features, rpb);
  return check_launch("head_backward_weight");
}

// ---- _reduce_ensemble (q_functions/ensemble_q_function.py:9-24) over the member axis of q[members][n], and the
// ensemble TD error EnsembleQFunction.compute_error (ensemble_q_function.py:81-106) — the callable Q-function API.
namespace d3b {

__global__ void ensemble_reduce_kernel(const float* __restrict__ q, long long stride_member, int n, int members,
                                       int mode, float lam, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v0 = q[i];
  float mn = v0, mx = v0, sum = v0;
  for (int e = 1; e < members; ++e) {
    float v = q[(long long)e * stride_member + i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    sum = __fadd_rn(sum, v);
  }
  float r;
  if (mode == 0) r = mn;
  else if (mode == 1) r = mx;
  else if (mode == 2) r = __fdiv_rn(sum, (float)members);
  else r = __fadd_rn(__fmul_rn(lam, mn), __fmul_rn(1.0f - lam, mx));   // lam * min + (1 - lam) * max
  out[i] = r;
}

// out[0] = sum_e mean_b loss(q_e[b] - y_b),  y = r + gamma * target * (1 - terminal); gamma per row when gamma_rows.
// One block, fixed summation order (bit-reproducible).
__global__ void td_error_kernel(const float* __restrict__ q, long long stride_member, const float* __restrict__ rew,
                                const float* __restrict__ target, const float* __restrict__ term,
                                const float* __restrict__ gamma_rows, float gamma, int n, int members, int huber,
                                float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  float total = 0.f;
  for (int e = 0; e < members; ++e) {
    float part = 0.f;
    for (int b = threadIdx.x; b < n; b += blockDim.x) {
      float g = gamma_rows ? gamma_rows[b] : gamma;
      float y = __fadd_rn(rew[b], __fmul_rn(__fmul_rn(g, target[b]), __fsub_rn(1.0f, term[b])));
      float d = __fsub_rn(q[(long long)e * stride_member + b], y);
      float l;
      if (huber) {  // compute_huber_loss (q_functions/utility.py:27-32), beta = 1
        float a = fabsf(d);
        l = a < 1.0f ? 0.5f * d * d : a - 0.5f;
      } else {
        l = d * d;
      }
      part += l;
    }
    part = block_sum(part);
    if (threadIdx.x == 0) total += part / (float)n;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = total;
}

}  // namespace d3b

extern "C" int d3b_ensemble_reduce(const float* q, int64_t stride_member, int n, int members, int mode, float lam,
                                   float* out, void* stream) {
  D3B_REQUIRE(n >= 0 && members >= 1 && mode >= 0 && mode <= 3, "ensemble_reduce: bad arguments (mode 0 min, 1 max, 2 mean, 3 mix)");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(q && out, "ensemble_reduce: null pointer");
  launch_pdl(ensemble_reduce_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, q,
             (long long)stride_member, n, members, mode, lam, out);
  return check_launch("ensemble_reduce");
}

extern "C" int d3b_td_error(const float* q, int64_t stride_member, const float* rewards, const float* target,
                            const float* terminals, const float* gamma_rows, float gamma, int n, int members,
                            int huber, float* out, void* stream) {
  D3B_REQUIRE(n >= 1 && members >= 1, "td_error: bad sizes");
  D3B_REQUIRE(q && rewards && target && terminals && out, "td_error: null pointer");
  launch_pdl(td_error_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, q, (long long)stride_member, rewards,
             target, terminals, gamma_rows, gamma, n, members, huber, out);
  return check_launch("td_error");
}
