// K10: fused multi-tensor Adam (+ optional Polyak soft_sync in the same pass), hard_sync, step
// counters and the Philox noise fill.  HBM-bandwidth bound: one pass over flat parameter arenas
// (28 B/param for Adam, +8 B/param when the target arena is synced in the same pass).
//
// Adam arithmetic restates torch 2.11 `_single_tensor_adam`, the optimizer the
// reference builds through AdamFactory (d3rlpy/models/optimizers.py:106-138):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(1-b2^t) + eps;
//   p.addcdiv_(m, denom, value=-lr/(1-b1^t))
// soft_sync restates d3rlpy/torch_utility.py:27-33 (two separately-rounded steps).
#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"

namespace d3b {

// the bias corrections need double pow/sqrt: evaluate them once per block, not once per thread -- the two pow() chains
// (a few hundred dependent FP64 instructions each) on two warps side by side; the callers issue their first operand
// loads BEFORE this call, so the chains run under the memory latency instead of in front of it.  blockDim.x >= 64.
__device__ __forceinline__ void adam_scalars_block(const int* step, double lr, double b1, double b2, double eps,
                                                   float& w1, float& fb2, float& w2, float& feps, float& neg_ss,
                                                   float& bc2_sqrt) {
  __shared__ float sc[6];
  if (threadIdx.x == 0) {
    double bc1 = 1.0 - pow(b1, (double)*step);
    sc[0] = (float)(1.0 - b1);
    sc[3] = (float)eps;
    sc[4] = (float)(-(lr / bc1));
  } else if (threadIdx.x == 32) {
    double bc2 = 1.0 - pow(b2, (double)*step);
    sc[1] = (float)b2;
    sc[2] = (float)(1.0 - b2);
    sc[5] = (float)sqrt(bc2);
  }
  __syncthreads();
  w1 = sc[0]; fb2 = sc[1]; w2 = sc[2]; feps = sc[3]; neg_ss = sc[4]; bc2_sqrt = sc[5];
}

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float w1, float fb2, float w2,
                                          float feps, float neg_ss, float bc2_sqrt) {
  // lerp (weight < 0.5 branch of ATen's lerp; its vectorised CPU form is fmadd(w, g - m, m))
  m = __fmaf_rn(w1, __fsub_rn(g, m), m);
  v = __fmul_rn(v, fb2);
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(w2, g), g));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), feps);
  return __fadd_rn(p, __fdiv_rn(__fmul_rn(neg_ss, m), denom));
}

// grid-stride, float4 main body + scalar tail.  `targ` (nullable) gets soft-synced with the NEW p.
// grads are zeroed after use so the next backward can accumulate with atomics (wgrad split-K).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   float* __restrict__ targ, long long n, const int* step, double lr,
                                                   double b1, double b2, double eps, float tau, int zero_grad,
                                                   float weight_decay) {
  pdl_trigger();
  pdl_wait();
  long long n4 = n >> 2;
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  // operands of the first pass in flight while the bias corrections are evaluated
  long long i = tid;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 P = zero4, G = zero4, M = zero4, V = zero4, T = zero4;
  if (i < n4) {
    P = ((float4*)p)[i]; G = ((float4*)g)[i]; M = ((float4*)m)[i]; V = ((float4*)v)[i];
    if (targ) T = ((float4*)targ)[i];
  }
  float w1, fb2, w2, feps, neg_ss, bc2s;
  adam_scalars_block(step, lr, b1, b2, eps, w1, fb2, w2, feps, neg_ss, bc2s);
  float one_m_tau = (float)(1.0 - (double)tau);  // python: (1 - tau) in double, then cast by mul_
  while (i < n4) {
    if (weight_decay != 0.f) {   // torch.optim.Adam: grad = grad.add(param, alpha=weight_decay) (L2, not decoupled)
      G.x = __fmaf_rn(weight_decay, P.x, G.x); G.y = __fmaf_rn(weight_decay, P.y, G.y);
      G.z = __fmaf_rn(weight_decay, P.z, G.z); G.w = __fmaf_rn(weight_decay, P.w, G.w);
    }
    P.x = adam_one(P.x, G.x, M.x, V.x, w1, fb2, w2, feps, neg_ss, bc2s);
    P.y = adam_one(P.y, G.y, M.y, V.y, w1, fb2, w2, feps, neg_ss, bc2s);
    P.z = adam_one(P.z, G.z, M.z, V.z, w1, fb2, w2, feps, neg_ss, bc2s);
    P.w = adam_one(P.w, G.w, M.w, V.w, w1, fb2, w2, feps, neg_ss, bc2s);
    ((float4*)p)[i] = P;
    ((float4*)m)[i] = M;
    ((float4*)v)[i] = V;
    if (zero_grad) ((float4*)g)[i] = zero4;
    if (targ) {
      T.x = __fadd_rn(__fmul_rn(T.x, one_m_tau), __fmul_rn(tau, P.x));
      T.y = __fadd_rn(__fmul_rn(T.y, one_m_tau), __fmul_rn(tau, P.y));
      T.z = __fadd_rn(__fmul_rn(T.z, one_m_tau), __fmul_rn(tau, P.z));
      T.w = __fadd_rn(__fmul_rn(T.w, one_m_tau), __fmul_rn(tau, P.w));
      ((float4*)targ)[i] = T;
    }
    i += stride;
    if (i < n4) {
      P = ((float4*)p)[i]; G = ((float4*)g)[i]; M = ((float4*)m)[i]; V = ((float4*)v)[i];
      if (targ) T = ((float4*)targ)[i];
    }
  }
  for (long long i = (n4 << 2) + tid; i < n; i += stride) {
    float M = m[i], V = v[i];
    float P = adam_one(p[i], weight_decay != 0.f ? __fmaf_rn(weight_decay, p[i], g[i]) : g[i], M, V, w1, fb2, w2, feps,
                       neg_ss, bc2s);
    p[i] = P;
    m[i] = M;
    v[i] = V;
    if (zero_grad) g[i] = 0.f;
    if (targ) targ[i] = __fadd_rn(__fmul_rn(targ[i], one_m_tau), __fmul_rn(tau, P));
  }
}


// Adam (+ Polyak) pass that also refreshes the bf16 K-major weight shadows the tensor-core kernels read
// (row-major W [N][ld8(K)] per trunk layer, for the params and — when synced — the target arena), so no
// separate conversion launch is needed after the optimizer step.
constexpr int MAX_SEG = 8;
struct ShadowSegs {
  long long param_off[MAX_SEG];  // fp32 offset inside one member block (multiple of 4)
  long long count[MAX_SEG];      // rows * cols
  int cols[MAX_SEG];
  long long shadow_off[MAX_SEG]; // bf16 offset inside one member's shadow block
  int ld[MAX_SEG];
  int n;
  long long member_size, shadow_member;
};

__global__ void __launch_bounds__(256) adam_shadow_kernel(float* __restrict__ p, float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          float* __restrict__ targ, long long n, const int* step,
                                                          double lr, double b1, double b2, double eps, float tau,
                                                          __nv_bfloat16* __restrict__ sh_p,
                                                          __nv_bfloat16* __restrict__ sh_t, ShadowSegs segs) {
  pdl_trigger();
  pdl_wait();
  long long n4 = n >> 2;  // arenas are padded to multiples of 4 floats per member
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  // operands of the first pass in flight while the bias corrections are evaluated
  long long i = tid;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 P = zero4, G = zero4, M = zero4, V = zero4, T0 = zero4;
  if (i < n4) {
    P = ((float4*)p)[i]; G = ((float4*)g)[i]; M = ((float4*)m)[i]; V = ((float4*)v)[i];
    if (targ) T0 = ((float4*)targ)[i];
  }
  float w1, fb2, w2, feps, neg_ss, bc2s;
  adam_scalars_block(step, lr, b1, b2, eps, w1, fb2, w2, feps, neg_ss, bc2s);
  float one_m_tau = (float)(1.0 - (double)tau);
  for (; i < n4; i += stride) {
    if (i != tid) {   // later passes (arenas beyond one grid of float4s)
      P = ((float4*)p)[i]; G = ((float4*)g)[i]; M = ((float4*)m)[i]; V = ((float4*)v)[i];
      if (targ) T0 = ((float4*)targ)[i];
    }
    P.x = adam_one(P.x, G.x, M.x, V.x, w1, fb2, w2, feps, neg_ss, bc2s);
    P.y = adam_one(P.y, G.y, M.y, V.y, w1, fb2, w2, feps, neg_ss, bc2s);
    P.z = adam_one(P.z, G.z, M.z, V.z, w1, fb2, w2, feps, neg_ss, bc2s);
    P.w = adam_one(P.w, G.w, M.w, V.w, w1, fb2, w2, feps, neg_ss, bc2s);
    ((float4*)p)[i] = P;
    ((float4*)m)[i] = M;
    ((float4*)v)[i] = V;
    ((float4*)g)[i] = zero4;
    float4 T = P;
    if (targ) {
      T = T0;
      T.x = __fadd_rn(__fmul_rn(T.x, one_m_tau), __fmul_rn(tau, P.x));
      T.y = __fadd_rn(__fmul_rn(T.y, one_m_tau), __fmul_rn(tau, P.y));
      T.z = __fadd_rn(__fmul_rn(T.z, one_m_tau), __fmul_rn(tau, P.z));
      T.w = __fadd_rn(__fmul_rn(T.w, one_m_tau), __fmul_rn(tau, P.w));
      ((float4*)targ)[i] = T;
    }
    // shadow refresh: the 4 elements belong to one allocation of one member (allocations are 4-aligned)
    long long e0 = i << 2;
    long long member = e0 / segs.member_size, off = e0 - member * segs.member_size;
    int sidx = -1;
#pragma unroll
    for (int k = 0; k < MAX_SEG; ++k)
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
          sh_p[d] = __float2bfloat16_rn(pv[q]);
          if (sh_t && targ) sh_t[d] = __float2bfloat16_rn(tv[q]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) soft_sync_kernel(float* __restrict__ targ, const float* __restrict__ p,
                                                        long long n, float tau) {
  float one_m_tau = (float)(1.0 - (double)tau);
  long long n4 = n >> 2;
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += stride) {
    float4 T = ((float4*)targ)[i];
    float4 P = __ldg((const float4*)p + i);
    T.x = __fadd_rn(__fmul_rn(T.x, one_m_tau), __fmul_rn(tau, P.x));
    T.y = __fadd_rn(__fmul_rn(T.y, one_m_tau), __fmul_rn(tau, P.y));
    T.z = __fadd_rn(__fmul_rn(T.z, one_m_tau), __fmul_rn(tau, P.z));
    T.w = __fadd_rn(__fmul_rn(T.w, one_m_tau), __fmul_rn(tau, P.w));
    ((float4*)targ)[i] = T;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += stride)
    targ[i] = __fadd_rn(__fmul_rn(targ[i], one_m_tau), __fmul_rn(tau, __ldg(p + i)));
}

__global__ void tick_kernel(int* counters, int n, unsigned mask) {
  int i = threadIdx.x;
  if (i < n && ((mask >> i) & 1u)) counters[i] += 1;
}

// segment layout: first n_normal floats ~ N(0,1), next n_uniform floats ~ U(-1,1).
// counter = (*draw_counter) so that graph replays advance the stream; bumped by tick_kernel.
__global__ void __launch_bounds__(256) noise_fill_kernel(float* __restrict__ out, long long n_normal,
                                                         long long n_uniform, unsigned long long seed,
                                                         const int* draw_counter) {
  pdl_trigger();
  pdl_wait();
  long long n = n_normal + n_uniform;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one Philox call -> 4 floats
  long long stride = (long long)gridDim.x * blockDim.x;
  uint32_t epoch = (uint32_t)(*draw_counter);
  for (; (q << 2) < n; q += stride) noise_quad(out, q, n_normal, n, seed, epoch);
}

}  // namespace d3b

using namespace d3b;

static int grid_for(long long n, int per_thread) {
  long long blocks = ceil_div_ll(n, 256LL * per_thread);
  long long cap = (long long)kNumSM * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int d3b_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* target,
                             int64_t n, const int* step, double lr, double beta1, double beta2, double eps,
                             float tau, int zero_grad, void* stream) {
  D3B_REQUIRE(n >= 0, "adam_step: n < 0");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(params && grads && exp_avg && exp_avg_sq && step, "adam_step: null pointer");
  D3B_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq |
               (uintptr_t)target) % 16 == 0,
              "adam_step: arenas must be 16-byte aligned");
  launch_pdl(adam_kernel, dim3(grid_for(n, 4)), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq,
             target, (long long)n, step, lr, beta1, beta2, eps, tau, zero_grad, 0.f);
  return check_launch("adam_step");
}

// adam_step with torch.optim.Adam's weight_decay (L2 term added to the gradient): AWAC's actor optimizer
// (d3rlpy/algos/awac.py:105, AdamFactory(weight_decay=1e-4))
extern "C" int d3b_adam_step_wd(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* target,
                                int64_t n, const int* step, double lr, double beta1, double beta2, double eps,
                                float weight_decay, float tau, int zero_grad, void* stream) {
  D3B_REQUIRE(n >= 0 && weight_decay >= 0.f, "adam_step_wd: bad arguments");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(params && grads && exp_avg && exp_avg_sq && step, "adam_step_wd: null pointer");
  D3B_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq |
               (uintptr_t)target) % 16 == 0,
              "adam_step_wd: arenas must be 16-byte aligned");
  launch_pdl(adam_kernel, dim3(grid_for(n, 4)), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq,
             target, (long long)n, step, lr, beta1, beta2, eps, tau, zero_grad, weight_decay);
  return check_launch("adam_step_wd");
}

// adam_step + bf16 shadow refresh in one pass.  table_host: n_segments x {param_off, rows, cols, shadow_off, ld}
// (int64), one entry per trunk weight matrix; shadow_target may be NULL.
extern "C" int d3b_adam_step_shadow(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* target,
                                    int64_t n, const int* step, double lr, double beta1, double beta2, double eps,
                                    float tau, void* shadow_params, void* shadow_target, const int64_t* table_host,
                                    int n_segments, int64_t member_size, int64_t shadow_member, void* stream) {
  D3B_REQUIRE(n >= 0 && n % 4 == 0 && member_size > 0 && member_size % 4 == 0, "adam_step_shadow: arenas must be padded to 4 floats");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(params && grads && exp_avg && exp_avg_sq && step && shadow_params && table_host,
              "adam_step_shadow: null pointer");
  D3B_REQUIRE(n_segments >= 1 && n_segments <= MAX_SEG, "adam_step_shadow: 1..8 shadow segments");
  D3B_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq |
               (uintptr_t)target) % 16 == 0,
              "adam_step_shadow: arenas must be 16-byte aligned");
  ShadowSegs segs{};
  segs.n = n_segments; segs.member_size = member_size; segs.shadow_member = shadow_member;
  for (int k = 0; k < n_segments; ++k) {
    const int64_t* t = table_host + 5 * k;
    D3B_REQUIRE(t[0] % 4 == 0 && t[1] > 0 && t[2] > 0 && t[4] >= t[2], "adam_step_shadow: bad segment");
    segs.param_off[k] = t[0]; segs.count[k] = t[1] * t[2]; segs.cols[k] = (int)t[2];
    segs.shadow_off[k] = t[3]; segs.ld[k] = (int)t[4];
  }
  launch_pdl(adam_shadow_kernel, dim3(grid_for(n, 4)), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq,
             target, (long long)n, step, lr, beta1, beta2, eps, tau, (__nv_bfloat16*)shadow_params,
             (__nv_bfloat16*)shadow_target, segs);
  return check_launch("adam_step_shadow");
}

extern "C" int d3b_soft_sync(float* target, const float* params, int64_t n, float tau, void* stream) {
  D3B_REQUIRE(n >= 0, "soft_sync: n < 0");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(target && params, "soft_sync: null pointer");
  D3B_REQUIRE(((uintptr_t)target | (uintptr_t)params) % 16 == 0, "soft_sync: arenas must be 16-byte aligned");
  soft_sync_kernel<<<grid_for(n, 4), 256, 0, (cudaStream_t)stream>>>(target, params, n, tau);
  return check_launch("soft_sync");
}

extern "C" int d3b_hard_sync(float* target, const float* params, int64_t n, void* stream) {
  D3B_REQUIRE(n >= 0, "hard_sync: n < 0");
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(target && params, "hard_sync: null pointer");
  D3B_CUDA(cudaMemcpyAsync(target, params, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return D3B_OK;
}

extern "C" int d3b_tick(int* counters, int n, unsigned mask, void* stream) {
  D3B_REQUIRE(counters && n > 0 && n <= 32, "tick: need 1..32 counters");
  tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counters, n, mask);
  return check_launch("tick");
}

extern "C" int d3b_noise_fill(float* out, int64_t n_normal, int64_t n_uniform, uint64_t seed,
                              const int* draw_counter, void* stream) {
  D3B_REQUIRE(n_normal >= 0 && n_uniform >= 0, "noise_fill: negative count");
  long long n = n_normal + n_uniform;
  if (n == 0) return D3B_OK;
  D3B_REQUIRE(out && draw_counter, "noise_fill: null pointer");
  launch_pdl(noise_fill_kernel, dim3(grid_for(n, 4)), dim3(256), 0, (cudaStream_t)stream, out, (long long)n_normal,
             (long long)n_uniform, (unsigned long long)seed, draw_counter);
  return check_launch("noise_fill");
}
