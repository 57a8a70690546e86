// K1 / K1b: HBM-resident replay buffer -> minibatch gather (vector rows, uint8 frame stacks,
// n-step return, terminal masks).  Bit-exact restatement, in flat-index form, of the reference's
// linked-list walk:
//   _assign_to_batch / _assign_observation / _assign_action  d3rlpy/dataset.pyx:1219-1342
//   _stack_frames                                            d3rlpy/dataset.pyx:1051-1096
// Data layout (DESIGN.md §HBM layout): step-indexed arrays O[S,...], A[S,...], R[S] plus one int4
// per TRANSITION {step, episode_start_step, episode_last_transition_step, flags}; flags bit 0 = the transition's
// `terminal`, bit 1 = its next_observation is the all-zero dummy.  Episode-built data sets both on the last transition
// of a terminal episode (dataset.pyx:86-96); the online ReplayBuffer also marks the transition INTO the terminal state
// as terminal while it keeps the real next observation (online/buffers.py:283-300), hence two bits.
#include "common.cuh"

namespace d3b {

struct RowInfo {
  int g, start, k, g2, terminal, zero_next;
};

__device__ __forceinline__ RowInfo row_info(const int4* __restrict__ meta, long long t, int n_steps) {
  int4 m = __ldg(meta + t);
  RowInfo r;
  r.g = m.x;
  r.start = m.y;
  int remain = m.z - m.x + 1;
  r.k = n_steps < remain ? n_steps : remain;
  r.g2 = r.g + r.k - 1;
  int flags = (r.k == 1) ? m.w : __ldg(meta + t + (r.k - 1)).w;
  r.terminal = flags & 1;
  r.zero_next = (flags >> 1) & 1;  // plain copies take next_observation verbatim (dataset.pyx:1243-1264)
  return r;
}

// n-step return: float accumulator, pow() in double, one rounding per step
// (dataset.pyx:1322-1330; `float n_step_return`, `gamma ** i` with float gamma -> double pow).
__device__ __forceinline__ float nstep_return(const float* __restrict__ rewards, int g, int k, float gamma) {
  float acc = 0.f;
  double gd = (double)gamma;
  double p = 1.0;
  for (int i = 0; i < k; ++i) {
    if (i > 0) p = pow(gd, (double)i);
    acc = (float)((double)acc + (double)__ldg(rewards + g + i) * p);
  }
  return acc;
}

// LPR lanes per minibatch row (16 = half-warp, 8 = quarter-warp): the fewer lanes per row, the more independent
// index -> metadata -> row chains are in flight per SM (measured on 1 M random c5 rows: 31 % of the HBM peak with a warp
// per row, 60 % with a half-warp, 72 % with a quarter-warp and 8 columns per lane; 4 or 2 lanes per row and 16 columns
// per lane are slower again, profiles/r1_ubench.md).  Every lane keeps 2*U loads in flight per pass of LPR*U columns.
template <int LPR, int U>
__global__ void __launch_bounds__(128) gather_vector_kernel(
    const float* __restrict__ obs, int O, const void* __restrict__ actions, int A, int discrete,
    const float* __restrict__ rewards, const int4* __restrict__ meta, const long long* __restrict__ indices,
    int B, int n_steps, float gamma, float* __restrict__ out_obs, void* __restrict__ out_act,
    float* __restrict__ out_rew, float* __restrict__ out_next, float* __restrict__ out_term,
    float* __restrict__ out_n, const float* __restrict__ sc_mean, const float* __restrict__ sc_std,
    float sc_eps) {
  int lane = threadIdx.x & (LPR - 1);
  int b = blockIdx.x * (blockDim.x / LPR) + threadIdx.x / LPR;
  if (b >= B) return;
  long long t = __ldg(indices + b);
  RowInfo r = row_info(meta, t, n_steps);
  if (O > 0) {
    const float* src = obs + (size_t)r.g * O;
    const float* nsrc = obs + (size_t)(r.g2 + 1) * O;
    float* d0 = out_obs + (size_t)b * O;
    float* d1 = out_next + (size_t)b * O;
    // 4 x 16 columns per pass with every load issued before the first store (8 independent loads in flight per
    // lane: the row gather is latency-bound on its index -> metadata -> row dependency chain)
    for (int j0 = 0; j0 < O; j0 += LPR * U) {
      float xs[U], ys[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int j = j0 + lane + LPR * u;
        xs[u] = j < O ? __ldg(src + j) : 0.f;
        ys[u] = (j < O && !r.zero_next) ? __ldg(nsrc + j) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int j = j0 + lane + LPR * u;
        if (j < O) {
          float x = xs[u], y = ys[u];
          if (sc_mean) {  // StandardScaler.transform fused (preprocessing/scalers.py:350-354)
            float m = __ldg(sc_mean + j), s = __ldg(sc_std + j) + sc_eps;
            x = __fdiv_rn(__fsub_rn(x, m), s);
            y = __fdiv_rn(__fsub_rn(y, m), s);
          }
          d0[j] = x;
          d1[j] = y;
        }
      }
    }
  }
  if (discrete) {
    if (lane == 0) ((int*)out_act)[b] = __ldg((const int*)actions + r.g);
  } else {
    const float* a = (const float*)actions + (size_t)r.g * A;
    for (int j = lane; j < A; j += LPR) ((float*)out_act)[(size_t)b * A + j] = __ldg(a + j);
  }
  if (lane == 0) {
    out_rew[b] = nstep_return(rewards, r.g, r.k, gamma);
    out_term[b] = (float)r.terminal;
    out_n[b] = (float)r.k;
  }
}

// One CTA per (row, obs|next, frame slot): copies one c*h*w uint8 frame with 16-byte vectors.
template <bool VEC16>
__global__ void __launch_bounds__(128) gather_frames_kernel(
    const uint8_t* __restrict__ frames, int frame_bytes, const int4* __restrict__ meta,
    const long long* __restrict__ indices, int B, int n_frames, int n_steps, uint8_t* __restrict__ out_obs,
    uint8_t* __restrict__ out_next) {
  int slot = blockIdx.x;
  int f = slot % n_frames;
  int which = (slot / n_frames) & 1;
  int b = slot / (2 * n_frames);
  if (b >= B) return;
  long long t = __ldg(indices + b);
  RowInfo r = row_info(meta, t, n_steps);
  int back = n_frames - 1 - f;  // channel block f holds the frame `back` steps in the past (oldest first)
  uint8_t* dst = (which ? out_next : out_obs) + ((size_t)b * n_frames + f) * frame_bytes;
  // stacking: terminal => all-zero next stack (dataset.pyx:1066-1068); single frames are copied verbatim
  bool zero = which && (n_frames > 1 ? r.terminal : r.zero_next);
  int j = (which ? r.g2 + 1 : r.g) - back;
  if (j < r.start) j = r.start;     // episode start repeated as padding (dataset.pyx:1089-1095)
  const uint8_t* src = frames + (size_t)j * frame_bytes;
  if (VEC16) {
    int nv = frame_bytes >> 4;
    const uint4* s4 = (const uint4*)src;
    uint4* d4 = (uint4*)dst;
    // 4 independent 16-byte loads in flight per thread before the first store (an 84x84 frame = 441 vectors is one
    // pass of a 128-thread CTA); 16 such CTAs are resident per SM
    for (int i0 = threadIdx.x; i0 < nv; i0 += 4 * blockDim.x) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int i = i0 + u * blockDim.x;
        v[u] = (i < nv && !zero) ? __ldg(s4 + i) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int i = i0 + u * blockDim.x;
        if (i < nv) d4[i] = v[u];
      }
    }
  } else {
    for (int i = threadIdx.x; i < frame_bytes; i += blockDim.x) dst[i] = zero ? 0 : __ldg(src + i);
  }
}

}  // namespace d3b

using namespace d3b;

extern "C" int d3b_gather_vector(const float* obs, int obs_dim, const void* actions, int act_dim, int discrete,
                                 const float* rewards, const void* meta, const int64_t* indices, int batch,
                                 int n_steps, float gamma, float* out_obs, void* out_act, float* out_rew,
                                 float* out_next, float* out_term, float* out_nsteps, const float* scaler_mean,
                                 const float* scaler_std, float scaler_eps, void* stream) {
  D3B_REQUIRE(batch >= 0 && n_steps >= 1 && obs_dim >= 0, "gather_vector: bad sizes B=%d n_steps=%d O=%d", batch,
              n_steps, obs_dim);
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(meta && indices && rewards && out_rew && out_term && out_nsteps && out_act && actions,
              "gather_vector: null pointer");
  D3B_REQUIRE(obs_dim == 0 || (obs && out_obs && out_next), "gather_vector: null observation pointer");
  D3B_REQUIRE((scaler_mean == nullptr) == (scaler_std == nullptr), "gather_vector: scaler mean/std must come together");
  cudaStream_t st = (cudaStream_t)stream;
  if (batch >= 128 * kNumSM) {
    // bandwidth regime (every SM holds many blocks): quarter-warp per row, 16 loads in flight per lane
    gather_vector_kernel<8, 8><<<ceil_div(batch, 16), 128, 0, st>>>(
        obs, obs_dim, actions, act_dim, discrete, rewards, (const int4*)meta, (const long long*)indices, batch, n_steps,
        gamma, out_obs, out_act, out_rew, out_next, out_term, out_nsteps, scaler_mean, scaler_std, scaler_eps);
  } else {
    // latency regime (batch-256 ... batch-8192 updates): half-warp per row, fewest serial passes over a row
    gather_vector_kernel<16, 4><<<ceil_div(batch, 8), 128, 0, st>>>(
        obs, obs_dim, actions, act_dim, discrete, rewards, (const int4*)meta, (const long long*)indices, batch, n_steps,
        gamma, out_obs, out_act, out_rew, out_next, out_term, out_nsteps, scaler_mean, scaler_std, scaler_eps);
  }
  return check_launch("gather_vector");
}

extern "C" int d3b_gather_frames(const uint8_t* frames, int frame_bytes, const void* meta, const int64_t* indices,
                                 int batch, int n_frames, int n_steps, uint8_t* out_obs, uint8_t* out_next,
                                 void* stream) {
  D3B_REQUIRE(batch >= 0 && n_frames >= 1 && n_steps >= 1 && frame_bytes > 0, "gather_frames: bad sizes");
  if (batch == 0) return D3B_OK;
  D3B_REQUIRE(frames && meta && indices && out_obs && out_next, "gather_frames: null pointer");
  bool vec = (frame_bytes % 16 == 0) && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)out_obs % 16 == 0) &&
             ((uintptr_t)out_next % 16 == 0);
  int grid = batch * 2 * n_frames;
  if (vec)
    gather_frames_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(frames, frame_bytes, (const int4*)meta,
                                                                       (const long long*)indices, batch, n_frames,
                                                                       n_steps, out_obs, out_next);
  else
    gather_frames_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(frames, frame_bytes, (const int4*)meta,
                                                                        (const long long*)indices, batch, n_frames,
                                                                        n_steps, out_obs, out_next);
  return check_launch("gather_frames");
}
