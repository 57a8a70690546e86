// bf16-mode support kernels around the tcgen05 GEMM (umma_gemm.cu): fp32 -> bf16 shadows (+ transposed
// copies so that every GEMM operand is K-major), narrow heads reading bf16 activations, bias-gradient
// column sums.  All HBM/L2-bandwidth bound element-wise or reduction work.
#include <cuda_bf16.h>

#include "common.cuh"

namespace d3b {

struct ShadowEntry {
  long long src_off;  // fp32 offset inside one member block
  int rows, cols;     // matrix [rows][cols], src leading dim == cols
  long long dst_off, ldd;    // bf16 row-major copy (ld >= cols), -1 = skip
  long long dstT_off, ldt;   // bf16 transposed copy [cols][ldt], -1 = skip
};
constexpr int MAX_SHADOW = 16;
struct ShadowTable {
  ShadowEntry e[MAX_SHADOW];
  int n;
};

// grid: (32x32 tiles of the largest matrix, table entry, member).  Tiles go through shared memory so that the
// row-major copy and the transposed copy are both written with coalesced rows (the element-wise version's
// transposed stores were 2-byte scatters: 35 us for the 512 x 3136 fc layer of the Nature-DQN encoder).
__global__ void __launch_bounds__(256) shadow_kernel(const float* __restrict__ src, long long src_member_stride,
                                                     __nv_bfloat16* __restrict__ dst, long long dst_member_stride,
                                                     ShadowTable t) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const ShadowEntry& s = t.e[blockIdx.y];
  const int tiles_c = (s.cols + 31) >> 5, tiles_r = (s.rows + 31) >> 5;
  const float* sp = src + (long long)blockIdx.z * src_member_stride + s.src_off;
  __nv_bfloat16* dp = dst + (long long)blockIdx.z * dst_member_stride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int tile_id = blockIdx.x; tile_id < tiles_c * tiles_r; tile_id += gridDim.x) {
    const int r0 = (tile_id / tiles_c) << 5, c0 = (tile_id % tiles_c) << 5;
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      const float v = (r < s.rows && c < s.cols) ? __ldg(sp + (long long)r * s.cols + c) : 0.f;
      tile[i][tx] = v;
      if (s.dst_off >= 0 && r < s.rows && c < s.cols) dp[s.dst_off + (long long)r * s.ldd + c] = __float2bfloat16_rn(v);
    }
    if (s.dstT_off >= 0) {
      __syncthreads();
      for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < s.rows && c < s.cols) dp[s.dstT_off + (long long)c * s.ldt + r] = __float2bfloat16_rn(tile[tx][i]);
      }
    }
    __syncthreads();
  }
}

// fp32 [rows][cols] (ld lds) -> bf16 [rows][ldd] and/or transposed bf16 [cols][ldt]; 32x32 smem tile transpose.
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ src, long long lds, int rows, int cols,
                                                      __nv_bfloat16* __restrict__ dst, long long ldd,
                                                      __nv_bfloat16* __restrict__ dstT, long long ldt) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int r = r0 + i, c = c0 + tx;
    float v = (r < rows && c < cols) ? __ldg(src + (long long)r * lds + c) : 0.f;
    tile[i][tx] = v;
    if (dst && r < rows && c < cols) dst[(long long)r * ldd + c] = __float2bfloat16_rn(v);
  }
  if (!dstT) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dstT[(long long)c * ldt + r] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// One warp per (member,row): Y[e][m][n] = act(X_bf16[e][m][:] . W[e][n][:] + b[e][n]), fp32 weights.
// The activation row is read ONCE into registers (KPL values per lane); four outputs are accumulated at a time
// with all their weight loads issued together (4*KPL independent loads in flight per lane).
template <int KPL>
__global__ void __launch_bounds__(256) head_forward_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long ldx,
                                                                long long sX, const float* __restrict__ W,
                                                                long long ldw, long long sW,
                                                                const float* __restrict__ bias, long long sB,
                                                                float* __restrict__ Y, long long ldy, long long sY,
                                                                int M, int N, int K, int E, int act_tanh) {
  int lane = threadIdx.x & 31;
  long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)M * E) return;
  int e = (int)(wid / M), m = (int)(wid % M);
  const __nv_bfloat16* x = X + (long long)e * sX + (long long)m * ldx;
  const float* w = W + (long long)e * sW;
  float mine = 0.f;
  if (K <= 32 * KPL) {
    float xr[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      xr[i] = k < K ? __bfloat162float(x[k]) : 0.f;
    }
    for (int n0 = 0; n0 < N; n0 += 4) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* w0 = w + (long long)n0 * ldw;
      const float* w1 = w + (long long)(n0 + 1 < N ? n0 + 1 : n0) * ldw;
      const float* w2 = w + (long long)(n0 + 2 < N ? n0 + 2 : n0) * ldw;
      const float* w3 = w + (long long)(n0 + 3 < N ? n0 + 3 : n0) * ldw;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        int k = lane + 32 * i;
        if (k < K) {
          s0 = fmaf(xr[i], __ldg(w0 + k), s0);
          s1 = fmaf(xr[i], __ldg(w1 + k), s1);
          s2 = fmaf(xr[i], __ldg(w2 + k), s2);
          s3 = fmaf(xr[i], __ldg(w3 + k), s3);
        }
      }
      s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
      if (lane == n0) mine = s0;
      if (lane == n0 + 1) mine = s1;
      if (lane == n0 + 2) mine = s2;
      if (lane == n0 + 3) mine = s3;
    }
  } else {
    for (int n = 0; n < N; ++n) {
      const float* wr = w + (long long)n * ldw;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s = fmaf(__bfloat162float(x[k]), __ldg(wr + k), s);
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

// dX_bf16[e][m][k] = (sum_n dY[e][m][n] W[e][n][k]) * [src_bf16[e][m][k] > 0]  (+ transposed copy [k][m])
// 32x32 tiles so that both the row-major and the transposed store are coalesced.
__global__ void __launch_bounds__(256) head_backward_data_bf16_kernel(
    const float* __restrict__ dY, long long lddy, long long sdY, const float* __restrict__ W, long long ldw,
    long long sW, __nv_bfloat16* __restrict__ dX, long long lddx, long long sdX, __nv_bfloat16* __restrict__ dXT,
    long long ldt, long long sdXT, const __nv_bfloat16* __restrict__ src, long long ldsrc, long long sSrc, int M,
    int N, int K) {
  __shared__ float tile[32][33];
  int e = blockIdx.z;
  int m0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* w = W + (long long)e * sW;
  for (int i = ty; i < 32; i += 8) {
    int m = m0 + i, k = k0 + tx;
    float s = 0.f;
    if (m < M && k < K) {
      const float* dy = dY + (long long)e * sdY + (long long)m * lddy;
      for (int n = 0; n < N; ++n) s = fmaf(__ldg(dy + n), __ldg(w + (long long)n * ldw + k), s);
      if (src && !(__bfloat162float(src[(long long)e * sSrc + (long long)m * ldsrc + k]) > 0.f)) s = 0.f;
      dX[(long long)e * sdX + (long long)m * lddx + k] = __float2bfloat16_rn(s);
    }
    tile[i][tx] = s;
  }
  if (!dXT) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int k = k0 + i, m = m0 + tx;
    if (m < M && k < K) dXT[(long long)e * sdXT + (long long)k * ldt + m] = __float2bfloat16_rn(tile[tx][i]);
  }
}

template <int NMAX>
__global__ void __launch_bounds__(256) head_backward_weight_bf16_kernel(
    const float* __restrict__ dY, long long lddy, long long sdY, const __nv_bfloat16* __restrict__ X, long long ldx,
    long long sX, float* __restrict__ dW, long long lddw, long long sdW, float* __restrict__ db, long long sdb, int M,
    int N, int K, int rows_per_block) {
  extern __shared__ float sdy[];
  int e = blockIdx.z;
  int m0 = blockIdx.y * rows_per_block;
  int rows = min(rows_per_block, M - m0);
  const float* dy = dY + (long long)e * sdY + (long long)m0 * lddy;
  for (int i = threadIdx.x; i < rows * N; i += blockDim.x) sdy[i] = __ldg(dy + (long long)(i / N) * lddy + (i % N));
  __syncthreads();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) {
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    const __nv_bfloat16* x = X + (long long)e * sX + (long long)m0 * ldx + k;
    for (int r = 0; r < rows; ++r) {
      float xv = __bfloat162float(x[(long long)r * ldx]);
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

// db[e][n] += sum_m dZ_bf16[e][m][n]; grid (ceil(N/256), row chunks, E)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dZ, long long ld,
                                                          long long sZ, float* __restrict__ db, long long sdb, int M,
                                                          int N, int rows_per_block) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int m0 = blockIdx.y * rows_per_block;
  int m1 = min(M, m0 + rows_per_block);
  const __nv_bfloat16* z = dZ + (long long)blockIdx.z * sZ + n;
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += __bfloat162float(z[(long long)m * ld]);
  atomicAdd(db + (long long)blockIdx.z * sdb + n, s);
}

// Wide inputs (BCQ's 750 / 300-feature heads over 25 600 rows): each lane reads 16-byte chunks of 8 bf16 activations
// and the matching weights as four 8-byte loads (the fp32 head rows are only 8-byte aligned when K is even), two
// outputs at a time.  The element-wise kernel above issues one 2-byte activation load and one 4-byte weight load per
// FMA.  Used where it measures faster (see d3b_head_forward_bf16).
template <int NCH>
__global__ void __launch_bounds__(256) head_forward_bf16_vec_kernel(const __nv_bfloat16* __restrict__ X, long long ldx,
                                                                    long long sX, const float* __restrict__ W,
                                                                    long long ldw, long long sW,
                                                                    const float* __restrict__ bias, long long sB,
                                                                    float* __restrict__ Y, long long ldy, long long sY,
                                                                    int M, int N, int K, int E, int act_tanh) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)M * E) return;
  const int e = (int)(wid / M), m = (int)(wid % M);
  const __nv_bfloat16* x = X + (long long)e * sX + (long long)m * ldx;
  const float* w = W + (long long)e * sW;
  const int nfull = K >> 3, tail = K & 7;
  float xr[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nfull) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + 8 * c));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        xr[i][2 * j] = f.x;
        xr[i][2 * j + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[i][j] = 0.f;
    }
  }
  const int kt = 8 * nfull + lane;
  const float xt = lane < tail ? __bfloat162float(x[kt]) : 0.f;
  float mine = 0.f;
  for (int n0 = 0; n0 < N; n0 += 2) {
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + j;
      if (n < N) {
        const float* wr = w + (long long)n * ldw;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = lane + 32 * i;
          if (c < nfull) {
            const float2* wp = reinterpret_cast<const float2*>(wr + 8 * c);
            const float2 a0 = __ldg(wp), a1 = __ldg(wp + 1), a2 = __ldg(wp + 2), a3 = __ldg(wp + 3);
            s[j] = fmaf(xr[i][0], a0.x, s[j]); s[j] = fmaf(xr[i][1], a0.y, s[j]);
            s[j] = fmaf(xr[i][2], a1.x, s[j]); s[j] = fmaf(xr[i][3], a1.y, s[j]);
            s[j] = fmaf(xr[i][4], a2.x, s[j]); s[j] = fmaf(xr[i][5], a2.y, s[j]);
            s[j] = fmaf(xr[i][6], a3.x, s[j]); s[j] = fmaf(xr[i][7], a3.y, s[j]);
          }
        }
        if (lane < tail) s[j] = fmaf(xt, __ldg(wr + kt), s[j]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
      s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
    }
    if (lane == n0) mine = s[0];
    if (lane == n0 + 1) mine = s[1];
  }
  if (lane < N) {
    float v = mine + (bias ? __ldg(bias + (long long)e * sB + lane) : 0.f);
    if (act_tanh) v = tanhf(v);
    Y[(long long)e * sY + (long long)m * ldy + lane] = v;
  }
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

// table: int64[n][7] = {src_off, rows, cols, dst_off, ldd, dstT_off, ldt} (host memory)
extern "C" int d3b_shadow_weights(const float* src, int64_t src_member_stride, void* dst_bf16,
                                  int64_t dst_member_stride, const int64_t* table_host, int n_entries, int members,
                                  void* stream) {
  D3B_REQUIRE(src && dst_bf16 && table_host, "shadow_weights: null pointer");
  D3B_REQUIRE(n_entries >= 1 && n_entries <= MAX_SHADOW && members >= 1, "shadow_weights: 1..%d entries", MAX_SHADOW);
  ShadowTable t{};
  t.n = n_entries;
  long long biggest = 0;
  for (int i = 0; i < n_entries; ++i) {
    const int64_t* r = table_host + 7 * i;
    t.e[i] = ShadowEntry{r[0], (int)r[1], (int)r[2], r[3], r[4], r[5], r[6]};
    long long tiles = ceil_div_ll(r[1], 32) * ceil_div_ll(r[2], 32);
    if (tiles > biggest) biggest = tiles;
  }
  dim3 grid((unsigned)std::min<long long>(biggest, 4 * kNumSM), n_entries, members);
  launch_pdl(shadow_kernel, grid, dim3(256), 0, ST, src, (long long)src_member_stride, (__nv_bfloat16*)dst_bf16,
             (long long)dst_member_stride, t);
  return check_launch("shadow_weights");
}

extern "C" int d3b_to_bf16(const float* src, int64_t lds, int rows, int cols, void* dst, int64_t ldd, void* dst_t,
                           int64_t ldt, void* stream) {
  D3B_REQUIRE(rows >= 0 && cols >= 1, "to_bf16: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(src && (dst || dst_t), "to_bf16: null pointer");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  launch_pdl(to_bf16_kernel, grid, dim3(256), 0, ST, src, (long long)lds, rows, cols, (__nv_bfloat16*)dst, (long long)ldd,
             (__nv_bfloat16*)dst_t, (long long)ldt);
  return check_launch("to_bf16");
}

extern "C" int d3b_head_forward_bf16(const void* x, int64_t ldx, int64_t stride_x, const float* w, int64_t ldw,
                                     int64_t stride_w, const float* bias, int64_t stride_b, float* y, int64_t ldy,
                                     int64_t stride_y, int rows, int out_features, int in_features, int members,
                                     int act_tanh, void* stream) {
  D3B_REQUIRE(rows >= 0 && in_features > 0 && members > 0, "head_forward_bf16: bad sizes");
  D3B_REQUIRE(out_features >= 1 && out_features <= 32, "head_forward_bf16: out_features %d not in 1..32",
              out_features);
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(x && w && y, "head_forward_bf16: null pointer");
  long long warps = (long long)rows * members;
  // measured (profiles/r2/head_bf16_bench.py): the chunked kernel wins for one or two outputs (25 600 x 300 -> 1, two
  // members: 32.3 -> 19.4 us) and for few rows (256 x 750 -> 6: 8.0 -> 6.1 us); with six outputs over 25 600 rows the
  // element-wise kernel's four-outputs-at-a-time loop is faster (32 us against 60), so it keeps those
  const bool vec_ok = in_features > 256 && in_features <= 1024 && ldx % 8 == 0 && stride_x % 8 == 0 &&
                      ((uintptr_t)x & 15) == 0 && ldw % 2 == 0 && stride_w % 2 == 0 && ((uintptr_t)w & 7) == 0 &&
                      (out_features <= 2 || warps <= 2048);
  if (vec_ok) {
    dim3 grid((unsigned)ceil_div_ll(warps, 8));
    if (in_features <= 768)
      launch_pdl(head_forward_bf16_vec_kernel<3>, grid, dim3(256), 0, ST, (const __nv_bfloat16*)x, (long long)ldx,
                 (long long)stride_x, w, (long long)ldw, (long long)stride_w, bias, (long long)stride_b, y,
                 (long long)ldy, (long long)stride_y, rows, out_features, in_features, members, act_tanh);
    else
      launch_pdl(head_forward_bf16_vec_kernel<4>, grid, dim3(256), 0, ST, (const __nv_bfloat16*)x, (long long)ldx,
                 (long long)stride_x, w, (long long)ldw, (long long)stride_w, bias, (long long)stride_b, y,
                 (long long)ldy, (long long)stride_y, rows, out_features, in_features, members, act_tanh);
    return check_launch("head_forward_bf16");
  }
  if (in_features <= 256)
    head_forward_bf16_kernel<8><<<(unsigned)ceil_div_ll(warps, 8), 256, 0, ST>>>(
        (const __nv_bfloat16*)x, ldx, stride_x, w, ldw, stride_w, bias, stride_b, y, ldy, stride_y, rows, out_features,
        in_features, members, act_tanh);
  else
    head_forward_bf16_kernel<24><<<(unsigned)ceil_div_ll(warps, 8), 256, 0, ST>>>(
        (const __nv_bfloat16*)x, ldx, stride_x, w, ldw, stride_w, bias, stride_b, y, ldy, stride_y, rows, out_features,
        in_features, members, act_tanh);
  return check_launch("head_forward_bf16");
}

extern "C" int d3b_head_backward_data_bf16(const float* dy, int64_t lddy, int64_t stride_dy, const float* w,
                                           int64_t ldw, int64_t stride_w, void* dx, int64_t lddx, int64_t stride_dx,
                                           void* dx_t, int64_t ldt, int64_t stride_dxt, const void* relu_src,
                                           int64_t ld_src, int64_t stride_src, int rows, int out_features,
                                           int in_features, int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && in_features > 0 && members > 0 && out_features >= 1, "head_backward_data_bf16: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && w && dx, "head_backward_data_bf16: null pointer");
  dim3 grid(ceil_div(in_features, 32), ceil_div(rows, 32), members);
  head_backward_data_bf16_kernel<<<grid, 256, 0, ST>>>(dy, lddy, stride_dy, w, ldw, stride_w, (__nv_bfloat16*)dx, lddx,
                                                       stride_dx, (__nv_bfloat16*)dx_t, ldt, stride_dxt,
                                                       (const __nv_bfloat16*)relu_src, ld_src, stride_src, rows,
                                                       out_features, in_features);
  return check_launch("head_backward_data_bf16");
}

extern "C" int d3b_head_backward_weight_bf16(const float* dy, int64_t lddy, int64_t stride_dy, const void* x,
                                             int64_t ldx, int64_t stride_x, float* dw, int64_t lddw,
                                             int64_t stride_dw, float* dbias, int64_t stride_db, int rows,
                                             int out_features, int in_features, int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && in_features > 0 && members > 0, "head_backward_weight_bf16: bad sizes");
  D3B_REQUIRE(out_features >= 1 && out_features <= 32, "head_backward_weight_bf16: out_features %d not in 1..32",
              out_features);
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && x && dw, "head_backward_weight_bf16: null pointer");
  int kblocks = ceil_div(in_features, 256);
  int rpb = rows / (2 * kNumSM / (kblocks * members) + 1);
  if (rpb < 4) rpb = 4;
  if (rpb > 64) rpb = 64;
  dim3 grid(kblocks, ceil_div(rows, rpb), members);
  size_t smem = (size_t)rpb * out_features * sizeof(float);
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
#define HBW(NM)                                                                                                   \
  head_backward_weight_bf16_kernel<NM><<<grid, 256, smem, ST>>>(dy, lddy, stride_dy, xb, ldx, stride_x, dw, lddw, \
                                                                stride_dw, dbias, stride_db, rows, out_features,  \
                                                                in_features, rpb)
  if (out_features <= 1) HBW(1);
  else if (out_features <= 8) HBW(8);
  else if (out_features <= 16) HBW(16);
  else HBW(32);
#undef HBW
  return check_launch("head_backward_weight_bf16");
}

extern "C" int d3b_colsum_bf16(const void* dz, int64_t ld, int64_t stride_z, float* dbias, int64_t stride_db,
                               int rows, int cols, int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && cols >= 1 && members >= 1, "colsum_bf16: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dz && dbias, "colsum_bf16: null pointer");
  int nblocks = ceil_div(cols, 256);
  int rpb = rows / (kNumSM / (nblocks * members) + 1);
  if (rpb < 8) rpb = 8;
  dim3 grid(nblocks, ceil_div(rows, rpb), members);
  colsum_bf16_kernel<<<grid, 256, 0, ST>>>((const __nv_bfloat16*)dz, ld, stride_z, dbias, stride_db, rows, cols, rpb);
  return check_launch("colsum_bf16");
}
