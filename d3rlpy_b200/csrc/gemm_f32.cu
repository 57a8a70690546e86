// fp32 (SIMT FFMA) batched-ensemble GEMM family: the "fp32 mode" (1e-5) dense path.
// One launch covers all E ensemble members (blockIdx.z), replacing the reference's Python loop over
// q_funcs (d3rlpy/models/torch/q_functions/ensemble_q_function.py:95-105,144-146,168-170) and the
// nn.Linear calls of the encoders (d3rlpy/models/torch/encoders.py:265-275).
//
//   forward : Y[e] = act(X[e] W[e]^T + b[e])                       (A row-contig, B row-contig)
//   dgrad   : dX[e] = (dY[e] W[e]) * [H[e] > 0]                    (A row-contig, B col-contig)
//   wgrad   : dW[e] += dY[e]^T X[e] ; db[e] += colsum(dY[e])       (A col-contig, B col-contig, split-R + RED)
//
// C[m][n] = sum_r A(m,r) B(r,n).  A_RC: A stored [m][r] (r contiguous) else [r][m];
// B_RC: B stored [n][r] (r contiguous) else [r][n].
#include "common.cuh"

namespace d3b {

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, R;
  long long lda, ldb, ldc;
  long long sA, sB, sC;  // per-member strides (0 = shared operand)
  int E, splits, r_chunk;
  const float* bias;
  long long sBias;
  int relu;
  const float* mask;
  long long ldmask, sMask;
  float* colsum;  // wgrad: bias gradient, [E][M]
  long long sColsum;
  int atomic;
  int vecA, vecB, vecC;
};

constexpr int PAD = 4;

// Loads a BX x BK tile into S[BK][BX+PAD] (zero-filled outside bounds).
template <bool RC, int BX, int BK, int NT>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, long long ld, int x0, int X, int r0,
                                          int Rend, float (*S)[BX + PAD], bool vec, int tid) {
  if (RC) {
    constexpr int VR = BK / 4;  // vectors per x-row
    constexpr int NV = BX * VR;
#pragma unroll
    for (int v = tid; v < NV; v += NT) {
      int x = v / VR, rv = (v % VR) * 4;
      int gx = x0 + x, gr = r0 + rv;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gx < X) {
        const float* p = base + (long long)gx * ld + gr;
        if (vec && gr + 3 < Rend) {
          val = __ldg((const float4*)p);
        } else {
          if (gr + 0 < Rend) val.x = __ldg(p + 0);
          if (gr + 1 < Rend) val.y = __ldg(p + 1);
          if (gr + 2 < Rend) val.z = __ldg(p + 2);
          if (gr + 3 < Rend) val.w = __ldg(p + 3);
        }
      }
      S[rv + 0][x] = val.x;
      S[rv + 1][x] = val.y;
      S[rv + 2][x] = val.z;
      S[rv + 3][x] = val.w;
    }
  } else {
    constexpr int VX = BX / 4;  // vectors per r-row
    constexpr int NV = BK * VX;
#pragma unroll
    for (int v = tid; v < NV; v += NT) {
      int r = v / VX, xv = (v % VX) * 4;
      int gr = r0 + r, gx = x0 + xv;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < Rend) {
        const float* p = base + (long long)gr * ld + gx;
        if (vec && gx + 3 < X) {
          val = __ldg((const float4*)p);
        } else {
          if (gx + 0 < X) val.x = __ldg(p + 0);
          if (gx + 1 < X) val.y = __ldg(p + 1);
          if (gx + 2 < X) val.z = __ldg(p + 2);
          if (gx + 3 < X) val.w = __ldg(p + 3);
        }
      }
      *(float4*)&S[r][xv] = val;
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN, bool A_RC, bool B_RC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_f32_kernel(GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TX = BN / TN;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int e = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int r_begin = split * g.r_chunk;
  const int r_end = min(g.R, r_begin + g.r_chunk);
  const float* A = g.A + (long long)e * g.sA;
  const float* B = g.B + (long long)e * g.sB;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float asum[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) asum[i] = 0.f;
  const bool do_colsum = (g.colsum != nullptr) && (blockIdx.x == 0) && (tx == 0);

  int buf = 0;
  if (r_begin < r_end) {
    load_tile<A_RC, BM, BK, NT>(A, g.lda, m0, g.M, r_begin, r_end, As[0], g.vecA, tid);
    load_tile<B_RC, BN, BK, NT>(B, g.ldb, n0, g.N, r_begin, r_end, Bs[0], g.vecB, tid);
  }
  __syncthreads();
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    int nxt = r0 + BK;
    if (nxt < r_end) {
      load_tile<A_RC, BM, BK, NT>(A, g.lda, m0, g.M, nxt, r_end, As[buf ^ 1], g.vecA, tid);
      load_tile<B_RC, BN, BK, NT>(B, g.ldb, n0, g.N, nxt, r_end, Bs[buf ^ 1], g.vecB, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        if (TM % 4 == 0) {
          float4 t = *(const float4*)&As[buf][kk][ty * TM + i];
          a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
        }
      }
      if (TM % 4 != 0) {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][ty * TM + i];
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 t = *(const float4*)&Bs[buf][kk][tx * TN + j];
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_colsum) {
#pragma unroll
        for (int i = 0; i < TM; ++i) asum[i] += a[i];
      }
    }
    __syncthreads();
    buf ^= 1;
  }

  float* C = g.C + (long long)e * g.sC;
  const float* bias = g.bias ? g.bias + (long long)e * g.sBias : nullptr;
  const float* mask = g.mask ? g.mask + (long long)e * g.sMask : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
    int n = n0 + tx * TN;
    float v[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float x = acc[i][j];
      int nn = n + j;
      if (nn < g.N) {
        if (bias) x += __ldg(bias + nn);
        if (g.relu) x = fmaxf(x, 0.f);
        if (mask) x = (__ldg(mask + (long long)m * g.ldmask + nn) > 0.f) ? x : 0.f;
      }
      v[j] = x;
    }
    float* dst = C + (long long)m * g.ldc + n;
    if (g.atomic) {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (n + j < g.N) atomicAdd(dst + j, v[j]);
    } else if (g.vecC && TN == 4 && n + 3 < g.N) {
      *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (n + j < g.N) dst[j] = v[j];
    }
  }
  if (do_colsum) {
    float* cs = g.colsum + (long long)e * g.sColsum;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int m = m0 + ty * TM + i;
      if (m < g.M) atomicAdd(cs + m, asum[i]);
    }
  }
}

template <bool A_RC, bool B_RC>
static int launch(GemmArgs& g, cudaStream_t st) {
  // big tile when there is enough M to fill the machine, else the small-M tile
  long long big_ctas = (long long)ceil_div(g.M, 128) * ceil_div(g.N, 64) * g.E * g.splits;
  if (big_ctas >= 2 * kNumSM) {
    dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 128), g.E * g.splits);
    gemm_f32_kernel<128, 64, 16, 8, 4, A_RC, B_RC><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 32), g.E * g.splits);
    gemm_f32_kernel<32, 64, 16, 2, 4, A_RC, B_RC><<<grid, 256, 0, st>>>(g);
  }
  return check_launch("gemm_f32");
}

static bool aligned16(const void* p) { return ((uintptr_t)p % 16) == 0; }

// 3xTF32 tensor-core engine (gemm_tf32x3.cu): same operands, same results to fp32 rounding
namespace tc32 {
int engine();
int ops();
int wgrad_split_mode();
int gemm(const float* a, long long lda, long long stride_a, int a_rc, const float* b, long long ldb,
         long long stride_b, int b_rc, float* c, long long ldc, long long stride_c, int m, int n, int r, int members,
         int splits, const float* bias, long long stride_bias, int relu, const float* mask, long long ld_mask,
         long long stride_mask, float* colsum, long long stride_colsum, int atomic, cudaStream_t stream);
}  // namespace tc32

// the tensor-core tile is 128 rows: below half a tile the SIMT small-M kernel wastes less
static bool use_tc(int m_rows, int op_bit, long long weight_elems = 0) {
  if (tc32::engine() != 1 || !(tc32::ops() & op_bit)) return false;
  // below half a row tile the SIMT small-M kernel wastes less — unless the layer is big: the Nature-DQN fc layer
  // (3136 -> 512) at batch 32 ran 272 us on 8 SIMT CTAs; the cluster split-K tensor-core configuration spreads the
  // reduction over the machine
  return m_rows >= 64 || (m_rows >= 16 && weight_elems >= (1LL << 20));
}

}  // namespace d3b

using namespace d3b;

extern "C" int d3b_linear_forward(const float* x, int64_t ldx, int64_t stride_x, const float* w, int64_t ldw,
                                  int64_t stride_w, const float* bias, int64_t stride_b, float* y, int64_t ldy,
                                  int64_t stride_y, int rows, int out_features, int in_features, int members,
                                  int relu, void* stream) {
  D3B_REQUIRE(rows >= 0 && out_features > 0 && in_features > 0 && members > 0, "linear_forward: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(x && w && y, "linear_forward: null pointer");
  D3B_REQUIRE(ldx >= in_features && ldw >= in_features && ldy >= out_features, "linear_forward: bad leading dims");
  if (use_tc(rows, 1, (long long)out_features * in_features))
    return tc32::gemm(x, ldx, stride_x, 1, w, ldw, stride_w, 1, y, ldy, stride_y, rows, out_features, in_features,
                      members, 1, bias, stride_b, relu, nullptr, 0, 0, nullptr, 0, 0, (cudaStream_t)stream);
  GemmArgs g{};
  g.A = x; g.B = w; g.C = y;
  g.M = rows; g.N = out_features; g.R = in_features;
  g.lda = ldx; g.ldb = ldw; g.ldc = ldy;
  g.sA = stride_x; g.sB = stride_w; g.sC = stride_y;
  g.E = members; g.splits = 1; g.r_chunk = in_features;
  g.bias = bias; g.sBias = stride_b; g.relu = relu;
  g.vecA = aligned16(x) && ldx % 4 == 0 && stride_x % 4 == 0;
  g.vecB = aligned16(w) && ldw % 4 == 0 && stride_w % 4 == 0;
  g.vecC = aligned16(y) && ldy % 4 == 0 && stride_y % 4 == 0;
  return launch<true, true>(g, (cudaStream_t)stream);
}

extern "C" int d3b_linear_backward_data(const float* dy, int64_t lddy, int64_t stride_dy, const float* w,
                                        int64_t ldw, int64_t stride_w, float* dx, int64_t lddx, int64_t stride_dx,
                                        const float* relu_src, int64_t ld_src, int64_t stride_src, int rows,
                                        int out_features, int in_features, int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && out_features > 0 && in_features > 0 && members > 0, "linear_backward_data: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && w && dx, "linear_backward_data: null pointer");
  if (use_tc(rows, 2, (long long)out_features * in_features))
    return tc32::gemm(dy, lddy, stride_dy, 1, w, ldw, stride_w, 0, dx, lddx, stride_dx, rows, in_features,
                      out_features, members, 1, nullptr, 0, 0, relu_src, ld_src, stride_src, nullptr, 0, 0,
                      (cudaStream_t)stream);
  GemmArgs g{};
  g.A = dy; g.B = w; g.C = dx;
  g.M = rows; g.N = in_features; g.R = out_features;
  g.lda = lddy; g.ldb = ldw; g.ldc = lddx;
  g.sA = stride_dy; g.sB = stride_w; g.sC = stride_dx;
  g.E = members; g.splits = 1; g.r_chunk = out_features;
  g.mask = relu_src; g.ldmask = ld_src; g.sMask = stride_src;
  g.vecA = aligned16(dy) && lddy % 4 == 0 && stride_dy % 4 == 0;
  g.vecB = aligned16(w) && ldw % 4 == 0 && stride_w % 4 == 0;
  g.vecC = aligned16(dx) && lddx % 4 == 0 && stride_dx % 4 == 0;
  return launch<true, false>(g, (cudaStream_t)stream);
}

extern "C" int d3b_linear_backward_weight(const float* dy, int64_t lddy, int64_t stride_dy, const float* x,
                                          int64_t ldx, int64_t stride_x, float* dw, int64_t lddw, int64_t stride_dw,
                                          float* dbias, int64_t stride_db, int rows, int out_features,
                                          int in_features, int members, void* stream) {
  D3B_REQUIRE(rows >= 0 && out_features > 0 && in_features > 0 && members > 0, "linear_backward_weight: bad sizes");
  if (rows == 0) return D3B_OK;
  D3B_REQUIRE(dy && x && dw, "linear_backward_weight: null pointer");
  if (use_tc(out_features, 4) && rows >= 32) {
    // split the minibatch-row reduction so that two CTAs per SM exist (the throughput configuration of the kernel:
    // measured 27.9 -> 19.2 us at the c2 shape and 1105 -> 683 us at the c5 shape against one CTA per SM; four per SM
    // is slower again, profiles/r2/tc32_wgrad.py)
    int bn = in_features > 64 ? 128 : (in_features > 32 ? 64 : 32);
    long long tiles = (long long)ceil_div(out_features, 128) * ceil_div(in_features, bn) * members;
    int want = (int)((2 * kNumSM) / tiles);
    const int mode = tc32::wgrad_split_mode();               // profiling knob
    if (mode == 1) want = (int)ceil_div_ll(kNumSM, tiles);
    if (mode == 2) want = (int)(kNumSM / tiles);
    if (mode == 3) want = (int)((4 * kNumSM) / tiles);
    int max_splits = ceil_div(rows, 64);
    int splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
    return tc32::gemm(dy, lddy, stride_dy, 0, x, ldx, stride_x, 0, dw, lddw, stride_dw, out_features, in_features,
                      rows, members, splits, nullptr, 0, 0, nullptr, 0, 0, dbias, stride_db, 1, (cudaStream_t)stream);
  }
  GemmArgs g{};
  g.A = dy; g.B = x; g.C = dw;
  g.M = out_features; g.N = in_features; g.R = rows;
  g.lda = lddy; g.ldb = ldx; g.ldc = lddw;
  g.sA = stride_dy; g.sB = stride_x; g.sC = stride_dw;
  g.E = members;
  // split the batch-row reduction so that the grid fills ~2 waves of the 148 SMs
  long long tiles = (long long)ceil_div(g.M, 32) * ceil_div(g.N, 64) * members;
  int want = (int)ceil_div_ll(2LL * kNumSM, tiles);
  int max_splits = ceil_div(rows, 64);
  int splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  int chunk = ceil_div(ceil_div(rows, splits), 16) * 16;
  g.splits = ceil_div(rows, chunk);
  g.r_chunk = chunk;
  g.colsum = dbias; g.sColsum = stride_db;
  g.atomic = 1;  // dW/db accumulate into the (pre-zeroed) gradient arena
  g.vecA = aligned16(dy) && lddy % 4 == 0 && stride_dy % 4 == 0;
  g.vecB = aligned16(x) && ldx % 4 == 0 && stride_x % 4 == 0;
  g.vecC = 0;
  return launch<false, false>(g, (cudaStream_t)stream);
}
