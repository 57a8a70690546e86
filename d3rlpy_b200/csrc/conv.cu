// K9 (encoder side): Nature-DQN convolution layers as im2col + the batched dense-layer GEMMs.
//
//   PixelEncoder.forward (d3rlpy/models/torch/encoders.py:130-162): conv(k8,s4) -> conv(k4,s2) ->
//   conv(k3,s1) (+ReLU each) -> flatten (NCHW order) -> fc + ReLU.  Activations are kept NHWC
//   ([(b,oh,ow)][oc] is exactly the GEMM output), patches are laid out in the weight's own (ic,kh,kw)
//   order so nn.Conv2d weights [oc][ic][kh][kw] are used in place as the K-major B operand, and the final
//   fc is the same operator with a kernel covering the whole feature map (its (c,h,w) patch order equals
//   the reference's NCHW flatten).  PixelScaler (x/255 as a true division, d3rlpy/preprocessing/scalers.py:109-110) and the
//   uint8 -> float cast (torch_utility.py:146-149) are fused into the first layer's patch load.
//   Backward: col2im gathers dPatches back onto the NHWC input and applies the producer's ReLU mask.
#include <cuda_bf16.h>

#include "common.cuh"

namespace d3b {

template <typename In>
__device__ __forceinline__ float load_scaled(const In* p, float divisor);
template <>
__device__ __forceinline__ float load_scaled<uint8_t>(const uint8_t* p, float scale) {
  return (float)__ldg(p) / scale;
}
template <>
__device__ __forceinline__ float load_scaled<float>(const float* p, float scale) {
  return __ldg(p) / scale;
}
template <>
__device__ __forceinline__ float load_scaled<__nv_bfloat16>(const __nv_bfloat16* p, float scale) {
  return __bfloat162float(*p) / scale;
}

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// patches[e][(b,oh,ow)][ic*k*k + kh*k + kw] = x[...] / divisor ; x[e][b][ic][oh*s+kh][ow*s+kw]  (generic element strides)
// One thread per (patch row, ic, kh): it writes the ksz consecutive kw entries (32-bit index arithmetic; the host
// checks the sizes).  With unit stride along the width (first layer: NCHW uint8 frames) the run is contiguous in
// the source as well.
template <typename In, typename Out>
__global__ void __launch_bounds__(256) im2col_kernel(const In* __restrict__ x, long long sx_member, long long sb,
                                                     long long sc, long long sh, long long sw,
                                                     Out* __restrict__ out, long long ldo, long long so_member,
                                                     int images, int C, int OH, int OW, int ksz, int stride,
                                                     float scale, int K_pad) {
  pdl_trigger();
  pdl_wait();
  const int K = C * ksz * ksz;
  const int runs = C * ksz;                         // (ic, kh) runs per patch row
  const int pad_runs = (K_pad - K + ksz - 1) / ksz; // zero padding up to the leading dimension, in runs of ksz
  const int per_row = runs + pad_runs;
  const unsigned rows = (unsigned)images * OH * OW;
  const unsigned total = rows * (unsigned)per_row;
  const In* xe = x + (long long)blockIdx.y * sx_member;
  Out* oe = out + (long long)blockIdx.y * so_member;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned row = idx / per_row;
    const int run = (int)(idx - row * per_row);
    Out* o = oe + (long long)row * ldo + run * ksz;
    if (run >= runs) {
      for (int j = 0; j < ksz && run * ksz + j < K_pad; ++j) store_out(o + j, 0.f);
      continue;
    }
    const int kh = run % ksz, ic = run / ksz;
    const unsigned ow = row % OW, t2 = row / OW;
    const unsigned oh = t2 % OH, b = t2 / OH;
    const In* src = xe + (long long)b * sb + (long long)ic * sc + (long long)(oh * stride + kh) * sh +
                    (long long)(ow * stride) * sw;
    for (int j = 0; j < ksz; ++j) store_out(o + j, load_scaled<In>(src + (long long)j * sw, scale));
  }
}

// dx[e][(b,ih,iw)][ic] = [y_prev > 0] * sum_{kh,kw : (ih-kh)%s==0, (iw-kw)%s==0} dpatch[e][(b,oh,ow)][ic*k*k+kh*k+kw]
template <typename Mask, typename Out>
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dpatch, long long ldp,
                                                     long long sp_member, const Mask* __restrict__ y_prev,
                                                     long long ldy, long long sy_member, Out* __restrict__ dx,
                                                     long long lddx, long long sdx_member, int images, int C, int H,
                                                     int W, int OH, int OW, int ksz, int stride) {
  const long long total = (long long)images * H * W * C;
  const float* pe = dpatch + (long long)blockIdx.y * sp_member;
  Out* de = dx + (long long)blockIdx.y * sdx_member;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ic = (int)(idx % C);
    long long pix = idx / C;  // (b, ih, iw)
    int iw = (int)(pix % W);
    long long t = pix / W;
    int ih = (int)(t % H);
    long long b = t / H;
    float acc = 0.f;
    bool live = true;
    if (y_prev) live = (float)y_prev[(long long)blockIdx.y * sy_member + pix * ldy + ic] > 0.f;
    if (live) {
      for (int kh = ih % stride; kh < ksz; kh += stride) {
        int oh = (ih - kh) / stride;
        if (ih < kh || oh >= OH) continue;
        for (int kw = iw % stride; kw < ksz; kw += stride) {
          int ow = (iw - kw) / stride;
          if (iw < kw || ow >= OW) continue;
          acc += __ldg(pe + ((b * OH + oh) * OW + ow) * ldp + (ic * ksz + kh) * ksz + kw);
        }
      }
    }
    store_out(de + pix * lddx + ic, acc);
  }
}

}  // namespace d3b

using namespace d3b;
#define ST ((cudaStream_t)stream)

static inline int grid_for(long long total) {
  long long g = (total + 255) / 256;
  long long cap = (long long)kNumSM * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int d3b_im2col(const void* x, int x_is_u8, int64_t stride_x, int64_t sb, int64_t sc, int64_t sh,
                          int64_t sw, void* out, int out_is_bf16, int64_t ldo, int64_t stride_o, int images,
                          int channels, int height, int width, int ksize, int stride, float divisor, int members,
                          void* stream) {
  D3B_REQUIRE(images >= 0 && channels >= 1 && ksize >= 1 && stride >= 1 && members >= 1, "im2col: bad sizes");
  D3B_REQUIRE(height >= ksize && width >= ksize, "im2col: kernel larger than the input");
  if (images == 0) return D3B_OK;
  D3B_REQUIRE(x && out, "im2col: null pointer");
  int OH = (height - ksize) / stride + 1, OW = (width - ksize) / stride + 1;
  int K = channels * ksize * ksize;
  D3B_REQUIRE(ldo >= K, "im2col: ldo < C*k*k");
  int K_pad = out_is_bf16 ? (int)ldo : K;  // bf16 rows are zero-padded to the 16-byte leading dimension
  long long per_row = (long long)channels * ksize + (K_pad - K + ksize - 1) / ksize;
  long long total = (long long)images * OH * OW * per_row;
  D3B_REQUIRE(total < (1LL << 31), "im2col: too many patch runs for one launch");
  dim3 grid(grid_for(total), members);
#define LAUNCH(IN, OUT)                                                                                               \
  launch_pdl(im2col_kernel<IN, OUT>, grid, dim3(256), 0, ST, (const IN*)x, (long long)stride_x, (long long)sb,        \
             (long long)sc, (long long)sh, (long long)sw, (OUT*)out, (long long)ldo, (long long)stride_o, images,     \
             channels, OH, OW, ksize, stride, divisor, K_pad)
  // x_is_u8: 0 = fp32 input, 1 = uint8 input, 2 = bf16 input (NHWC activations of the tensor-core path)
  if (x_is_u8 == 1 && out_is_bf16) LAUNCH(uint8_t, __nv_bfloat16);
  else if (x_is_u8 == 1) LAUNCH(uint8_t, float);
  else if (x_is_u8 == 2 && out_is_bf16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (x_is_u8 == 2) LAUNCH(__nv_bfloat16, float);
  else if (out_is_bf16) LAUNCH(float, __nv_bfloat16);
  else LAUNCH(float, float);
#undef LAUNCH
  return check_launch("im2col");
}

extern "C" int d3b_col2im(const float* dpatch, int64_t ldp, int64_t stride_p, const void* y_prev, int y_is_bf16,
                          int64_t ldy, int64_t stride_y, void* dx, int64_t lddx, int64_t stride_dx, int images,
                          int channels, int height, int width, int ksize, int stride, int members, void* stream) {
  D3B_REQUIRE(images >= 0 && channels >= 1 && ksize >= 1 && stride >= 1 && members >= 1, "col2im: bad sizes");
  D3B_REQUIRE(height >= ksize && width >= ksize, "col2im: kernel larger than the input");
  if (images == 0) return D3B_OK;
  D3B_REQUIRE(dpatch && dx, "col2im: null pointer");
  int OH = (height - ksize) / stride + 1, OW = (width - ksize) / stride + 1;
  long long total = (long long)images * height * width * channels;
  dim3 grid(grid_for(total), members);
  // bf16 mask => tensor-core path: the gradient is written as bf16 too (operand of the next layer's GEMMs)
  if (y_is_bf16)
    col2im_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, ST>>>(
        dpatch, ldp, stride_p, (const __nv_bfloat16*)y_prev, ldy, stride_y, (__nv_bfloat16*)dx, lddx, stride_dx, images,
        channels, height, width, OH, OW, ksize, stride);
  else
    col2im_kernel<float, float><<<grid, 256, 0, ST>>>(dpatch, ldp, stride_p, (const float*)y_prev, ldy, stride_y,
                                                      (float*)dx, lddx, stride_dx, images, channels, height, width, OH,
                                                      OW, ksize, stride);
  return check_launch("col2im");
}
