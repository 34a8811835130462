// SAM ViT image encoder assembly (modeling_sam.py:1058-1072): preprocessing, patch-embed im2col,
// LayerNorm (+ window partition), neck; the contractions run on the tcgen05 GEMM / attention kernels.
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ptx.cuh"

namespace ysi {

// ------------------------------------------------------------------------------------------------
// a1, resize step: torchvision's uint8 bilinear resize with antialias=True (image_processing_sam.py:205-250 ->
// image_processing_backends.py resize) is Pillow's fixed-point separable resampler: per output index a
// window [xmin, xmin+xsize) of the input, triangle-filter weights normalised in double and quantised to
// int16 with `prec` fractional bits (as many as keep the largest weight below 2^15), accumulation in int32,
// (acc + 2^(prec-1)) >> prec clamped to uint8; horizontal pass first, then vertical, each rounding to uint8.
// build_resize_tables restates aten/src/ATen/native/cpu/UpSampleKernel.cpp (_compute_indices_int16_weights_aa).
// ------------------------------------------------------------------------------------------------
ResizeTables build_resize_tables(int in_size, int out_size) {
  ResizeTables t;
  t.in_size = in_size; t.out_size = out_size;
  const double scale = static_cast<double>(in_size) / out_size;
  const double support = scale >= 1.0 ? scale : 1.0;                 // interp_size / 2 = 1 for the triangle filter
  const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
  t.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  t.xmin.assign(out_size, 0); t.xsize.assign(out_size, 0);
  std::vector<double> w(static_cast<size_t>(out_size) * t.ksize, 0.0);
  double wmax = 0.0;
  for (int i = 0; i < out_size; ++i) {
    const double center = scale * (i + 0.5);
    int lo = static_cast<int>(center - support + 0.5); if (lo < 0) lo = 0;
    int hi = static_cast<int>(center + support + 0.5); if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    t.xmin[i] = lo; t.xsize[i] = n;
    double total = 0.0;
    for (int j = 0; j < n; ++j) {
      const double x = (j + lo - center + 0.5) * invscale;
      const double v = x < 0 ? 1.0 + x : 1.0 - x;
      w[static_cast<size_t>(i) * t.ksize + j] = v > 0.0 ? v : 0.0;
      total += w[static_cast<size_t>(i) * t.ksize + j];
    }
    for (int j = 0; j < n; ++j) {
      if (total != 0.0) w[static_cast<size_t>(i) * t.ksize + j] /= total;
      if (w[static_cast<size_t>(i) * t.ksize + j] > wmax) wmax = w[static_cast<size_t>(i) * t.ksize + j];
    }
  }
  int prec = 0;
  for (prec = 0; prec < 22; ++prec) {
    const int next = static_cast<int>(0.5 + wmax * (1 << (prec + 1)));
    if (next >= (1 << 15)) break;
  }
  t.prec = prec;
  t.weights.assign(w.size(), 0);
  for (size_t k = 0; k < w.size(); ++k) t.weights[k] = static_cast<int16_t>(w[k] * (1 << prec) + 0.5);   // weights are >= 0
  return t;
}

// one thread = one output pixel (3 channels); horizontal: dst[b,y,i,:] from src[b,y,xmin[i]..,:]
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, int n, int H, int row_stride, size_t img_stride,
                                const int* __restrict__ xmin, const int* __restrict__ xsize, const int16_t* __restrict__ wts,
                                int ksize, int prec, int ow, uint8_t* __restrict__ dst) {
  const long long total = static_cast<long long>(n) * H * ow;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % ow);
    const long long t = idx / ow;
    const int y = static_cast<int>(t % H), b = static_cast<int>(t / H);
    const uint8_t* sp = src + b * img_stride + static_cast<size_t>(y) * row_stride + static_cast<size_t>(xmin[i]) * 3;
    const int16_t* w = wts + static_cast<size_t>(i) * ksize;
    const int cnt = xsize[i];
    int a0 = 1 << (prec - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int wj = w[j];
      a0 += sp[3 * j] * wj; a1 += sp[3 * j + 1] * wj; a2 += sp[3 * j + 2] * wj;
    }
    uint8_t* dp = dst + idx * 3;
    dp[0] = static_cast<uint8_t>(min(max(a0 >> prec, 0), 255));
    dp[1] = static_cast<uint8_t>(min(max(a1 >> prec, 0), 255));
    dp[2] = static_cast<uint8_t>(min(max(a2 >> prec, 0), 255));
  }
}

// vertical: dst[b,i,x,:] from src[b,ymin[i]..,x,:]   (src pitch = row_stride bytes)
__global__ void resize_v_kernel(const uint8_t* __restrict__ src, int n, int row_stride, size_t img_stride, int W,
                                const int* __restrict__ ymin, const int* __restrict__ ysize, const int16_t* __restrict__ wts,
                                int ksize, int prec, int oh, uint8_t* __restrict__ dst) {
  const long long total = static_cast<long long>(n) * oh * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const long long t = idx / W;
    const int i = static_cast<int>(t % oh), b = static_cast<int>(t / oh);
    const uint8_t* sp = src + b * img_stride + static_cast<size_t>(ymin[i]) * row_stride + static_cast<size_t>(x) * 3;
    const int16_t* w = wts + static_cast<size_t>(i) * ksize;
    const int cnt = ysize[i];
    int a0 = 1 << (prec - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int wj = w[j];
      const uint8_t* q = sp + static_cast<size_t>(j) * row_stride;
      a0 += q[0] * wj; a1 += q[1] * wj; a2 += q[2] * wj;
    }
    uint8_t* dp = dst + idx * 3;
    dp[0] = static_cast<uint8_t>(min(max(a0 >> prec, 0), 255));
    dp[1] = static_cast<uint8_t>(min(max(a1 >> prec, 0), 255));
    dp[2] = static_cast<uint8_t>(min(max(a2 >> prec, 0), 255));
  }
}

void launch_resize_h(const uint8_t* src, int n, int H, int row_stride, size_t img_stride, const ResizeTablesDev& t, uint8_t* dst,
                     cudaStream_t s) {
  const long long total = static_cast<long long>(n) * H * t.out_size;
  resize_h_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32)), 256, 0, s>>>(
      src, n, H, row_stride, img_stride, t.xmin, t.xsize, t.weights, t.ksize, t.prec, t.out_size, dst);
  YSI_CUDA(cudaGetLastError());
}
void launch_resize_v(const uint8_t* src, int n, int row_stride, size_t img_stride, int W, const ResizeTablesDev& t, uint8_t* dst,
                     cudaStream_t s) {
  const long long total = static_cast<long long>(n) * t.out_size * W;
  resize_v_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32)), 256, 0, s>>>(
      src, n, row_stride, img_stride, W, t.xmin, t.xsize, t.weights, t.ksize, t.prec, t.out_size, dst);
  YSI_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// a1, normalise + pad + patchify: (x - 255*mean) / (255*std) in fp32, exactly as tvF.normalize
// (image_processing_backends.py rescale_and_normalize: mean,std pre-multiplied by 1/rescale_factor); pixels
// outside the resized image [src_h, src_w] are the zero padding applied AFTER normalisation (exactly 0.0).
// One thread = 8 horizontally adjacent pixels of one channel of one patch row.
// A row (token) = patch (py,px); column = c*256 + ky*16 + kx   (Conv2d weight [D,3,16,16] flattened)
// ------------------------------------------------------------------------------------------------
// The patch-embed A operand carries every pixel as a two-term split  v = hi + lo  (both op16, lo = the rounding
// residual of hi): columns [0,768) hold hi, [768,1536) hold lo, and the weight matrix is [W | W], so the tensor
// core sums hi*W + lo*W in fp32 and the input rounding of this one GEMM -- the largest single error term of the
// encoder, since its output IS the residual stream -- drops from 2^-9 / 2^-12 to fp32 level for 0.1 ms per step.
__device__ __forceinline__ void store_patch_hi_lo(op16* dst, const float (&v)[8]) {
  float lo[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) lo[k] = v[k] - op2f(f2op(v[k]));
  uint4 o;
  o.x = pack_op16x2(v[0], v[1]); o.y = pack_op16x2(v[2], v[3]);
  o.z = pack_op16x2(v[4], v[5]); o.w = pack_op16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = o;
  o.x = pack_op16x2(lo[0], lo[1]); o.y = pack_op16x2(lo[2], lo[3]);
  o.z = pack_op16x2(lo[4], lo[5]); o.w = pack_op16x2(lo[6], lo[7]);
  *reinterpret_cast<uint4*>(dst + 768) = o;
}

__global__ void preprocess_kernel(const uint8_t* __restrict__ rgb, int n, int src_h, int src_w, int row_stride, size_t img_stride,
                                  float m0, float m1, float m2, float s0, float s1, float s2,
                                  float* __restrict__ pix, op16* __restrict__ a_patch) {
  const long long total = static_cast<long long>(n) * 3 * 1024 * 128;   // 8-pixel groups
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xg = static_cast<int>(i & 127);
    const int y = static_cast<int>((i >> 7) & 1023);
    const int c = static_cast<int>((i >> 17) % 3);
    const int b = static_cast<int>(i / (3ll << 17));
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const uint8_t* src = rgb + b * img_stride + static_cast<size_t>(y) * row_stride + static_cast<size_t>(xg) * 8 * 3 + c;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = (y < src_h && xg * 8 + k < src_w) ? __fdiv_rn(__fsub_rn(static_cast<float>(src[3 * k]), mean), sd) : 0.0f;
    if (pix) {
      float4* d = reinterpret_cast<float4*>(pix + ((static_cast<size_t>(b) * 3 + c) * 1024 + y) * 1024 + xg * 8);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (a_patch) {
      const int py = y >> 4, ky = y & 15, px = xg >> 1, kx0 = (xg & 1) * 8;
      store_patch_hi_lo(a_patch + (static_cast<size_t>(b) * 4096 + py * 64 + px) * PATCH_K + c * 256 + ky * 16 + kx0, v);
    }
  }
}

void launch_preprocess(const uint8_t* rgb, int n, int src_h, int src_w, int row_stride, size_t img_stride, const float* mean255,
                       const float* std255, float* pixel_values, op16* a_patch, cudaStream_t s) {
  const long long total = static_cast<long long>(n) * 3 * 1024 * 128;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  preprocess_kernel<<<grid, 256, 0, s>>>(rgb, n, src_h, src_w, row_stride, img_stride, mean255[0], mean255[1], mean255[2],
                                         std255[0], std255[1], std255[2], pixel_values, a_patch);
  YSI_CUDA(cudaGetLastError());
}

__global__ void im2col_patch_f32_kernel(const float* __restrict__ pix, int n, op16* __restrict__ a_patch) {
  const long long total = static_cast<long long>(n) * 3 * 1024 * 128;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xg = static_cast<int>(i & 127);
    const int y = static_cast<int>((i >> 7) & 1023);
    const int c = static_cast<int>((i >> 17) % 3);
    const int b = static_cast<int>(i / (3ll << 17));
    const float4* sp = reinterpret_cast<const float4*>(pix + ((static_cast<size_t>(b) * 3 + c) * 1024 + y) * 1024 + xg * 8);
    const float4 a = sp[0], d = sp[1];
    const int py = y >> 4, ky = y & 15, px = xg >> 1, kx0 = (xg & 1) * 8;
    const float v[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
    store_patch_hi_lo(a_patch + (static_cast<size_t>(b) * 4096 + py * 64 + px) * PATCH_K + c * 256 + ky * 16 + kx0, v);
  }
}

void launch_im2col_patch_f32(const float* pixel_values, int n, op16* a_patch, cudaStream_t s) {
  const long long total = static_cast<long long>(n) * 3 * 1024 * 128;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  im2col_patch_f32_kernel<<<grid, 256, 0, s>>>(pixel_values, n, a_patch);
  YSI_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim of fp32 rows, one warp per OUTPUT row.
//   WINDOWED: output rows are in window-partition order (25 windows x 196 per image, 64->70 zero pad,
//             modeling_sam.py:900-922); pad rows are written as zeros (they become q=k=v=bias).
// Output op16 (GEMM operand) and/or fp32.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int window_row_to_token(int wrow) {   // row within one image's 4900 -> token or -1
  const int w = wrow / 196, l = wrow - w * 196;
  const int wy = w / 5, wx = w - wy * 5, ly = l / 14, lx = l - ly * 14;
  const int y = wy * 14 + ly, x = wx * 14 + lx;
  return (y < 64 && x < 64) ? y * 64 + x : -1;
}

// SPLIT: the op16 output row is [hi(D) | lo(D)] (two-term split, see store_patch_hi_lo) for the neck's convolutions.
template <bool WINDOWED, bool SPLIT = false>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int rows_out, int D, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, op16* __restrict__ out_bf, float* __restrict__ out_f, int reverse) {
  // reverse: rows are processed last-to-first when the producing GEMM wrote (reduce-added) x in ascending tile order --
  // the rows most likely still in the L2 are then the last ones -- and first-to-last after a GEMM that ran descending
  // (one row per warp, rows / 8 CTAs. A grid-stride loop over 2 / 3 / 4 resident CTAs per SM was measured in round 2:
  // 0.95 / 0.81 / 0.97 ms per batch against 0.78 -- the hardware's CTA dispatch balances this bandwidth-bound kernel better.)
  const int warp_lin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_lin >= rows_out) return;
  const int warp_global = reverse ? rows_out - 1 - warp_lin : warp_lin;
  int src_row = warp_global;
  if (WINDOWED) {
    const int img = warp_global / 4900;
    const int tok = window_row_to_token(warp_global - img * 4900);
    if (tok < 0) {
      if (out_bf) {
        uint4* o = reinterpret_cast<uint4*>(out_bf + static_cast<size_t>(warp_global) * D);
        for (int i = lane; i < D / 8; i += 32) o[i] = make_uint4(0, 0, 0, 0);
      }
      return;
    }
    src_row = img * 4096 + tok;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(src_row) * D);
  constexpr int MAXV = 10;               // D <= 1280
  float4 v[MAXV];
  const int nv = D / 4;                  // float4 per row
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) { v[k] = xr[i]; sum += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
  const float mean = sum / D;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
  const float rstd = rsqrtf(sq / D + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      const float4 g = __ldg(g4 + i), bb = __ldg(b4 + i);
      float4 y;
      y.x = (v[k].x - mean) * rstd * g.x + bb.x;
      y.y = (v[k].y - mean) * rstd * g.y + bb.y;
      y.z = (v[k].z - mean) * rstd * g.z + bb.z;
      y.w = (v[k].w - mean) * rstd * g.w + bb.w;
      if (out_f) reinterpret_cast<float4*>(out_f + static_cast<size_t>(warp_global) * D)[i] = y;
      if (out_bf) {
        uint2 o;
        o.x = pack_op16x2(y.x, y.y);
        o.y = pack_op16x2(y.z, y.w);
        if (SPLIT) {
          op16* orow = out_bf + static_cast<size_t>(warp_global) * 2 * D;
          reinterpret_cast<uint2*>(orow)[i] = o;
          o.x = pack_op16x2(y.x - op2f(f2op(y.x)), y.y - op2f(f2op(y.y)));
          o.y = pack_op16x2(y.z - op2f(f2op(y.z)), y.w - op2f(f2op(y.w)));
          reinterpret_cast<uint2*>(orow + D)[i] = o;
        } else {
          reinterpret_cast<uint2*>(out_bf + static_cast<size_t>(warp_global) * D)[i] = o;
        }
      }
    }
  }
}

void launch_layernorm(const float* x, int rows_out, int D, const float* gamma, const float* beta, float eps,
                      op16* out_bf, float* out_f, bool windowed, cudaStream_t s, bool split, bool reverse) {
  YSI_CHECK(D % 8 == 0 && D <= 1280, "LayerNorm width must be a multiple of 8 and <= 1280");
  const int blocks = ceil_div(rows_out, 8);
  if (split)
    layernorm_kernel<false, true><<<blocks, 256, 0, s>>>(x, rows_out, D, gamma, beta, eps, out_bf, out_f, reverse ? 1 : 0);
  else if (windowed)
    layernorm_kernel<true><<<blocks, 256, 0, s>>>(x, rows_out, D, gamma, beta, eps, out_bf, out_f, reverse ? 1 : 0);
  else
    layernorm_kernel<false><<<blocks, 256, 0, s>>>(x, rows_out, D, gamma, beta, eps, out_bf, out_f, reverse ? 1 : 0);
  YSI_CUDA(cudaGetLastError());
}

__global__ void build_win_row_map_kernel(int* map, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = i / 4900;
  const int tok = window_row_to_token(i - img * 4900);
  map[i] = tok < 0 ? -1 : img * 4096 + tok;
}

void launch_build_win_row_map(int* map, int n_images, cudaStream_t s) {
  const int total = n_images * 4900;
  build_win_row_map_kernel<<<ceil_div(total, 256), 256, 0, s>>>(map, total);
  YSI_CUDA(cudaGetLastError());
}

__global__ void build_tok_win_map_kernel(int* map, int total) {     // token row -> row in window-partition order
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = i >> 12, tok = i & 4095, y = tok >> 6, x = tok & 63;
  map[i] = img * 4900 + ((y / 14) * 5 + x / 14) * 196 + (y % 14) * 14 + x % 14;
}

void launch_build_tok_win_map(int* map, int n_images, cudaStream_t s) {
  const int total = n_images * 4096;
  build_tok_win_map_kernel<<<ceil_div(total, 256), 256, 0, s>>>(map, total);
  YSI_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// LayerNorm folded into the GEMMs around it (round 2).  LN(x) W^T + b, with LN(x) = (x - mean) * rstd * gamma + beta, is
//     rstd * [ (gamma . x) W^T - mean * cs ] + wb + b,      cs[n] = sum_k gamma[k] W[n,k],  wb[n] = sum_k beta[k] W[n,k]
// so the qkv / fc1 GEMM can run on  x16 = op16(gamma . x)  of the RAW residual stream and apply mean / rstd per row in its
// epilogue (EpiStaged::prefetch / run).  x16 and the row statistics come out of the epilogue that produced x -- the proj / fc2
// residual adds and, for the first layer, the patch-embed GEMM (EpiResidLN: x_new is in registers there anyway) -- so the
// LayerNorm passes over the fp32 stream (read 100 MB + write 50 MB each at 8 ViT-B images) disappear.
// Rounding: the operand is rounded at gamma * x instead of at LN(x): the same relative step, on a value that still carries
// the row mean, which for these residual streams is small against the row's standard deviation.
// Row statistics: every producer warp writes (mean, sum of squared deviations) of its 96 / 128 columns of the row; the consumer
// merges the at most LN_STAT_SLOTS partials pairwise-stably (Chan et al.) in a fixed order -- no E[x^2] - mean^2 cancellation.
// ------------------------------------------------------------------------------------------------
// 3x3 / pad 1 im2col over the 64x64 token grid, tap-major columns; a row of `in` is the 256 channels as a two-term
// split [hi(256) | lo(256)] (NECK_C2 = 512 values):
//   A[t, (ky*3+kx)*512 + c] = in[(y+ky-1, x+kx-1), c]  (zero outside the grid)
__global__ void im2col_3x3_kernel(const op16* __restrict__ in, int n, op16* __restrict__ out) {
  const long long total = static_cast<long long>(n) * 4096 * 9 * 64;   // uint4 (8 value) groups
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i & 63);
    const int tap = static_cast<int>((i >> 6) % 9);
    const long long t = i / (9 * 64);
    const int tok = static_cast<int>(t & 4095);
    const int b = static_cast<int>(t >> 12);
    const int y = (tok >> 6) + tap / 3 - 1, x = (tok & 63) + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < 64 && x >= 0 && x < 64)
      v = *reinterpret_cast<const uint4*>(in + (static_cast<size_t>(b) * 4096 + y * 64 + x) * NECK_C2 + cg * 8);
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * NECK_K2 + tap * NECK_C2 + cg * 8) = v;
  }
}

// fp32 [rows, D] -> op16 [rows, 2D] = [hi | lo] (two-term split, see store_patch_hi_lo)
__global__ void cast_split_op16_kernel(const float* __restrict__ in, op16* __restrict__ out, long long n8, int d8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    const long long row = i / d8, c8 = i - row * d8;
    uint4 o;
    o.x = pack_op16x2(a.x, a.y); o.y = pack_op16x2(a.z, a.w);
    o.z = pack_op16x2(b.x, b.y); o.w = pack_op16x2(b.z, b.w);
    reinterpret_cast<uint4*>(out)[row * 2 * d8 + c8] = o;
    o.x = pack_op16x2(a.x - op2f(f2op(a.x)), a.y - op2f(f2op(a.y))); o.y = pack_op16x2(a.z - op2f(f2op(a.z)), a.w - op2f(f2op(a.w)));
    o.z = pack_op16x2(b.x - op2f(f2op(b.x)), b.y - op2f(f2op(b.y))); o.w = pack_op16x2(b.z - op2f(f2op(b.z)), b.w - op2f(f2op(b.w)));
    reinterpret_cast<uint4*>(out)[row * 2 * d8 + d8 + c8] = o;
  }
}

// ------------------------------------------------------------------------------------------------
void encoder_forward(const EncoderW& w, const EncoderWork& work, int n, float* emb_out, float* hidden_dump,
                     cudaStream_t s, int64_t* launches, Profiler* prof) {
  YSI_CHECK(n >= 1 && n <= work.cap, "encoder batch exceeds the workspace");
  const int D = w.D, T = n * 4096, TW = n * 4900;
  int64_t nl = 0;
  // LayerNorm folded into the neighbouring GEMMs (see the comment on the folded LayerNorm above). YSI_LN_FUSED is a bit mask: 1 = LayerNorm1 (fc2 of the
  // previous layer -> qkv), 2 = LayerNorm2 (proj -> fc1), 0 = the separate LayerNorm kernels.
  static const int ln_fused_env = [] { const char* e = getenv("YSI_LN_FUSED"); return e ? atoi(e) : 1; }();
  static const bool pair_env = [] { const char* a = getenv("YSI_GEMM_PAIR"); const char* b = getenv("YSI_GEMM_STAGED");
                                    return (!a || atoi(a) != 0) && (!b || atoi(b) != 0); }();
  const int slots = gemm_ln_stat_slots(T, D);
  const bool ln_ok = pair_env && T >= 2048 && D % 256 == 0 && w.mlp % 256 == 0 && slots <= LN_STAT_SLOTS && w.residual_mode == 2;
  const bool ln1_fused = ln_ok && (ln_fused_env & 1), ln2_fused = ln_ok && (ln_fused_env & 2);
  // patch embed: x = A * Wp^T + b + pos_embed   (modeling_sam.py:128, 1065-1066)
  {
    GemmEpilogue ep;
    ep.bias = w.b_patch; ep.add_src = w.pos_embed; ep.add_mod = 4096; ep.ld_add = D;
    ep.out_f32 = work.x; ep.ld_out = D;
    if (ln1_fused) {      // the first layer's LayerNorm1 operand and statistics come out of this epilogue as well
      const bool g0 = w.layers[0].is_global != 0;
      ep.stats_out = work.ln_stats; ep.ld_x16 = D; ep.x16_gamma = w.layers[0].ln1_g;
      if (g0) { ep.x16_out = work.h; } else { ep.x16_out = work.h_win; ep.x16_rowmap = work.tok_win_map; }
    }
    ProfScope ps(prof, KC_GEMM_PATCH, 2.0 * T * D * 768);      // algorithmic FLOPs: the lo term is overhead
    gemm_op16(work.a_patch, PATCH_K, w.w_patch, PATCH_K, T, D, PATCH_K, ep, s); ++nl;
  }
  if (hidden_dump) YSI_CUDA(cudaMemcpyAsync(hidden_dump, work.x, sizeof(float) * T * D, cudaMemcpyDeviceToDevice, s));
  for (int li = 0; li < w.L; ++li) {
    const EncoderLayerW& lw = w.layers[li];
    const bool glob = lw.is_global != 0;
    const int rows = glob ? T : TW;
    // algorithmic FLOPs (pad rows / pad keys excluded) are attached to every record for the roofline
    if (!ln1_fused) {
      ProfScope ps(prof, KC_LAYERNORM); launch_layernorm(work.x, rows, D, lw.ln1_g, lw.ln1_b, 1e-6f, work.h, nullptr, !glob, s, false, /*reverse=*/li == 0); ++nl;
    }
    {
      GemmEpilogue ep;
      ep.bias = lw.b_qkv; ep.out_op16 = work.qkv; ep.ld_out_op16 = 3 * D;
      ep.col_scale = attn_k_scale(w.head_dim); ep.scale_c0 = D; ep.scale_c1 = 2 * D;     // K in log2 units for the attention kernel
      if (ln1_fused) {
        ep.ln_stats = work.ln_stats; ep.ln_np = slots; ep.ln_dim = D; ep.ln_rowmap = glob ? nullptr : work.win_row_map;
        ep.ln_cs = lw.cs_qkv; ep.ln_wb = lw.wb_qkv; ep.ln_bw = lw.bw_qkv;
      }
      ProfScope ps(prof, KC_GEMM_QKV, 2.0 * T * 3 * D * D);
      gemm_op16((ln1_fused && !glob) ? work.h_win : work.h, D, lw.w_qkv, D, rows, 3 * D, D, ep, s); ++nl;
    }
    {
      const double S = glob ? 64.0 : 14.0, tok = glob ? 4096.0 : 196.0, nseq = glob ? n : n * 25.0;
      // QK^T + PV (4 * T^2 * hd per head) + rel-pos terms (2 * 2 * T * S * hd per head)
      ProfScope ps(prof, glob ? KC_ATTN_GLOBAL : KC_ATTN_WINDOW, nseq * w.heads * (4.0 * tok * tok * w.head_dim + 4.0 * tok * S * w.head_dim));
      launch_encoder_attention(work.qkv, lw.rel_tab, work.attn, glob ? n : n * 25, glob ? 4096 : 196, w.heads, w.head_dim, glob, !glob, s); ++nl;
    }
    {
      GemmEpilogue ep;   // x += attn * Wproj^T + b ; the attention kernel already un-partitioned the windows
      ep.bias = lw.b_proj; ep.out_f32 = work.x; ep.ld_out = D; ep.accumulate = w.residual_mode;
      if (ln2_fused) {      // LayerNorm2's operand copy (token order) and statistics come out of this epilogue
        ep.stats_out = work.ln_stats; ep.x16_out = work.h; ep.ld_x16 = D; ep.x16_gamma = lw.ln2_g;
      }
      ProfScope ps(prof, KC_GEMM_PROJ, 2.0 * T * D * D);
      gemm_op16(work.attn, D, lw.w_proj, D, T, D, D, ep, s); ++nl;
    }
    if (!ln2_fused) { ProfScope ps(prof, KC_LAYERNORM); launch_layernorm(work.x, T, D, lw.ln2_g, lw.ln2_b, 1e-6f, work.h, nullptr, false, s, false, /*reverse=*/true); ++nl; }
    {
      GemmEpilogue ep;
      ep.bias = lw.b_fc1; ep.act = ACT_GELU; ep.out_op16 = work.u; ep.ld_out_op16 = w.mlp;
      if (ln2_fused) { ep.ln_stats = work.ln_stats; ep.ln_np = slots; ep.ln_dim = D; ep.ln_cs = lw.cs_fc1; ep.ln_wb = lw.wb_fc1; ep.ln_bw = lw.bw_fc1; }
      ProfScope ps(prof, KC_GEMM_FC1, 2.0 * T * w.mlp * D);
      gemm_op16(work.h, D, lw.w_fc1, D, T, w.mlp, D, ep, s); ++nl;
    }
    {
      GemmEpilogue ep;
      ep.bias = lw.b_fc2; ep.out_f32 = work.x; ep.ld_out = D; ep.accumulate = w.residual_mode;
      ep.reverse_m = 1;      // u (201 MB at 8 images) was written first-to-last by fc1: read its L2-resident tail first
      if (ln1_fused && li + 1 < w.L) {      // the next layer's LayerNorm1: operand copy in that layer's row order
        const EncoderLayerW& nw = w.layers[li + 1];
        ep.stats_out = work.ln_stats; ep.ld_x16 = D; ep.x16_gamma = nw.ln1_g;
        if (nw.is_global) { ep.x16_out = work.h; } else { ep.x16_out = work.h_win; ep.x16_rowmap = work.tok_win_map; }
      }
      ProfScope ps(prof, KC_GEMM_FC2, 2.0 * T * w.mlp * D);
      gemm_op16(work.u, w.mlp, lw.w_fc2, w.mlp, T, D, w.mlp, ep, s); ++nl;
    }
    if (hidden_dump)
      YSI_CUDA(cudaMemcpyAsync(hidden_dump + static_cast<size_t>(li + 1) * T * D, work.x, sizeof(float) * T * D,
                               cudaMemcpyDeviceToDevice, s));
  }
  // neck (modeling_sam.py:985-992): 1x1 conv -> LN2d -> 3x3 conv -> LN2d, all in token-major (NHWC) layout.
  // Both convolutions take their input as a two-term op16 split [hi | lo] against [W | W]: their input roundings
  // would otherwise be the two largest error terms of the image embedding (they act on the whole residual stream /
  // the whole normalised map, not on a small branch), and the neck is < 2 % of the encoder's FLOPs.
  {
    ProfScope ps(prof, KC_NECK, 2.0 * T * 256 * (D + 2304.0));
    const long long n8 = static_cast<long long>(T) * D / 8;
    cast_split_op16_kernel<<<static_cast<int>(std::min<long long>((n8 + 255) / 256, 148 * 16)), 256, 0, s>>>(work.x, work.u, n8, D / 8);
    YSI_CUDA(cudaGetLastError()); ++nl;
    GemmEpilogue ep;
    ep.out_f32 = work.n1; ep.ld_out = 256;
    gemm_op16(work.u, 2 * D, w.w_neck1, 2 * D, T, 256, 2 * D, ep, s); ++nl;
    launch_layernorm(work.n1, T, 256, w.neck_ln1_g, w.neck_ln1_b, 1e-6f, work.n1b, nullptr, false, s, /*split=*/true); ++nl;
    static const bool implicit = [] { const char* e = getenv("YSI_NECK_IMPLICIT"); return e ? atoi(e) != 0 : true; }();
    if (implicit) {
      // 3x3 convolution as an implicit GEMM: the nine shifted views of n1b are fetched by 4-D TMA boxes (zero fill = padding)
      gemm_conv3x3_grid(work.n1b, n, NECK_C2, w.w_neck2, work.n2, 256, s); ++nl;
    } else {
      const long long tot = static_cast<long long>(T) * 9 * 64;
      im2col_3x3_kernel<<<static_cast<int>(std::min<long long>((tot + 255) / 256, 148 * 16)), 256, 0, s>>>(work.n1b, n, work.a_neck);
      YSI_CUDA(cudaGetLastError()); ++nl;
      GemmEpilogue ep2;
      ep2.out_f32 = work.n2; ep2.ld_out = 256;
      gemm_op16(work.a_neck, NECK_K2, w.w_neck2, NECK_K2, T, 256, NECK_K2, ep2, s); ++nl;
    }
    launch_layernorm(work.n2, T, 256, w.neck_ln2_g, w.neck_ln2_b, 1e-6f, nullptr, emb_out, false, s); ++nl;
  }
  *launches += nl;
}

}  // namespace ysi
