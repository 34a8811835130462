// Prompt encoder + two-way-transformer mask decoder + ConvTranspose upscaler, batched over all boxes
// (modeling_sam.py:546-566, 647-698, 273-405, 461-543).
//
// Image-side work (4096 tokens x 256 channels per box) runs on the tcgen05 GEMM:
//   k/v/q projections of the cross attentions, the image->token out-projection (+ residual), and the
//   two stride-2 transposed convolutions as GEMMs with fused LayerNorm/GELU/pixel-shuffle and
//   GELU + hypernetwork-dot epilogues (the [32,256,256] upscaled embedding is never materialised).
// Token-side work (7 tokens per box) runs in small fp32 CUDA-core kernels, one CTA per box.
// Block-0 image-side projections are box independent (keys = image_emb + no_mask_embed for every box)
// and are computed once per image.
#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "kernels.h"

namespace ysi {

constexpr int NT = 7;       // tokens per box: iou, 4 mask tokens, 2 box corners
constexpr int C = 256;
constexpr float TWO_PI = 6.283185307179586f;

// ---------------------------------------------------------------------------------------------------
// positional encoding (SamPositionalEmbedding.forward :552-566): c in [0,1]^2 -> [sin(2pi (2c-1)G), cos(...)]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pe_encode(float cx, float cy, const float* __restrict__ gauss, int j, float& s, float& c) {
  // coordinates @ positional_embedding ([..,2] x [2,128]) then * 2*pi, fp32
  float v = __fadd_rn(__fmul_rn(cx, gauss[j]), __fmul_rn(cy, gauss[128 + j]));
  v = __fmul_rn(TWO_PI, v);
  s = sinf(v);
  c = cosf(v);
}

// image-wide PE (:1128-1139): grid ((j+0.5)/64, (i+0.5)/64) -> token-major [4096,256]
__global__ void image_pe_kernel(const float* __restrict__ gauss, float* __restrict__ pe) {
  const int tok = blockIdx.x, j = threadIdx.x;   // 128 threads
  const int i = tok >> 6, jj = tok & 63;
  const float y = (static_cast<float>(i + 1) - 0.5f) / 64.0f, x = (static_cast<float>(jj + 1) - 0.5f) / 64.0f;
  float s, c;
  pe_encode(2.0f * x - 1.0f, 2.0f * y - 1.0f, gauss, j, s, c);
  pe[tok * 256 + j] = s;
  pe[tok * 256 + 128 + j] = c;
}

void launch_image_pe(const float* gauss, float* image_pe, cudaStream_t s) {
  image_pe_kernel<<<4096, 128, 0, s>>>(gauss, image_pe);
  YSI_CUDA(cudaGetLastError());
}

// tokens [nb,7,256] = [iou_token; mask_tokens(4); corner1; corner2]  (:489-496, :647-656)
__global__ void prompt_tokens_kernel(const double* __restrict__ boxes1024, DecoderW w, float* __restrict__ tok0,
                                     float* __restrict__ sparse_out) {
  const int b = blockIdx.x, t = threadIdx.x;   // 256 threads
  float* out = tok0 + static_cast<size_t>(b) * NT * C;
  out[t] = w.iou_token[t];
  for (int m = 0; m < 4; ++m) out[(1 + m) * C + t] = w.mask_tokens[m * C + t];
  const int j = t & 127;
  for (int corner = 0; corner < 2; ++corner) {
    // fp64: (box + 0.5) / 1024, 2c - 1; then cast to fp32 (:649-651, :556-562)
    const double bx = (boxes1024[b * 4 + 2 * corner] + 0.5) / 1024.0;
    const double by = (boxes1024[b * 4 + 2 * corner + 1] + 0.5) / 1024.0;
    const float cx = static_cast<float>(2.0 * bx - 1.0), cy = static_cast<float>(2.0 * by - 1.0);
    float s, c;
    pe_encode(cx, cy, w.gauss, j, s, c);
    const float v = (t < 128 ? s : c) + w.point_embed[(2 + corner) * C + t];
    out[(5 + corner) * C + t] = v;
    if (sparse_out) sparse_out[(static_cast<size_t>(b) * 2 + corner) * C + t] = v;
  }
}

// keys0 = emb + no_mask_embed ; op16 copies of keys0 and keys0 + pe   (:499, :320-321)
__global__ void prep_keys_kernel(const float* __restrict__ emb, const float* __restrict__ no_mask,
                                 const float* __restrict__ pe, long long n4, float* __restrict__ keys,
                                 op16* __restrict__ keys_bf, op16* __restrict__ keyspos_bf) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i & 63);
    const long long row = i >> 6;
    float4 e = reinterpret_cast<const float4*>(emb)[i];
    const float4 nm = reinterpret_cast<const float4*>(no_mask)[c4];
    const float4 p = reinterpret_cast<const float4*>(pe)[(row & 4095) * 64 + c4];
    e.x += nm.x; e.y += nm.y; e.z += nm.z; e.w += nm.w;
    reinterpret_cast<float4*>(keys)[i] = e;
    uint2 o;
    o.x = pack_op16x2(e.x, e.y); o.y = pack_op16x2(e.z, e.w);
    reinterpret_cast<uint2*>(keys_bf)[i] = o;
    o.x = pack_op16x2(e.x + p.x, e.y + p.y); o.y = pack_op16x2(e.z + p.z, e.w + p.w);
    reinterpret_cast<uint2*>(keyspos_bf)[i] = o;
  }
}

// LayerNorm4 over per-box keys (warp per row) -> fp32 keys + op16 keys as a two-term split [hi(256) | lo(256)] (row
// pitch KEYS_LD; the attention projections read the hi half, the upscaler both) + op16(keys + pe)   (:343-347)
constexpr int KEYS_LD = 512;

__global__ void __launch_bounds__(256)
keys_ln_kernel(const float* __restrict__ in, long long rows, const float* __restrict__ g, const float* __restrict__ bta,
               const float* __restrict__ pe, float* __restrict__ keys, op16* __restrict__ keys_bf,
               op16* __restrict__ keyspos_bf, int want_lo) {
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(in + row * C);
  float4 v[2] = {xr[lane], xr[lane + 32]};
  float sum = (v[0].x + v[0].y) + (v[0].z + v[0].w) + (v[1].x + v[1].y) + (v[1].z + v[1].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
  const float rstd = rsqrtf(sq / C + 1e-6f);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = lane + 32 * k;
    const float4 gg = reinterpret_cast<const float4*>(g)[i], bb = reinterpret_cast<const float4*>(bta)[i];
    const float4 p = reinterpret_cast<const float4*>(pe)[(row & 4095) * 64 + i];
    float4 y;
    y.x = (v[k].x - mean) * rstd * gg.x + bb.x; y.y = (v[k].y - mean) * rstd * gg.y + bb.y;
    y.z = (v[k].z - mean) * rstd * gg.z + bb.z; y.w = (v[k].w - mean) * rstd * gg.w + bb.w;
    if (keys) reinterpret_cast<float4*>(keys + row * C)[i] = y;       // fp32 copy: only the next block's residual needs it
    uint2 o;
    o.x = pack_op16x2(y.x, y.y); o.y = pack_op16x2(y.z, y.w);
    reinterpret_cast<uint2*>(keys_bf + row * KEYS_LD)[i] = o;
    if (want_lo) {                                                   // lo terms: only the upscaler (after the last block) reads them
      o.x = pack_op16x2(y.x - op2f(f2op(y.x)), y.y - op2f(f2op(y.y))); o.y = pack_op16x2(y.z - op2f(f2op(y.z)), y.w - op2f(f2op(y.w)));
      reinterpret_cast<uint2*>(keys_bf + row * KEYS_LD + C)[i] = o;
    }
    o.x = pack_op16x2(y.x + p.x, y.y + p.y); o.y = pack_op16x2(y.z + p.z, y.w + p.w);
    reinterpret_cast<uint2*>(keyspos_bf + row * C)[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// token-side building blocks. The 7 tokens of every box are rows of one [R = 7 nb, C] fp32 matrix, and each
// linear layer is one launch over all boxes: a warp owns one output column (its weight row lives in
// registers) and walks the rows four at a time, so the grid is N/8 CTAs instead of one CTA per box.
// ---------------------------------------------------------------------------------------------------
struct TokLin {
  const float* X;  const float* Xadd;  int ldx;     // input rows (optionally X + Xadd), row pitch ldx
  const float* W;  const float* b;                   // [N, K], [N]
  const float* res;  int ldres;                      // optional residual added to the output
  float* Y;  int ldy;
  int K, N, relu;                                    // K multiple of 128, <= 2048
  const op16* W3;                                    // optional [N, 3K] three-term split of W (tensor-core path at many boxes)
};
struct TokLin3 { TokLin t[3]; };

constexpr int TOK_ROWS = 8;     // rows per CTA

__global__ void __launch_bounds__(256)
tok_linear_kernel(TokLin3 P, int R) {
  const TokLin& p = P.t[blockIdx.z];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  const int nk4 = p.K >> 7;
  float4 w[16];
  const float4* wr = reinterpret_cast<const float4*>(p.W + static_cast<size_t>(n) * p.K);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < nk4) w[i] = __ldg(wr + i * 32 + lane);
  const float bias = p.b ? __ldg(p.b + n) : 0.f;
  const int rbeg = blockIdx.y * TOK_ROWS, rend = min(R, rbeg + TOK_ROWS);
  for (int r0 = rbeg; r0 < rend; r0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < nk4) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          if (r0 + rr < rend) {
            float4 x = __ldg(reinterpret_cast<const float4*>(p.X + static_cast<size_t>(r0 + rr) * p.ldx) + i * 32 + lane);
            if (p.Xadd) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p.Xadd + static_cast<size_t>(r0 + rr) * p.ldx) + i * 32 + lane);
              x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
            }
            acc[rr] = fmaf(x.x, w[i].x, fmaf(x.y, w[i].y, fmaf(x.z, w[i].z, fmaf(x.w, w[i].w, acc[rr]))));
          }
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[rr] += __shfl_xor_sync(0xFFFFFFFFu, acc[rr], o);
    }
    if (lane < 4 && r0 + lane < rend) {
      float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + bias;
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.res) v += p.res[static_cast<size_t>(r0 + lane) * p.ldres + n];
      p.Y[static_cast<size_t>(r0 + lane) * p.ldy + n] = v;
    }
  }
}

// Same contract as tok_linear_kernel for many rows (config 4: 32 boxes/image -> R = 7*256): classic shared-memory
// tiled fp32 GEMM, 64 x 64 output tile per CTA, 4 x 4 outputs per thread, K in slabs of 32, so weights and
// activations are each read from L2 once per tile row / column instead of once per 8 x 8 block.
constexpr int TG_BN = 64, TG_BK = 32;

// BM = 64 or 32 rows per CTA (32: twice the CTAs for the narrow layers, whose 64-row grids leave SMs idle). The next K slab
// is fetched into registers while the current one is multiplied, so a CTA is not exposed to one L2 round trip per slab
// (round-2 ncu of the unpipelined version: 12 % warps active, long_scoreboard the top stall, 179 us for the K = 2048 layer).
template <int BM>
__global__ void __launch_bounds__(256)
tok_gemm_kernel(TokLin3 P, int R) {
  const TokLin& p = P.t[blockIdx.z];
  __shared__ float Xs[BM][TG_BK + 1];
  __shared__ float Ws[TG_BN][TG_BK + 1];
  constexpr int RT = BM / 16;          // output rows per thread
  constexpr int XL = BM / 32;          // float4 loads of X per thread and slab
  const int n0 = blockIdx.x * TG_BN, r0 = blockIdx.y * BM;
  if (n0 >= p.N) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[RT][4];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: 256 threads x float4 = 32 rows x 32 k per pass
  const int lrow = tid >> 3, lk4 = tid & 7;
  float4 xr[XL], wr[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < XL; ++h) {
      const int rr = lrow + 32 * h;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + rr < R) {
        x = __ldg(reinterpret_cast<const float4*>(p.X + static_cast<size_t>(r0 + rr) * p.ldx + k0) + lk4);
        if (p.Xadd) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p.Xadd + static_cast<size_t>(r0 + rr) * p.ldx + k0) + lk4);
          x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
        }
      }
      xr[h] = x;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rr = lrow + 32 * h;
      wr[h] = n0 + rr < p.N ? __ldg(reinterpret_cast<const float4*>(p.W + static_cast<size_t>(n0 + rr) * p.K + k0) + lk4)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < p.K; k0 += TG_BK) {
#pragma unroll
    for (int h = 0; h < XL; ++h) {
      const int rr = lrow + 32 * h;
      Xs[rr][lk4 * 4] = xr[h].x; Xs[rr][lk4 * 4 + 1] = xr[h].y; Xs[rr][lk4 * 4 + 2] = xr[h].z; Xs[rr][lk4 * 4 + 3] = xr[h].w;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rr = lrow + 32 * h;
      Ws[rr][lk4 * 4] = wr[h].x; Ws[rr][lk4 * 4 + 1] = wr[h].y; Ws[rr][lk4 * 4 + 2] = wr[h].z; Ws[rr][lk4 * 4 + 3] = wr[h].w;
    }
    __syncthreads();
    if (k0 + TG_BK < p.K) fetch(k0 + TG_BK);
#pragma unroll
    for (int k = 0; k < TG_BK; ++k) {
      float a[RT], b[4];
#pragma unroll
      for (int i = 0; i < RT; ++i) a[i] = Xs[ty * RT + i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[tx + 16 * j][k];
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const int r = r0 + ty * RT + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= p.N) continue;
      float v = acc[i][j] + (p.b ? __ldg(p.b + n) : 0.f);
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.res) v += p.res[static_cast<size_t>(r) * p.ldres + n];
      p.Y[static_cast<size_t>(r) * p.ldy + n] = v;
    }
  }
}

// fp32 [R, K] (row pitch ldx; optionally X + Xadd) -> op16 [R, 3K] = [x_hi | x_lo | x_hi]: the activation side of the three-term
// split GEMM
__global__ void tok_split3_kernel(const float* __restrict__ X, const float* __restrict__ Xadd, int ldx, int R, int K, op16* __restrict__ A3) {
  const long long n4 = static_cast<long long>(R) * (K / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / (K / 4)), k4 = static_cast<int>(i - static_cast<long long>(r) * (K / 4));
    float4 v = *reinterpret_cast<const float4*>(X + static_cast<size_t>(r) * ldx + 4 * k4);
    if (Xadd) {
      const float4 a = *reinterpret_cast<const float4*>(Xadd + static_cast<size_t>(r) * ldx + 4 * k4);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    uint2 hi, lo;
    hi.x = pack_op16x2(v.x, v.y); hi.y = pack_op16x2(v.z, v.w);
    lo.x = pack_op16x2(v.x - op2f(f2op(v.x)), v.y - op2f(f2op(v.y))); lo.y = pack_op16x2(v.z - op2f(f2op(v.z)), v.w - op2f(f2op(v.w)));
    op16* row = A3 + static_cast<size_t>(r) * 3 * K + 4 * k4;
    *reinterpret_cast<uint2*>(row) = hi;
    *reinterpret_cast<uint2*>(row + K) = lo;
    *reinterpret_cast<uint2*>(row + 2 * K) = hi;
  }
}

// LayerNorm over rows of 256 (warp per row) -> out, and optionally out + pe
__global__ void __launch_bounds__(256)
tok_layernorm_kernel(const float* __restrict__ in, int R, const float* __restrict__ g, const float* __restrict__ b, float eps,
                     const float* __restrict__ pe, float* __restrict__ out, float* __restrict__ out_pe) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* x = in + static_cast<size_t>(row) * C;
  float v[8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = x[lane + 32 * k]; sum += v[k]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; sq += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = lane + 32 * k;
    const float y = (v[k] - mean) * rstd * g[c] + b[c];
    out[static_cast<size_t>(row) * C + c] = y;
    if (out_pe) out_pe[static_cast<size_t>(row) * C + c] = y + pe[static_cast<size_t>(row) * C + c];
  }
}

// self-attention core over the 7 tokens of one box: 8 heads x 32 dims, scale 32^-0.5   (:208-231)
__global__ void __launch_bounds__(256)
tok_self_attn_core_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                          float* __restrict__ out) {
  __shared__ float s_q[NT * C], s_k[NT * C], s_v[NT * C];
  __shared__ float s_p[8 * NT * NT];
  const int b = blockIdx.x, t = threadIdx.x;
  const size_t base = static_cast<size_t>(b) * NT * C;
  for (int i = t; i < NT * C; i += 256) { s_q[i] = q[base + i]; s_k[i] = k[base + i]; s_v[i] = v[base + i]; }
  __syncthreads();
  for (int i = t; i < 8 * NT * NT; i += 256) {
    const int h = i / (NT * NT), r = (i / NT) % NT, c = i % NT;
    float acc = 0.f;
    for (int d = 0; d < 32; ++d) acc = fmaf(s_q[r * C + h * 32 + d], s_k[c * C + h * 32 + d], acc);
    s_p[i] = acc * 0.17677669529663687f;
  }
  __syncthreads();
  if (t < 8 * NT) {
    float* row = s_p + t * NT;
    float m = row[0];
    for (int c = 1; c < NT; ++c) m = fmaxf(m, row[c]);
    float sum = 0.f;
    for (int c = 0; c < NT; ++c) { row[c] = expf(row[c] - m); sum += row[c]; }
    for (int c = 0; c < NT; ++c) row[c] /= sum;
  }
  __syncthreads();
  for (int i = t; i < NT * C; i += 256) {
    const int r = i / C, ch = i % C, h = ch / 32;
    float acc = 0.f;
    for (int c = 0; c < NT; ++c) acc = fmaf(s_p[(h * NT + r) * NT + c], s_v[c * C + ch], acc);
    out[base + i] = acc;
  }
}

// 16 consecutive values of an fp32 or op16 row (one head's slice of an image-side K / V / Q projection)
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(op16 v) { return op2f(v); }
__device__ __forceinline__ void load16(const float* p, float (&o)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float4 v = __ldg(q + i); o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w; }
}
__device__ __forceinline__ void load16(const op16* p, float (&o)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint4 v = __ldg(q + i);
    const float2 a = unpack_op16x2(v.x), b = unpack_op16x2(v.y), c = unpack_op16x2(v.z), d = unpack_op16x2(v.w);
    o[8 * i] = a.x; o[8 * i + 1] = a.y; o[8 * i + 2] = b.x; o[8 * i + 3] = b.y;
    o[8 * i + 4] = c.x; o[8 * i + 5] = c.y; o[8 * i + 6] = d.x; o[8 * i + 7] = d.y;
  }
}

// token -> image attention core: 7 queries x 4096 keys, 8 heads x 16, scale 0.25   (:324-327, :398-401)
// grid (8 heads, nb, T2I_SPLIT key ranges): every CTA produces the un-normalised partial (max, sum, P.V) of its
// 1024 keys; t2i_merge_kernel combines the ranges. K: fp32 rows of pitch ldk (head h at columns h*16..),
// V: pitch ldv. group: box -> source sequence (or null = box).
constexpr int T2I_SPLIT = 4;
constexpr int T2I_KEYS = 4096 / T2I_SPLIT;
constexpr int T2I_PART = NT * 2 + NT * 16;     // per (box, head, split): m[7], l[7], o[7][16]

template <typename T>
__global__ void __launch_bounds__(256)
t2i_attention_kernel(const float* __restrict__ q_t2i, const T* __restrict__ K, int ldk, const T* __restrict__ V,
                     int ldv, const int* __restrict__ group, float* __restrict__ part) {
  __shared__ float s_sc[NT * T2I_KEYS];        // 28 KB
  __shared__ float s_qh[NT * 16];
  __shared__ float s_red[8 * NT];
  __shared__ float s_max[NT];
  __shared__ float s_part[2 * NT * 16];
  const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const size_t seq = group ? group[b] : b;
  const T* Kb = K + (seq * 4096 + static_cast<size_t>(sp) * T2I_KEYS) * ldk + h * 16;
  const T* Vb = V + (seq * 4096 + static_cast<size_t>(sp) * T2I_KEYS) * ldv + h * 16;
  float* pout = part + ((static_cast<size_t>(b) * 8 + h) * T2I_SPLIT + sp) * T2I_PART;
  if (t < NT * 16) s_qh[t] = q_t2i[(static_cast<size_t>(b) * NT + t / 16) * 128 + h * 16 + (t % 16)];
  __syncthreads();
  float lmax[NT];
#pragma unroll
  for (int r = 0; r < NT; ++r) lmax[r] = -INFINITY;
#pragma unroll
  for (int it = 0; it < T2I_KEYS / 256; ++it) {
    const int key = t + it * 256;
    float kk[16];
    load16(Kb + static_cast<size_t>(key) * ldk, kk);
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      const float* q = s_qh + r * 16;
      float a = q[0] * kk[0] + q[1] * kk[1] + q[2] * kk[2] + q[3] * kk[3];
      a += q[4] * kk[4] + q[5] * kk[5] + q[6] * kk[6] + q[7] * kk[7];
      a += q[8] * kk[8] + q[9] * kk[9] + q[10] * kk[10] + q[11] * kk[11];
      a += q[12] * kk[12] + q[13] * kk[13] + q[14] * kk[14] + q[15] * kk[15];
      a *= 0.25f;
      s_sc[r * T2I_KEYS + key] = a;
      lmax[r] = fmaxf(lmax[r], a);
    }
  }
#pragma unroll
  for (int r = 0; r < NT; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax[r] = fmaxf(lmax[r], __shfl_xor_sync(0xFFFFFFFFu, lmax[r], o));
    if (lane == 0) s_red[warp * NT + r] = lmax[r];
  }
  __syncthreads();
  if (t < NT) {
    float m = s_red[t];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w * NT + t]);
    s_max[t] = m;
    pout[t] = m;
  }
  __syncthreads();
  float lsum[NT];
#pragma unroll
  for (int r = 0; r < NT; ++r) lsum[r] = 0.f;
#pragma unroll
  for (int it = 0; it < T2I_KEYS / 256; ++it) {
    const int key = t + it * 256;
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      const float pv = expf(s_sc[r * T2I_KEYS + key] - s_max[r]);
      s_sc[r * T2I_KEYS + key] = pv;
      lsum[r] += pv;
    }
  }
#pragma unroll
  for (int r = 0; r < NT; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum[r] += __shfl_xor_sync(0xFFFFFFFFu, lsum[r], o);
    if (lane == 0) s_red[warp * NT + r] = lsum[r];
  }
  __syncthreads();
  if (t < NT) {
    float sm = 0.f;
    for (int w = 0; w < 8; ++w) sm += s_red[w * NT + t];
    pout[NT + t] = sm;
  }
  // o[r][d] = sum_key p[r][key] V[key][d]; 224 threads = 2 key halves x 7 x 16
  if (t < 2 * NT * 16) {
    const int half = t / (NT * 16), idx = t % (NT * 16), r = idx / 16, d = idx % 16;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const float* pr = s_sc + r * T2I_KEYS + half * (T2I_KEYS / 2);
    const T* vp = Vb + static_cast<size_t>(half) * (T2I_KEYS / 2) * ldv + d;
    auto vat = [&](int key) { return to_f32(__ldg(vp + static_cast<size_t>(key) * ldv)); };
    for (int key = 0; key < T2I_KEYS / 2; key += 4) {
      acc0 = fmaf(pr[key], vat(key), acc0);
      acc1 = fmaf(pr[key + 1], vat(key + 1), acc1);
      acc2 = fmaf(pr[key + 2], vat(key + 2), acc2);
      acc3 = fmaf(pr[key + 3], vat(key + 3), acc3);
    }
    s_part[t] = (acc0 + acc1) + (acc2 + acc3);
  }
  __syncthreads();
  if (t < NT * 16) pout[2 * NT + t] = s_part[t] + s_part[NT * 16 + t];
}

// Many boxes (configs[3]: 32 per image): one CTA per (head, box) streams all 4096 keys ONCE through shared memory
// (coalesced float4 loads, 256 keys per tile) and keeps a running (max, sum, P.V) per query row in registers -- lane =
// (key lane 0..3, row 0..7): the 8 row-lanes of a key read its K / V slice as one broadcast -- instead of writing the
// 7 x 1024 scores to shared memory, re-reading them twice and fetching V element-wise seven times.
constexpr int T2O_TK = 256;

template <typename T>
__global__ void __launch_bounds__(256)
t2i_attention_online_kernel(const float* __restrict__ q_t2i, const T* __restrict__ K, int ldk, const T* __restrict__ V,
                            int ldv, const int* __restrict__ group, float* __restrict__ attn_out) {
  __shared__ float4 s_k[T2O_TK][4];
  __shared__ float4 s_v[T2O_TK][4];
  __shared__ float s_red[8][NT][18];
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int r = lane & 7, kl = lane >> 3;
  const size_t seq = group ? group[b] : b;
  const T* Kb = K + seq * 4096 * ldk + h * 16;
  const T* Vb = V + seq * 4096 * ldv + h * 16;
  float q[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)      // scale 16^-0.5 and log2(e): softmax in base 2 (ex2.approx is one MUFU)
    q[i] = r < NT ? (0.25f * 1.4426950408889634f) * q_t2i[(static_cast<size_t>(b) * NT + r) * 128 + h * 16 + i] : 0.f;
  float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = 0.f;
  for (int tile = 0; tile < 4096 / T2O_TK; ++tile) {
    {
      const size_t key = static_cast<size_t>(tile) * T2O_TK + t;
      float kk[16], vv[16];
      load16(Kb + key * ldk, kk);
      load16(Vb + key * ldv, vv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s_k[t][i] = make_float4(kk[4 * i], kk[4 * i + 1], kk[4 * i + 2], kk[4 * i + 3]);
        s_v[t][i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int kk = warp * 32 + kl + 4 * i;
      float sc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 k4 = s_k[kk][j];
        sc = fmaf(q[4 * j], k4.x, sc); sc = fmaf(q[4 * j + 1], k4.y, sc);
        sc = fmaf(q[4 * j + 2], k4.z, sc); sc = fmaf(q[4 * j + 3], k4.w, sc);
      }
      if (sc > m) {                         // new running maximum of this lane's key subset: rescale (rare after the first keys)
        const float f = ex2_approx(m - sc);
        l *= f;
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] *= f;
        m = sc;
      }
      const float pv = ex2_approx(sc - m);
      l += pv;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v4 = s_v[kk][j];
        o[4 * j] = fmaf(pv, v4.x, o[4 * j]); o[4 * j + 1] = fmaf(pv, v4.y, o[4 * j + 1]);
        o[4 * j + 2] = fmaf(pv, v4.z, o[4 * j + 2]); o[4 * j + 3] = fmaf(pv, v4.w, o[4 * j + 3]);
      }
    }
    __syncthreads();
  }
  // merge the 4 key lanes of the warp (every lane has seen keys: m is finite), then the 8 warps
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xFFFFFFFFu, m, off), l2 = __shfl_xor_sync(0xFFFFFFFFu, l, off);
    const float mn = fmaxf(m, m2), f1 = ex2_approx(m - mn), f2 = ex2_approx(m2 - mn);
    l = l * f1 + l2 * f2;
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = o[d] * f1 + __shfl_xor_sync(0xFFFFFFFFu, o[d], off) * f2;
    m = mn;
  }
  if (kl == 0 && r < NT) {
    s_red[warp][r][0] = m; s_red[warp][r][1] = l;
#pragma unroll
    for (int d = 0; d < 16; ++d) s_red[warp][r][2 + d] = o[d];
  }
  __syncthreads();
  if (t < NT * 16) {
    const int rr = t / 16, d = t % 16;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) mm = fmaxf(mm, s_red[w][rr][0]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float f = ex2_approx(s_red[w][rr][0] - mm);
      ll = fmaf(f, s_red[w][rr][1], ll);
      oo = fmaf(f, s_red[w][rr][2 + d], oo);
    }
    attn_out[(static_cast<size_t>(b) * NT + rr) * 128 + h * 16 + d] = oo / ll;
  }
}

// 16 consecutive values of a K / V row, fetched one iteration ahead in their storage type and converted when consumed
template <typename T> struct Raw16;
template <> struct Raw16<float> {
  float4 v[4];
  __device__ __forceinline__ void load(const float* p) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
  }
  __device__ __forceinline__ void get(float (&o)[16]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[4 * i] = v[i].x; o[4 * i + 1] = v[i].y; o[4 * i + 2] = v[i].z; o[4 * i + 3] = v[i].w; }
  }
};
template <> struct Raw16<op16> {
  uint4 v[2];
  __device__ __forceinline__ void load(const op16* p) {
    v[0] = __ldg(reinterpret_cast<const uint4*>(p));
    v[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float (&o)[16]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 a = unpack_op16x2(v[i].x), b = unpack_op16x2(v[i].y), c = unpack_op16x2(v[i].z), d = unpack_op16x2(v[i].w);
      o[8 * i] = a.x; o[8 * i + 1] = a.y; o[8 * i + 2] = b.x; o[8 * i + 3] = b.y;
      o[8 * i + 4] = c.x; o[8 * i + 5] = c.y; o[8 * i + 6] = d.x; o[8 * i + 7] = d.y;
    }
  }
};

// Very many boxes (configs[3] at batch 8: 256 per launch): one CTA = (key range, box) for ALL 8 heads. lane = (key slot 0..3,
// head 0..7), 4 warps -> 16 keys per iteration; a thread owns one key at a time for the 7 query rows of its head:
//  * K / V come straight from global memory, one iteration ahead -- the 8 head-lanes of a key read one contiguous 256-byte
//    (op16) row, so nothing is staged in shared memory and no barrier sits in the key loop;
//  * q is read from shared memory (head stride 116 floats: the 8 heads of an instruction hit 8 different bank quads, the 4
//    key slots broadcast); the running (reference, sum, P.V) of the 7 rows stay in registers: 224 FMAs per 28 LDS.128;
//  * fp32 has the range to never rescale in steady state: the reference exponent only moves when a score exceeds it by 2^32.
// The streaming kernel above spends its time in shared-memory reads (ncu: l1tex 90 %, issue 23-30 %).
// Output: the (max, sum, P.V) partials of t2i_attention_kernel (max converted to natural-log units), merged by t2i_merge_kernel.
constexpr int T2R_THREADS = 128;
constexpr int T2R_QP = NT * 16 + 4;          // floats per head in shared memory

template <typename T>
__global__ void __launch_bounds__(T2R_THREADS)
t2i_attention_rows_kernel(const float* __restrict__ q_t2i, const T* __restrict__ K, int ldk, const T* __restrict__ V,
                          int ldv, const int* __restrict__ group, float* __restrict__ part) {
  __shared__ __align__(16) float s_q[8 * T2R_QP];
  __shared__ float s_red[4][8][T2I_PART];
  const int sp = blockIdx.x, b = blockIdx.y, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int h = lane & 7, ks = warp * 4 + (lane >> 3);
  for (int i = t; i < NT * 128; i += T2R_THREADS) {      // scale 16^-0.5 and log2(e): softmax in base 2
    const int r = i >> 7, c = i & 127;
    s_q[(c >> 4) * T2R_QP + r * 16 + (c & 15)] = (0.25f * 1.4426950408889634f) * q_t2i[(static_cast<size_t>(b) * NT + r) * 128 + c];
  }
  __syncthreads();
  const size_t seq = group ? group[b] : b;
  const T* Kb = K + (seq * 4096 + static_cast<size_t>(sp) * T2I_KEYS + ks) * ldk + h * 16;
  const T* Vb = V + (seq * 4096 + static_cast<size_t>(sp) * T2I_KEYS + ks) * ldv + h * 16;
  float m[NT], l[NT], o[NT][16];
#pragma unroll
  for (int r = 0; r < NT; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) o[r][d] = 0.f;
  }
  Raw16<T> kn, vn;
  kn.load(Kb);
  vn.load(Vb);
  constexpr int NIT = T2I_KEYS / 16;
  const float4* qh = reinterpret_cast<const float4*>(s_q + h * T2R_QP);
  for (int it = 0; it < NIT; ++it) {
    float kk[16], vv[16];
    kn.get(kk);
    vn.get(vv);
    if (it + 1 < NIT) {
      kn.load(Kb + static_cast<size_t>(it + 1) * 16 * ldk);
      vn.load(Vb + static_cast<size_t>(it + 1) * 16 * ldv);
    }
    // all 7 scores first, ONE (rarely taken) branch for the reference update, then the 7 exponentials and the P.V updates:
    // a branch per row would fence the scheduler into one row's dependent FMA chain at a time
    float sc[NT];
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const float4 qa = qh[r * 4 + j], qb = qh[r * 4 + j + 1];
        s0 = fmaf(qa.x, kk[4 * j], s0); s0 = fmaf(qa.y, kk[4 * j + 1], s0); s0 = fmaf(qa.z, kk[4 * j + 2], s0); s0 = fmaf(qa.w, kk[4 * j + 3], s0);
        s1 = fmaf(qb.x, kk[4 * j + 4], s1); s1 = fmaf(qb.y, kk[4 * j + 5], s1); s1 = fmaf(qb.z, kk[4 * j + 6], s1); s1 = fmaf(qb.w, kk[4 * j + 7], s1);
      }
      sc[r] = s0 + s1;
    }
    bool moved = false;
#pragma unroll
    for (int r = 0; r < NT; ++r) moved |= sc[r] > m[r] + 32.0f;
    if (moved) {                               // first key of the thread, then (practically) never again
#pragma unroll
      for (int r = 0; r < NT; ++r) {
        if (sc[r] > m[r] + 32.0f) {
          const float f = ex2_approx(m[r] - sc[r]);
          l[r] *= f;
#pragma unroll
          for (int d = 0; d < 16; ++d) o[r][d] *= f;
          m[r] = sc[r];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      const float pv = ex2_approx(sc[r] - m[r]);
      l[r] += pv;
#pragma unroll
      for (int d = 0; d < 16; ++d) o[r][d] = fmaf(pv, vv[d], o[r][d]);
    }
  }
  // merge the 4 key slots of the warp (every slot has seen keys: m is finite), then the 4 warps through shared memory
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      const float m2 = __shfl_xor_sync(0xFFFFFFFFu, m[r], off), l2 = __shfl_xor_sync(0xFFFFFFFFu, l[r], off);
      const float mn = fmaxf(m[r], m2), f1 = ex2_approx(m[r] - mn), f2 = ex2_approx(m2 - mn);
      l[r] = l[r] * f1 + l2 * f2;
#pragma unroll
      for (int d = 0; d < 16; ++d) o[r][d] = o[r][d] * f1 + __shfl_xor_sync(0xFFFFFFFFu, o[r][d], off) * f2;
      m[r] = mn;
    }
  }
  if (lane < 8) {
    float* dst = &s_red[warp][h][0];
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      dst[r] = m[r]; dst[NT + r] = l[r];
#pragma unroll
      for (int d = 0; d < 16; ++d) dst[2 * NT + r * 16 + d] = o[r][d];
    }
  }
  __syncthreads();
  for (int i = t; i < 8 * NT * 16; i += T2R_THREADS) {
    const int hh = i / (NT * 16), r = (i / 16) % NT, d = i & 15;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) mm = fmaxf(mm, s_red[w][hh][r]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float f = ex2_approx(s_red[w][hh][r] - mm);
      ll = fmaf(f, s_red[w][hh][NT + r], ll);
      oo = fmaf(f, s_red[w][hh][2 * NT + r * 16 + d], oo);
    }
    float* pout = part + ((static_cast<size_t>(b) * 8 + hh) * T2I_SPLIT + sp) * T2I_PART;
    pout[2 * NT + r * 16 + d] = oo;
    if (d == 0) { pout[r] = mm * 0.6931471805599453f; pout[NT + r] = ll; }      // t2i_merge_kernel works in natural-log units
  }
}

// combine the key ranges: out[b][r][h*16+d] = sum_s e^(m_s-m) o_s / sum_s e^(m_s-m) l_s
__global__ void __launch_bounds__(256)
t2i_merge_kernel(const float* __restrict__ part, float* __restrict__ attn_out) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < NT * 128; i += 256) {
    const int r = i / 128, hd = i % 128, h = hd / 16, d = hd % 16;
    const float* pp = part + (static_cast<size_t>(b) * 8 + h) * T2I_SPLIT * T2I_PART;
    float m = -INFINITY;
#pragma unroll
    for (int sp = 0; sp < T2I_SPLIT; ++sp) m = fmaxf(m, pp[sp * T2I_PART + r]);
    float l = 0.f, o = 0.f;
#pragma unroll
    for (int sp = 0; sp < T2I_SPLIT; ++sp) {
      const float f = expf(pp[sp * T2I_PART + r] - m);
      l = fmaf(f, pp[sp * T2I_PART + NT + r], l);
      o = fmaf(f, pp[sp * T2I_PART + 2 * NT + r * 16 + d], o);
    }
    attn_out[(static_cast<size_t>(b) * NT + r) * 128 + hd] = o / l;
  }
}

// image -> token attention (:337-341): every image token attends over the 7 tokens of its box.
// Q: fp32 [*,ldq] (columns 128..255 of the fused k|q projection). One thread = (token, head).
template <typename T>
__global__ void __launch_bounds__(256)
i2t_attention_kernel(const T* __restrict__ Q, int ldq, const int* __restrict__ group, const float* __restrict__ k_tok,
                     const float* __restrict__ v_tok, op16* __restrict__ out) {
  // k / v of the box's 7 tokens, head h at floats [h*20, h*20+16) of a 160-float row: the 8 heads of a quarter-warp
  // then read their 16-byte pieces from 8 different bank groups (a 64-byte head stride would be a 4-way conflict)
  __shared__ __align__(16) float s_k[NT * 160], s_v[NT * 160];
  const int b = blockIdx.y, t = threadIdx.x;
  for (int i = t; i < NT * 128; i += 256) {
    const int rr = i >> 7, c = i & 127, j = rr * 160 + (c >> 4) * 20 + (c & 15);
    s_k[j] = k_tok[static_cast<size_t>(b) * NT * 128 + i];
    s_v[j] = v_tok[static_cast<size_t>(b) * NT * 128 + i];
  }
  __syncthreads();
  const int tok = blockIdx.x * 32 + (t >> 3), h = t & 7;
  const size_t seq = group ? group[b] : b;
  float qq[16];
  load16(Q + (seq * 4096 + tok) * ldq + h * 16, qq);
  float4 q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float4(qq[4 * i], qq[4 * i + 1], qq[4 * i + 2], qq[4 * i + 3]);
  float sc[NT], m = -INFINITY;
#pragma unroll
  for (int r = 0; r < NT; ++r) {
    const float4* kp = reinterpret_cast<const float4*>(s_k + r * 160 + h * 20);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 k4 = kp[j];
      a = fmaf(q[j].x, k4.x, a); a = fmaf(q[j].y, k4.y, a); a = fmaf(q[j].z, k4.z, a); a = fmaf(q[j].w, k4.w, a);
    }
    sc[r] = a * (0.25f * 1.4426950408889634f);     // scale 16^-0.5, base-2 softmax
    m = fmaxf(m, sc[r]);
  }
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < NT; ++r) { sc[r] = ex2_approx(sc[r] - m); sum += sc[r]; }
  const float inv = 1.0f / sum;
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
  for (int r = 0; r < NT; ++r) {
    const float4* vp = reinterpret_cast<const float4*>(s_v + r * 160 + h * 20);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v4 = vp[j];
      o[4 * j] = fmaf(sc[r], v4.x, o[4 * j]); o[4 * j + 1] = fmaf(sc[r], v4.y, o[4 * j + 1]);
      o[4 * j + 2] = fmaf(sc[r], v4.z, o[4 * j + 2]); o[4 * j + 3] = fmaf(sc[r], v4.w, o[4 * j + 3]);
    }
  }
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] *= inv;
  uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * 4096 + tok) * 128 + h * 16);
  uint4 v0, v1;
  v0.x = pack_op16x2(o[0], o[1]); v0.y = pack_op16x2(o[2], o[3]); v0.z = pack_op16x2(o[4], o[5]); v0.w = pack_op16x2(o[6], o[7]);
  v1.x = pack_op16x2(o[8], o[9]); v1.y = pack_op16x2(o[10], o[11]); v1.z = pack_op16x2(o[12], o[13]); v1.w = pack_op16x2(o[14], o[15]);
  dst[0] = v0;
  dst[1] = v1;
}

// Same contract, four image tokens per thread (tokens tok0 + 32 i of the CTA's 128): the 7 token keys / values of the
// thread's head are read from shared memory once per FOUR tokens -- the one-token kernel above is bound by those reads
// (ncu at 256 boxes: l1tex 95 %, issue 34 %). grid (32, nb).
template <typename T>
__global__ void __launch_bounds__(256)
i2t_attention4_kernel(const T* __restrict__ Q, int ldq, const int* __restrict__ group, const float* __restrict__ k_tok,
                      const float* __restrict__ v_tok, op16* __restrict__ out) {
  __shared__ __align__(16) float s_k[NT * 160], s_v[NT * 160];
  const int b = blockIdx.y, t = threadIdx.x;
  for (int i = t; i < NT * 128; i += 256) {
    const int rr = i >> 7, c = i & 127, j = rr * 160 + (c >> 4) * 20 + (c & 15);
    s_k[j] = k_tok[static_cast<size_t>(b) * NT * 128 + i];
    s_v[j] = v_tok[static_cast<size_t>(b) * NT * 128 + i];
  }
  __syncthreads();
  const int tok0 = blockIdx.x * 128 + (t >> 3), h = t & 7;
  const size_t seq = group ? group[b] : b;
  float q[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i) load16(Q + (seq * 4096 + tok0 + 32 * i) * ldq + h * 16, q[i]);
  float sc[4][NT];
#pragma unroll
  for (int r = 0; r < NT; ++r) {
    const float4* kp = reinterpret_cast<const float4*>(s_k + r * 160 + h * 20);
    const float4 k0 = kp[0], k1 = kp[1], k2 = kp[2], k3 = kp[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = 0.f, c = 0.f;
      a = fmaf(q[i][0], k0.x, a); a = fmaf(q[i][1], k0.y, a); a = fmaf(q[i][2], k0.z, a); a = fmaf(q[i][3], k0.w, a);
      c = fmaf(q[i][4], k1.x, c); c = fmaf(q[i][5], k1.y, c); c = fmaf(q[i][6], k1.z, c); c = fmaf(q[i][7], k1.w, c);
      a = fmaf(q[i][8], k2.x, a); a = fmaf(q[i][9], k2.y, a); a = fmaf(q[i][10], k2.z, a); a = fmaf(q[i][11], k2.w, a);
      c = fmaf(q[i][12], k3.x, c); c = fmaf(q[i][13], k3.y, c); c = fmaf(q[i][14], k3.z, c); c = fmaf(q[i][15], k3.w, c);
      sc[i][r] = (a + c) * (0.25f * 1.4426950408889634f);     // scale 16^-0.5, base-2 softmax
    }
  }
  float o[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m = sc[i][0];
#pragma unroll
    for (int r = 1; r < NT; ++r) m = fmaxf(m, sc[i][r]);
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < NT; ++r) { sc[i][r] = ex2_approx(sc[i][r] - m); sum += sc[i][r]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int r = 0; r < NT; ++r) sc[i][r] *= inv;
#pragma unroll
    for (int d = 0; d < 16; ++d) o[i][d] = 0.f;
  }
#pragma unroll
  for (int r = 0; r < NT; ++r) {
    const float4* vp = reinterpret_cast<const float4*>(s_v + r * 160 + h * 20);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v4 = vp[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i][4 * j] = fmaf(sc[i][r], v4.x, o[i][4 * j]); o[i][4 * j + 1] = fmaf(sc[i][r], v4.y, o[i][4 * j + 1]);
        o[i][4 * j + 2] = fmaf(sc[i][r], v4.z, o[i][4 * j + 2]); o[i][4 * j + 3] = fmaf(sc[i][r], v4.w, o[i][4 * j + 3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * 4096 + tok0 + 32 * i) * 128 + h * 16);
    uint4 v0, v1;
    v0.x = pack_op16x2(o[i][0], o[i][1]); v0.y = pack_op16x2(o[i][2], o[i][3]); v0.z = pack_op16x2(o[i][4], o[i][5]); v0.w = pack_op16x2(o[i][6], o[i][7]);
    v1.x = pack_op16x2(o[i][8], o[i][9]); v1.y = pack_op16x2(o[i][10], o[i][11]); v1.z = pack_op16x2(o[i][12], o[i][13]); v1.w = pack_op16x2(o[i][14], o[i][15]);
    dst[0] = v0;
    dst[1] = v1;
  }
}

// ---------------------------------------------------------------------------------------------------
// upscaler epilogues
// ---------------------------------------------------------------------------------------------------
// ConvTranspose2d(256->64,k2,s2) as a GEMM with N = 4 sub-positions x 64 channels; per sub-position:
// + bias, LayerNorm over the 64 channels (eps 1e-6), GELU, op16, stored pixel-shuffled   (:515-520)
struct EpiConvT1 {
  const float *bias, *g, *b;
  op16* out;   // [nb*16384, 128]: 64 channels as a two-term split [hi | lo]
  __device__ __forceinline__ void finish(EpiCtx&) const {}
  // M is a multiple of 4096 (whole boxes), so every row of a tile exists; the 32 rows of a warp are 32 consecutive x of one
  // (box, y): for a sub-position their [hi | lo] segments lie 512 bytes apart and are written as two coalesced slabs.
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx& ctx) const {
    const int box = ctx.row0 >> 12, tok0 = ctx.row0 & 4095, y = tok0 >> 6, x0 = tok0 & 63;
    for (int sp = c_begin / 64; sp < c_end / 64; ++sp) {
      float mean, rstd;
      {
        uint32_t r0[32], r1[32];
        tmem_ld_x32(taddr_row + sp * 64, r0);
        tmem_ld_x32(taddr_row + sp * 64 + 32, r1);
        tmem_ld_wait();
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float x = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]) + __ldg(bias + i);
          if (i < 32) r0[i] = __float_as_uint(x); else r1[i - 32] = __float_as_uint(x);
          sum += x;
        }
        mean = sum * (1.0f / 64.0f);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) { const float d = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]) - mean; sq += d * d; }
        rstd = rsqrtf(sq * (1.0f / 64.0f) + 1e-6f);
      }
      const int dy = sp >> 1, dx = sp & 1;
      op16* dst0 = out + (static_cast<size_t>(box) * 16384 + (2 * y + dy) * 128 + (2 * x0 + dx)) * 128;     // row x0 of the slab
      // 32 channels at a time (re-read from the accumulator: holding all 64 values across the GELU chains spills): hi and lo
      // segments of 64 bytes per row, one 2 KB slab each
      for (int h = 0; h < 2; ++h) {
        uint32_t a[32];
        tmem_ld_x32(taddr_row + sp * 64 + 32 * h, a);
        tmem_ld_wait();
        const float *bh = bias + 32 * h, *gh = g + 32 * h, *bth = b + 32 * h;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {          // packed erf-GELU (|abs err| <= 1.5e-7), as in the fc1 epilogue
            const int i = 8 * c + 2 * k;
            const float x0v = __uint_as_float(a[i]) + __ldg(bh + i), x1v = __uint_as_float(a[i + 1]) + __ldg(bh + i + 1);
            const float2 gg = gelu_erf2(make_float2((x0v - mean) * rstd * __ldg(gh + i) + __ldg(bth + i),
                                                    (x1v - mean) * rstd * __ldg(gh + i + 1) + __ldg(bth + i + 1)));
            hw[k] = pack_op16x2(gg.x, gg.y);
            lw[k] = pack_op16x2(gg.x - op2f(f2op(gg.x)), gg.y - op2f(f2op(gg.y)));
          }
          slab64_put(ctx.smem, ctx.lane, c, hw[0], hw[1], hw[2], hw[3]);
          slab64_put(ctx.smem + 2048u, ctx.lane, c, lw[0], lw[1], lw[2], lw[3]);
        }
        slab64_flush(ctx.smem, ctx.lane, dst0 + 32 * h, 2 * 128 * sizeof(op16));
        slab64_flush(ctx.smem + 2048u, ctx.lane, dst0 + 64 + 32 * h, 2 * 128 * sizeof(op16));
      }
    }
  }
};
template <>
struct EpiTraits<EpiConvT1> {
  static constexpr bool FULL_ROW = false;
  static constexpr int WARP_SMEM = 4096;
};

// image->token out-projection (+ bias, + residual keys) fused with LayerNorm4 and the three operand copies the next
// stages read (:343-347, what keys_ln_kernel did from an fp32 intermediate): keys (fp32, optional), op16 keys as
// [hi | lo] (lo optional) and op16(keys + pe).  N = BN = 256: a tile holds whole rows (EpiTraits::FULL_ROW); a thread
// owns one row, keeps the pre-LayerNorm values in its accumulator columns (tcgen05.st) between the mean, variance and
// output passes, and every global access goes through the warp's shared-memory slab (full 128-byte lines).
struct EpiKeysLN {
  const float* bias;       // [256] out_proj bias
  const float* res;        // residual rows (fp32, pitch 256): the previous block's keys
  const int* res_group;    // block 0: box -> image (the residual is the image's keys, shared by its boxes); null: one row per GEMM row
  const float *g, *b;      // LayerNorm4
  const float* pe;         // [4096,256] image positional encoding
  float* keys;             // fp32 [M,256] or null
  op16* keys_bf;           // [M,KEYS_LD]: hi at column 0, lo at column 256 (want_lo)
  op16* keyspos_bf;        // [M,256]
  int want_lo;
  __device__ __forceinline__ void finish(EpiCtx&) const {}
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx& ctx) const {
    const int lane = ctx.lane, row0 = ctx.row0;          // 32 consecutive tokens of one box (M is a multiple of 4096)
    const size_t rrow0 = res_group ? static_cast<size_t>(res_group[row0 >> 12]) * 4096 + (row0 & 4095) : static_cast<size_t>(row0);
    const float* rbase = res + rrow0 * C;
    const float* pbase = pe + static_cast<size_t>(row0 & 4095) * C;
    // pass 1: v = (acc + bias) + residual, kept in the accumulator columns; row sum. The residual slab of the next 32 columns
    // is in flight while the current one is transposed and added (a warp otherwise has 4 KB of loads outstanding at a time,
    // and eight such warps per SM do not cover the memory latency: 3.1 TB/s measured without the prefetch).
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    uint4 nxt[8];
    slab_load_issue(lane, rbase, C * sizeof(float), nxt);
#pragma unroll
    for (int c = 0; c < C; c += 32) {
      uint32_t a[32], r[32];
      uint4 cur[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      if (c + 32 < C) slab_load_issue(lane, rbase + c + 32, C * sizeof(float), nxt);
      else slab_load_issue(lane, pbase, C * sizeof(float), nxt);            // first pe slab of pass 3
      tmem_ld_x32(taddr_row + c, a);
      slab_load_finish(ctx.smem, lane, cur, r);
      tmem_ld_wait();
      const float4* b4 = reinterpret_cast<const float4*>(bias + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = __ldg(b4 + i);
        const float v0 = (__uint_as_float(a[4 * i]) + bb.x) + __uint_as_float(r[4 * i]);
        const float v1 = (__uint_as_float(a[4 * i + 1]) + bb.y) + __uint_as_float(r[4 * i + 1]);
        const float v2 = (__uint_as_float(a[4 * i + 2]) + bb.z) + __uint_as_float(r[4 * i + 2]);
        const float v3 = (__uint_as_float(a[4 * i + 3]) + bb.w) + __uint_as_float(r[4 * i + 3]);
        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
        a[4 * i] = __float_as_uint(v0); a[4 * i + 1] = __float_as_uint(v1);
        a[4 * i + 2] = __float_as_uint(v2); a[4 * i + 3] = __float_as_uint(v3);
      }
      tmem_st_x32p(taddr_row + c, a);
    }
    tmem_st_wait();
    const float mean = ((s0 + s1) + (s2 + s3)) / C;
    // pass 2: centred second moment
    s0 = s1 = s2 = s3 = 0.f;
    for (int c = 0; c < C; c += 32) {
      uint32_t a[32];
      tmem_ld_x32(taddr_row + c, a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float d0 = __uint_as_float(a[i]) - mean, d1 = __uint_as_float(a[i + 1]) - mean;
        const float d2 = __uint_as_float(a[i + 2]) - mean, d3 = __uint_as_float(a[i + 3]) - mean;
        s0 += d0 * d0; s1 += d1 * d1; s2 += d2 * d2; s3 += d3 * d3;
      }
    }
    const float rstd = rsqrtf(((s0 + s1) + (s2 + s3)) / C + 1e-6f);
    // pass 3: normalise 32 columns at a time and write: fp32 rows as one 128-byte slab, the op16 outputs as 64-byte slabs
    // (two fit the warp's 4 KB); the pe slab of the next 32 columns is in flight meanwhile
#pragma unroll
    for (int c = 0; c < C; c += 32) {
      uint32_t a[32], pp[32];
      uint4 cur[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      if (c + 32 < C) slab_load_issue(lane, pbase + c + 32, C * sizeof(float), nxt);
      tmem_ld_x32(taddr_row + c, a);
      slab_load_finish(ctx.smem, lane, cur, pp);
      tmem_ld_wait();
      const float4* g4 = reinterpret_cast<const float4*>(g + c);
      const float4* b4 = reinterpret_cast<const float4*>(b + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 gg = __ldg(g4 + i), bb = __ldg(b4 + i);
        a[4 * i] = __float_as_uint((__uint_as_float(a[4 * i]) - mean) * rstd * gg.x + bb.x);
        a[4 * i + 1] = __float_as_uint((__uint_as_float(a[4 * i + 1]) - mean) * rstd * gg.y + bb.y);
        a[4 * i + 2] = __float_as_uint((__uint_as_float(a[4 * i + 2]) - mean) * rstd * gg.z + bb.z);
        a[4 * i + 3] = __float_as_uint((__uint_as_float(a[4 * i + 3]) - mean) * rstd * gg.w + bb.w);
      }
      if (keys) slab_store(ctx.smem, lane, a, keys + static_cast<size_t>(row0) * C + c, C * sizeof(float));
#pragma unroll
      for (int p4 = 0; p4 < 4; ++p4) {          // 8 columns = one 16-byte piece of the hi / lo slabs
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = __uint_as_float(a[8 * p4 + k]);
        slab64_put(ctx.smem, lane, p4, pack_op16x2(y[0], y[1]), pack_op16x2(y[2], y[3]), pack_op16x2(y[4], y[5]), pack_op16x2(y[6], y[7]));
        if (want_lo)
          slab64_put(ctx.smem + 2048u, lane, p4, pack_op16x2(y[0] - op2f(f2op(y[0])), y[1] - op2f(f2op(y[1]))),
                     pack_op16x2(y[2] - op2f(f2op(y[2])), y[3] - op2f(f2op(y[3]))),
                     pack_op16x2(y[4] - op2f(f2op(y[4])), y[5] - op2f(f2op(y[5]))),
                     pack_op16x2(y[6] - op2f(f2op(y[6])), y[7] - op2f(f2op(y[7]))));
      }
      slab64_flush(ctx.smem, lane, keys_bf + static_cast<size_t>(row0) * KEYS_LD + c, KEYS_LD * sizeof(op16));
      if (want_lo) slab64_flush(ctx.smem + 2048u, lane, keys_bf + static_cast<size_t>(row0) * KEYS_LD + C + c, KEYS_LD * sizeof(op16));
#pragma unroll
      for (int p4 = 0; p4 < 4; ++p4) {
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = __uint_as_float(a[8 * p4 + k]) + __uint_as_float(pp[8 * p4 + k]);
        slab64_put(ctx.smem, lane, p4, pack_op16x2(y[0], y[1]), pack_op16x2(y[2], y[3]), pack_op16x2(y[4], y[5]), pack_op16x2(y[6], y[7]));
      }
      slab64_flush(ctx.smem, lane, keyspos_bf + static_cast<size_t>(row0) * C + c, C * sizeof(op16));
    }
  }
};
template <>
struct EpiTraits<EpiKeysLN> {
  static constexpr bool FULL_ROW = true;
  static constexpr int WARP_SMEM = 4096;
};

// ConvTranspose2d(64->32,k2,s2) as a GEMM with N = 4 x 32; per sub-position: + bias, GELU, dot with the
// box's hypernetwork vector -> one low-res logit   (:521, :531)
struct EpiConvT2 {
  const float* bias;    // [32]
  const float* hyper;   // [nb,32]
  float* low;           // [nb,256,256]
  __device__ __forceinline__ void finish(EpiCtx&) const {}
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx&) const {
    const bool active = row < M;
    const int box = row >> 14, pos = row & 16383, Y = pos >> 7, X = pos & 127;
    for (int sp = c_begin / 32; sp < c_end / 32; ++sp) {
      uint32_t r[32];
      tmem_ld_x32(taddr_row + sp * 32, r);
      tmem_ld_wait();
      if (!active) continue;
      float acc = 0.f;
      const float4* b4 = reinterpret_cast<const float4*>(bias);
      const float4* h4 = reinterpret_cast<const float4*>(hyper + box * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {          // 128-bit loads of the bias / hypernetwork vectors (64 scalar loads per 32 values before)
        const float4 bb = __ldg(b4 + i), hh = __ldg(h4 + i);
        const float2 g0 = gelu_erf2(make_float2(__uint_as_float(r[4 * i]) + bb.x, __uint_as_float(r[4 * i + 1]) + bb.y));
        const float2 g1 = gelu_erf2(make_float2(__uint_as_float(r[4 * i + 2]) + bb.z, __uint_as_float(r[4 * i + 3]) + bb.w));
        acc = fmaf(g0.x, hh.x, acc); acc = fmaf(g0.y, hh.y, acc);
        acc = fmaf(g1.x, hh.z, acc); acc = fmaf(g1.y, hh.w, acc);
      }
      low[static_cast<size_t>(box) * 65536 + (2 * Y + (sp >> 1)) * 256 + 2 * X + (sp & 1)] = acc;
    }
  }
};

// ---------------------------------------------------------------------------------------------------
static TokLin tok_lin(const float* X, const float* Xadd, int ldx, const float* W, const float* b, int K, int N, float* Y, int ldy,
                      const float* res = nullptr, int ldres = 0, int relu = 0, const op16* W3 = nullptr) {
  TokLin t;
  t.X = X; t.Xadd = Xadd; t.ldx = ldx; t.W = W; t.b = b; t.res = res; t.ldres = ldres; t.Y = Y; t.ldy = ldy;
  t.K = K; t.N = N; t.relu = relu; t.W3 = W3;
  return t;
}

// token-side linear layers on the tensor cores (many boxes): Y = act((X [+ Xadd]) W^T + b) (+ res) through three-term op16 splits
// on both sides (x_hi W_hi + x_lo W_hi + x_hi W_lo: fp32-level accuracy, the dropped term is 2^-22 relative). Returns the launches.
static int tok_linear_tc(const TokLin* t, int count, int R, op16* a3, cudaStream_t s) {
  int nl = 0;
  const float *lastX = nullptr, *lastA = nullptr;
  int lastK = 0;
  for (int i = 0; i < count; ++i) {
    const TokLin& p = t[i];
    if (p.X != lastX || p.Xadd != lastA || p.K != lastK) {        // consecutive layers on the same input share the split
      const long long n4 = static_cast<long long>(R) * (p.K / 4);
      tok_split3_kernel<<<static_cast<int>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, s>>>(p.X, p.Xadd, p.ldx, R, p.K, a3);
      YSI_CUDA(cudaGetLastError());
      lastX = p.X; lastA = p.Xadd; lastK = p.K; ++nl;
    }
    GemmEpilogue ep;
    ep.bias = p.b; ep.act = p.relu ? ACT_RELU : ACT_NONE; ep.out_f32 = p.Y; ep.ld_out = p.ldy; ep.narrow_tiles = 1;
    if (p.res) { ep.add_src = p.res; ep.ld_add = p.ldres; ep.add_mod = R; }
    gemm_op16(a3, 3 * p.K, p.W3, 3 * p.K, R, p.N, 3 * p.K, ep, s); ++nl;
  }
  return nl;
}

// Row threshold of the tensor-core path (YSI_DEC_TC_MIN_ROWS; default: always). Introduced for many boxes (7 x 256 rows: 0.77 ->
// 0.45 ms); at 1 box per image (56 rows per launch) the kernels are hardly shorter when timed alone (0.42 -> 0.38 ms) but the
// step is 1.7 % faster: the decoder shares the GPU with the next batch's encoder, and a few short GEMM CTAs disturb it less
// than the 19 wide CUDA-core launches they replace.
static int tok_tc_min_rows() {
  static const int v = [] { const char* e = getenv("YSI_DEC_TC_MIN_ROWS"); return e ? atoi(e) : 1; }();
  return v;
}

static int launch_tok_linear(const TokLin* t, int count, int R, cudaStream_t s, op16* a3 = nullptr) {
  static const bool tok_tc = [] { const char* e = getenv("YSI_DEC_MLP_TC"); return e ? atoi(e) != 0 : true; }();
  if (tok_tc && a3 && R >= tok_tc_min_rows()) {
    bool all = true;
    for (int i = 0; i < count; ++i) all = all && t[i].W3 != nullptr && t[i].N % 64 == 0 && t[i].ldx % 4 == 0;
    if (all) return tok_linear_tc(t, count, R, a3, s);
  }
  TokLin3 P;
  int nmax = 0;
  for (int i = 0; i < 3; ++i) {
    P.t[i] = t[i < count ? i : 0];
    if (i < count) {
      YSI_CHECK(t[i].K % 128 == 0 && t[i].K <= 2048 && t[i].N % 8 == 0 && t[i].ldx % 4 == 0, "token linear shape");
      nmax = t[i].N > nmax ? t[i].N : nmax;
    }
  }
  if (R >= 256) {
    const int nt = ceil_div(nmax, TG_BN);
    if (nt * ceil_div(R, 64) * count >= 296) tok_gemm_kernel<64><<<dim3(nt, ceil_div(R, 64), count), 256, 0, s>>>(P, R);
    else tok_gemm_kernel<32><<<dim3(nt, ceil_div(R, 32), count), 256, 0, s>>>(P, R);
  }
  else
    tok_linear_kernel<<<dim3(nmax / 8, ceil_div(R, TOK_ROWS), count), 256, 0, s>>>(P, R);
  return 1;
}

// token -> image attention: few boxes -> key ranges split over CTAs + merge (parallelism); 16..63 -> one streaming CTA
// per (head, box); from 64 boxes up the register-resident kernel (one CTA per key range and box, all heads)
template <typename T>
static int launch_t2i(const float* q, const T* K, int ldk, const T* V, int ldv, const int* group, float* part,
                      float* attn_out, int nb, cudaStream_t s) {
  if (nb >= 64) {
    t2i_attention_rows_kernel<T><<<dim3(T2I_SPLIT, nb), T2R_THREADS, 0, s>>>(q, K, ldk, V, ldv, group, part);
    t2i_merge_kernel<<<nb, 256, 0, s>>>(part, attn_out);
    return 2;
  }
  if (nb >= 16) {
    t2i_attention_online_kernel<T><<<dim3(8, nb), 256, 0, s>>>(q, K, ldk, V, ldv, group, attn_out);
    return 1;
  }
  t2i_attention_kernel<T><<<dim3(8, nb, T2I_SPLIT), 256, 0, s>>>(q, K, ldk, V, ldv, group, part);
  t2i_merge_kernel<<<nb, 256, 0, s>>>(part, attn_out);
  return 2;
}

void decoder_forward(const DecoderW& w, const DecoderWork& wk, const float* emb, int n_img, int nb, float* low_res_out,
                     float* sparse_out, cudaStream_t s, int64_t* launches, Profiler* prof) {
  YSI_CHECK(n_img >= 1 && n_img <= wk.cap_img && nb >= 1 && nb <= wk.cap_box, "decoder batch exceeds the workspace");
  int64_t nl = 0;
  const int TI = n_img * 4096, TB = nb * 4096;
  const int R = NT * nb;                                  // token rows of the whole batch
  // token-side scratch (all [R,256] unless noted)
  float* t_q = wk.tok_ws;            float* t_k = t_q + static_cast<size_t>(R) * C;
  float* t_v = t_k + static_cast<size_t>(R) * C;   float* t_a = t_v + static_cast<size_t>(R) * C;
  float* t_tmp = t_a + static_cast<size_t>(R) * C; float* t_qpe = t_tmp + static_cast<size_t>(R) * C;
  float* t_hid = t_qpe + static_cast<size_t>(R) * C;       // [R,2048]
  float* t_h0 = t_hid + static_cast<size_t>(R) * 2048;     // [nb,256]
  float* t_h1 = t_h0 + static_cast<size_t>(nb) * C;        // [nb,256]
  float* t_part = t_h1 + static_cast<size_t>(nb) * C;      // [nb,8,T2I_SPLIT,T2I_PART] partial token->image attention
  const int ln_blocks = ceil_div(R, 8);
  prompt_tokens_kernel<<<nb, 256, 0, s>>>(wk.boxes1024, w, wk.tok0, sparse_out); ++nl;
  {
    const long long n4 = static_cast<long long>(TI) * 64;
    prep_keys_kernel<<<static_cast<int>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, s>>>(
        emb, w.no_mask_embed, w.image_pe, n4, wk.keys0, wk.keys0_bf, wk.keyspos0_bf); ++nl;
  }
  YSI_CUDA(cudaGetLastError());
  for (int li = 0; li < 2; ++li) {
    const DecLayerW& lw = w.layers[li];
    const bool first = li == 0;
    const bool per_img = li == 0;                       // block 0: keys are shared by all boxes of an image
    const int rows = per_img ? TI : TB;
    const op16* a_keys = per_img ? wk.keys0_bf : wk.keys_bf;
    const op16* a_keyspos = per_img ? wk.keyspos0_bf : wk.keyspos_bf;
    // block 0: k|q and v projections once per image, fp32; block 1 and the final attention: per box, stored as op16 --
    // these [boxes x 4096 x 256] intermediates are what the decoder's time goes into at many boxes per image
    const int* group = per_img ? wk.box_img : nullptr;
    {
      // self attention (:316-321; block 0 has neither PE nor residual), LN1, token->image query projection
      ProfScope ps(prof, KC_DEC_TOKEN);
      const float* qin = first ? wk.tok0 : wk.queries;
      const float* qadd = first ? nullptr : wk.tok0;
      const TokLin qkv[3] = {tok_lin(qin, qadd, C, lw.self_attn.wq, lw.self_attn.bq, C, C, t_q, C, nullptr, 0, 0, lw.self_attn.wq3),
                             tok_lin(qin, qadd, C, lw.self_attn.wk, lw.self_attn.bk, C, C, t_k, C, nullptr, 0, 0, lw.self_attn.wk3),
                             tok_lin(qin, nullptr, C, lw.self_attn.wv, lw.self_attn.bv, C, C, t_v, C, nullptr, 0, 0, lw.self_attn.wv3)};
      nl += launch_tok_linear(qkv, 3, R, s, wk.tok_a3);
      tok_self_attn_core_kernel<<<nb, 256, 0, s>>>(t_q, t_k, t_v, t_a); ++nl;
      const TokLin o = tok_lin(t_a, nullptr, C, lw.self_attn.wo, lw.self_attn.bo, C, C, t_tmp, C, first ? nullptr : wk.queries, C, 0, lw.self_attn.wo3);
      nl += launch_tok_linear(&o, 1, R, s, wk.tok_a3);
      tok_layernorm_kernel<<<ln_blocks, 256, 0, s>>>(t_tmp, R, lw.ln1_g, lw.ln1_b, 1e-6f, wk.tok0, wk.queries, t_qpe); ++nl;
      const TokLin tq = tok_lin(t_qpe, nullptr, C, lw.t2i.wq, lw.t2i.bq, C, 128, wk.q_t2i, 128, nullptr, 0, 0, lw.t2i.wq3);
      nl += launch_tok_linear(&tq, 1, R, s, wk.tok_a3);
      YSI_CUDA(cudaGetLastError());
    }
    {
      ProfScope ps(prof, KC_DEC_GEMM, 2.0 * rows * 384 * 256);
      GemmEpilogue ep;
      ep.bias = lw.b_kq_img;
      if (per_img) { ep.out_f32 = wk.kq0; ep.ld_out = 256; } else { ep.out_op16 = wk.kq16; ep.ld_out_op16 = 256; }
      gemm_op16(a_keyspos, C, lw.w_kq_img, C, rows, 256, C, ep, s); ++nl;
      GemmEpilogue ev;
      ev.bias = lw.t2i.bv;
      if (per_img) { ev.out_f32 = wk.v0; ev.ld_out = 128; } else { ev.out_op16 = wk.v16; ev.ld_out_op16 = 128; }
      gemm_op16(a_keys, per_img ? C : KEYS_LD, lw.w_v_img, C, rows, 128, C, ev, s); ++nl;
    }
    { ProfScope ps(prof, KC_DEC_ATTN);
      if (per_img) nl += launch_t2i(wk.q_t2i, wk.kq0, 256, wk.v0, 128, group, t_part, wk.attn_t2i, nb, s);
      else nl += launch_t2i(wk.q_t2i, wk.kq16, 256, wk.v16, 128, group, t_part, wk.attn_t2i, nb, s); }
    {
      // queries += out_proj(attn); LN2; MLP; LN3; image->token key / value projections   (:328-341)
      ProfScope ps(prof, KC_DEC_TOKEN);
      const TokLin o = tok_lin(wk.attn_t2i, nullptr, 128, lw.t2i.wo, lw.t2i.bo, 128, C, t_tmp, C, wk.queries, C, 0, lw.t2i.wo3);
      nl += launch_tok_linear(&o, 1, R, s, wk.tok_a3);
      tok_layernorm_kernel<<<ln_blocks, 256, 0, s>>>(t_tmp, R, lw.ln2_g, lw.ln2_b, 1e-6f, nullptr, wk.queries, nullptr); ++nl;
      // (from 64 boxes per launch these run on the tensor cores through three-term op16 splits: launch_tok_linear / tok_linear_tc)
      const TokLin f1 = tok_lin(wk.queries, nullptr, C, lw.w_fc1, lw.b_fc1, C, 2048, t_hid, 2048, nullptr, 0, 1, lw.w_fc1_s3);
      nl += launch_tok_linear(&f1, 1, R, s, wk.tok_a3);
      const TokLin f2 = tok_lin(t_hid, nullptr, 2048, lw.w_fc2, lw.b_fc2, 2048, C, t_tmp, C, wk.queries, C, 0, lw.w_fc2_s3);
      nl += launch_tok_linear(&f2, 1, R, s, wk.tok_a3);
      tok_layernorm_kernel<<<ln_blocks, 256, 0, s>>>(t_tmp, R, lw.ln3_g, lw.ln3_b, 1e-6f, wk.tok0, wk.queries, t_qpe); ++nl;
      const TokLin kv[2] = {tok_lin(t_qpe, nullptr, C, lw.i2t.wk, lw.i2t.bk, C, 128, wk.k_tok, 128, nullptr, 0, 0, lw.i2t.wk3),
                            tok_lin(wk.queries, nullptr, C, lw.i2t.wv, lw.i2t.bv, C, 128, wk.v_tok, 128, nullptr, 0, 0, lw.i2t.wv3)};
      nl += launch_tok_linear(kv, 2, R, s, wk.tok_a3);
      YSI_CUDA(cudaGetLastError());
    }
    { ProfScope ps(prof, KC_DEC_ATTN);
      if (nb >= 32) {        // enough boxes to fill the GPU with a quarter of the CTAs: four tokens per thread
        if (per_img) i2t_attention4_kernel<float><<<dim3(32, nb), 256, 0, s>>>(wk.kq0 + 128, 256, group, wk.k_tok, wk.v_tok, wk.attn_i2t);
        else i2t_attention4_kernel<op16><<<dim3(32, nb), 256, 0, s>>>(wk.kq16 + 128, 256, group, wk.k_tok, wk.v_tok, wk.attn_i2t);
      } else if (per_img) i2t_attention_kernel<float><<<dim3(128, nb), 256, 0, s>>>(wk.kq0 + 128, 256, group, wk.k_tok, wk.v_tok, wk.attn_i2t);
      else i2t_attention_kernel<op16><<<dim3(128, nb), 256, 0, s>>>(wk.kq16 + 128, 256, group, wk.k_tok, wk.v_tok, wk.attn_i2t);
      ++nl; }
    YSI_CUDA(cudaGetLastError());
    {
      // keys = keys_prev + out_proj(attn) ; then LN4
      ProfScope ps(prof, KC_DEC_GEMM, 2.0 * TB * 256 * 128);
      const bool last = li == 1;
      static const bool fused_ln = [] { const char* e = getenv("YSI_DEC_FUSED_LN"); return e ? atoi(e) != 0 : true; }();
      if (fused_ln) {
        // one launch: the [boxes x 4096, 256] fp32 pre-LayerNorm intermediate never leaves the SM (EpiKeysLN)
        const CUtensorMap tmA = make_tmap_op16_2d(wk.attn_i2t, TB, 128, 128, GEMM_BM);
        const CUtensorMap tmB = make_tmap_op16_2d(lw.w_i2t_out, 256, 128, 128, 256);
        EpiKeysLN e{lw.i2t.bo, per_img ? wk.keys0 : wk.keys, per_img ? wk.box_img : nullptr, lw.ln4_g, lw.ln4_b, w.image_pe,
                    last ? nullptr : wk.keys, wk.keys_bf, wk.keyspos_bf, last ? 1 : 0};
        launch_gemm<256>(tmA, tmB, TB, 256, 128, e, s); ++nl;
      } else {
        GemmEpilogue ep;
        ep.bias = lw.i2t.bo; ep.out_f32 = wk.kq; ep.ld_out = 256;      // kq (per-box) is free now: reuse as pre-LN buffer
        ep.add_src = per_img ? wk.keys0 : wk.keys; ep.ld_add = 256;
        if (per_img) { ep.add_mod = 4096; ep.add_group = wk.box_img; } else { ep.add_mod = TB; }
        // block 1 reads kq (q columns) in i2t above, which is complete before this GEMM starts (same stream)
        gemm_op16(wk.attn_i2t, 128, lw.w_i2t_out, 128, TB, 256, 128, ep, s); ++nl;
        keys_ln_kernel<<<ceil_div(TB, 8), 256, 0, s>>>(wk.kq, TB, lw.ln4_g, lw.ln4_b, w.image_pe, last ? nullptr : wk.keys, wk.keys_bf,
                                                       wk.keyspos_bf, last ? 1 : 0); ++nl;
      }
      YSI_CUDA(cudaGetLastError());
    }
  }
  // final token -> image attention (:394-404); t_qpe = queries + pe from LN3 of block 1
  {
    ProfScope ps(prof, KC_DEC_TOKEN);
    const TokLin fq = tok_lin(t_qpe, nullptr, C, w.final_attn.wq, w.final_attn.bq, C, 128, wk.q_t2i, 128, nullptr, 0, 0, w.final_attn.wq3);
    nl += launch_tok_linear(&fq, 1, R, s, wk.tok_a3);
  }
  {
    ProfScope ps(prof, KC_DEC_GEMM, 2.0 * TB * 256 * 256);
    GemmEpilogue ek;
    ek.bias = w.final_attn.bk; ek.out_op16 = wk.kq16; ek.ld_out_op16 = 256;    // K in columns 0..127 of kq16
    gemm_op16(wk.keyspos_bf, C, w.w_k_final, C, TB, 128, C, ek, s); ++nl;
    GemmEpilogue ev;
    ev.bias = w.final_attn.bv; ev.out_op16 = wk.v16; ev.ld_out_op16 = 128;
    gemm_op16(wk.keys_bf, KEYS_LD, w.w_v_final, C, TB, 128, C, ev, s); ++nl;
  }
  { ProfScope ps(prof, KC_DEC_ATTN); nl += launch_t2i(wk.q_t2i, wk.kq16, 256, wk.v16, 128, nullptr, t_part, wk.attn_t2i, nb, s); }
  {
    // queries += out_proj(attn); layer_norm_final_attn (eps 1e-5); hypernetwork MLP of mask token 0 (= token row 1)
    ProfScope ps(prof, KC_DEC_TOKEN);
    const TokLin o = tok_lin(wk.attn_t2i, nullptr, 128, w.final_attn.wo, w.final_attn.bo, 128, C, t_tmp, C, wk.queries, C, 0, w.final_attn.wo3);
    nl += launch_tok_linear(&o, 1, R, s, wk.tok_a3);
    tok_layernorm_kernel<<<ln_blocks, 256, 0, s>>>(t_tmp, R, w.lnf_g, w.lnf_b, 1e-5f, nullptr, wk.queries, nullptr); ++nl;
    const TokLin h0 = tok_lin(wk.queries + C, nullptr, NT * C, w.hy_w0, w.hy_b0, C, C, t_h0, C, nullptr, 0, 1);
    launch_tok_linear(&h0, 1, nb, s); ++nl;
    const TokLin h1 = tok_lin(t_h0, nullptr, C, w.hy_w1, w.hy_b1, C, C, t_h1, C, nullptr, 0, 1);
    launch_tok_linear(&h1, 1, nb, s); ++nl;
    const TokLin h2 = tok_lin(t_h1, nullptr, C, w.hy_w2, w.hy_b2, C, 32, wk.hyper, 32);
    launch_tok_linear(&h2, 1, nb, s); ++nl;
    YSI_CUDA(cudaGetLastError());
  }
  // upscaler (:515-531)
  {
    ProfScope ps(prof, KC_DEC_UPSCALE, 2.0 * TB * 256 * 256 + 2.0 * nb * 16384.0 * 128 * 64);
    // both transposed convolutions contract [hi | lo] input splits against [W | W]: these two inputs feed the logits
    // directly, so their op16 rounding would otherwise dominate the decoder's error
    const CUtensorMap tmA = make_tmap_op16_2d(wk.keys_bf, TB, KEYS_LD, KEYS_LD, GEMM_BM);
    const CUtensorMap tmB = make_tmap_op16_2d(w.w_ct1, 256, KEYS_LD, KEYS_LD, 256);
    EpiConvT1 e1{w.b_ct1, w.lnu_g, w.lnu_b, wk.up1};
    launch_gemm<256>(tmA, tmB, TB, 256, KEYS_LD, e1, s); ++nl;
    const int M2 = nb * 16384;
    const CUtensorMap tmA2 = make_tmap_op16_2d(wk.up1, M2, 128, 128, GEMM_BM);
    const CUtensorMap tmB2 = make_tmap_op16_2d(w.w_ct2, 128, 128, 128, 128);
    EpiConvT2 e2{w.b_ct2, wk.hyper, low_res_out};
    launch_gemm<128>(tmA2, tmB2, M2, 128, 128, e2, s); ++nl;
  }
  *launches += nl;
}

}  // namespace ysi
