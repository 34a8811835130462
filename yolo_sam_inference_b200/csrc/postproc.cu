// a6 + a7 of SURVEY.md section 8: bilinear upsample x2 + threshold fused with the per-mask morphometrics.
//
//  K1 upsample_stats   logits[256,256] -> mask bytes + {area, centroid sums, bbox, perimeter code
//                      histogram, intensity histogram, first mixed 2x2 cell}      (HBM-bound part)
//  K2 contour_hull_disk one CTA per mask: trace contours[0] from the first mixed cell, convex hull
//                      (exact integer arithmetic on doubled coordinates), inclusive hull raster
//                      -> hull area + hull perimeter histogram; centre-disk brightness sums.
//
// Arithmetic contracts (all verified bit-for-bit against the CPU oracle in tests/):
//  * bilinear taps are evaluated exactly like ATen's CPU kernel as built in torch 2.11
//    (UpSampleKernel.cpp, HelperInterpLinear):  src = max(fma(scale, dst+0.5, -0.5), 0),
//    out = fma(w0, p0, rn(w1*p1)) along W, then the same along H   (image_processing_sam.py:423-427)
//  * perimeter codes follow skimage.measure.perimeter(neighborhood=4): border = mask & ~erode4(mask),
//    code = 1 + 2*#border 4-neighbours + 10*#border diagonal neighbours (utils/metrics.py:65,69)
//  * contour / hull / raster follow find_contours(level 0.5, 'low') -> ConvexHull -> polygon2mask
//    (utils/metrics.py:31-48), see oracle/metrics_oracle.py for the restated library algorithms.
#include <limits.h>

#include "kernels.h"

namespace ysi {

PostGeom make_post_geom(int H, int W) {
  // image_processing_sam.py:113-122 (_get_preprocess_shape), python float == C double
  PostGeom g;
  g.H = H;
  g.W = W;
  const int longest = H > W ? H : W;
  const double scale = 1024 * 1.0 / longest;
  g.rh = static_cast<int>(H * scale + 0.5);
  g.rw = static_cast<int>(W * scale + 0.5);
  return g;
}

// bin of a perimeter code, -1 if the code has zero weight
__constant__ signed char c_code_bin[50] = {
    -1, -1, -1, -1, -1, 0,  -1, 1,  -1, -1, -1, -1, -1, 2,  -1, 3,  -1, 4,  -1, -1, -1, 5,  -1, 6,  -1,
    7,  -1, 8,  -1, -1, -1, -1, -1, 9,  -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};

struct Lin {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Lin lin_index(int dst, int in_size, int out_size, float scale) {
  Lin r;
  if (in_size == out_size) {
    r.i0 = dst; r.i1 = dst; r.l0 = 1.0f; r.l1 = 0.0f;
    return r;
  }
  float src = fmaxf(__fmaf_rn(scale, static_cast<float>(dst) + 0.5f, -0.5f), 0.0f);
  int i0 = static_cast<int>(floorf(src));
  if (i0 > in_size - 1) i0 = in_size - 1;
  r.i0 = i0;
  r.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  r.l1 = fminf(fmaxf(__fsub_rn(src, static_cast<float>(i0)), 0.0f), 1.0f);
  r.l0 = __fsub_rn(1.0f, r.l1);
  return r;
}

__device__ __forceinline__ float lerp_torch(float w0, float p0, float w1, float p1) {
  return __fmaf_rn(w0, p0, __fmul_rn(w1, p1));
}

// value at (i,k) of the 256->1024 interpolated plane
__device__ __forceinline__ float stage1(const float* __restrict__ low, int i, int k) {
  const Lin a = lin_index(i, 256, 1024, 0.25f);
  const Lin b = lin_index(k, 256, 1024, 0.25f);
  const float* r0 = low + a.i0 * 256;
  const float* r1 = low + a.i1 * 256;
  const float top = lerp_torch(b.l0, __ldg(r0 + b.i0), b.l1, __ldg(r0 + b.i1));
  const float bot = lerp_torch(b.l0, __ldg(r1 + b.i0), b.l1, __ldg(r1 + b.i1));
  return lerp_torch(a.l0, top, a.l1, bot);
}

struct PostParams {
  PostGeom g;
  float sh, sw;   // stage-2 scales: float(rh)/float(H), float(rw)/float(W)
};

__device__ __forceinline__ float logit_at(const float* __restrict__ low, const PostParams& p, int r, int c) {
  if (p.g.rh == p.g.H && p.g.rw == p.g.W) return stage1(low, r, c);
  const Lin a = lin_index(r, p.g.rh, p.g.H, p.sh);
  const Lin b = lin_index(c, p.g.rw, p.g.W, p.sw);
  const float top = lerp_torch(b.l0, stage1(low, a.i0, b.i0), b.l1, stage1(low, a.i0, b.i1));
  const float bot = lerp_torch(b.l0, stage1(low, a.i1, b.i0), b.l1, stage1(low, a.i1, b.i1));
  return lerp_torch(a.l0, top, a.l1, bot);
}

constexpr int TR = 32;    // tile rows
constexpr int TC = 128;   // tile cols (one warp row = 32 lanes x 4 pixels)
constexpr int MP = TC + 8;  // smem pitch

__global__ void init_stats_kernel(MaskStatsDev* stats, int nmask) {
  const int m = blockIdx.x;
  if (m >= nmask) return;
  MaskStatsDev* s = stats + m;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s->mask_hist[i] = 0;
  if (threadIdx.x < YSI_PERIM_BINS) s->perim_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    s->area = 0; s->sum_r = 0; s->sum_c = 0;
    s->min_r = INT_MAX; s->min_c = INT_MAX; s->max_r = -1; s->max_c = -1;
    s->first_cell = 0xFFFFFFFFu;
  }
}

void launch_init_stats(MaskStatsDev* stats, int nmask, cudaStream_t s) {
  if (nmask <= 0) return;
  init_stats_kernel<<<nmask, 256, 0, s>>>(stats, nmask);
  YSI_CUDA(cudaGetLastError());
}

// MODE 0: mask = upsampled logit > 0 (and write it); MODE 1: mask read from masks_in
template <int MODE>
__global__ void __launch_bounds__(256)
upsample_stats_kernel(const float* __restrict__ low_all, const uint8_t* __restrict__ masks_in, PostParams p,
                      const uint16_t* __restrict__ sum3_all, const int* __restrict__ mask_image,
                      uint8_t* __restrict__ masks_out, float* __restrict__ up_logits, MaskStatsDev* __restrict__ stats) {
  __shared__ uint8_t sm[TR + 4][MP];   // mask with halo 2, origin (R0-2, C0-2)
  __shared__ uint8_t sb[TR + 2][MP];   // border with halo 1, origin (R0-1, C0-1)
  __shared__ unsigned int s_hist[8][256];
  __shared__ unsigned long long s_acc[3];
  __shared__ int s_mm[4];
  __shared__ unsigned int s_first;
  __shared__ unsigned int s_perim[YSI_PERIM_BINS];

  const int H = p.g.H, W = p.g.W;
  const int m = blockIdx.z;
  const int R0 = blockIdx.y * TR, C0 = blockIdx.x * TC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t plane = static_cast<size_t>(H) * W;
  const float* low = MODE == 0 ? low_all + static_cast<size_t>(m) * 65536 : nullptr;
  const uint8_t* min_ = MODE == 1 ? masks_in + m * plane : nullptr;
  const bool want_hist = sum3_all != nullptr;
  const uint16_t* sum3 = want_hist ? sum3_all + static_cast<size_t>(mask_image ? mask_image[m] : 0) * plane : nullptr;

  if (want_hist)
    for (int i = tid; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
  if (tid < 3) s_acc[tid] = 0;
  if (tid == 0) { s_mm[0] = INT_MAX; s_mm[1] = INT_MAX; s_mm[2] = -1; s_mm[3] = -1; s_first = 0xFFFFFFFFu; }
  if (tid < YSI_PERIM_BINS) s_perim[tid] = 0;

  // ---- A: mask values on the halo'd tile
  for (int i = tid; i < (TR + 4) * (TC + 4); i += 256) {
    const int lr = i / (TC + 4), lc = i - lr * (TC + 4);
    const int r = R0 - 2 + lr, c = C0 - 2 + lc;
    uint8_t v = 0;
    if (r >= 0 && r < H && c >= 0 && c < W) {
      if (MODE == 0) {
        const float x = logit_at(low, p, r, c);
        v = x > 0.0f ? 1 : 0;
        if (up_logits && lr >= 2 && lr < TR + 2 && lc >= 2 && lc < TC + 2) up_logits[m * plane + static_cast<size_t>(r) * W + c] = x;
      } else {
        v = min_[static_cast<size_t>(r) * W + c] ? 1 : 0;
      }
    }
    sm[lr][lc] = v;
  }
  __syncthreads();
  // ---- B: border = mask & ~erosion(cross); outside the image counts as background
  for (int i = tid; i < (TR + 2) * (TC + 2); i += 256) {
    const int lr = i / (TC + 2), lc = i - lr * (TC + 2);   // border-tile coords; mask-tile coords are +1
    const int mr = lr + 1, mc = lc + 1;
    const uint8_t v = sm[mr][mc];
    sb[lr][lc] = v & (1 ^ (sm[mr - 1][mc] & sm[mr + 1][mc] & sm[mr][mc - 1] & sm[mr][mc + 1]));
  }
  __syncthreads();
  // ---- C: interior pixels. warp -> rows warp, warp+8, ...; lane -> 4 consecutive columns
  unsigned long long area = 0, sum_r = 0, sum_c = 0;
  int mnr = INT_MAX, mnc = INT_MAX, mxr = -1, mxc = -1;
  unsigned int first = 0xFFFFFFFFu;
  unsigned int pc[YSI_PERIM_BINS];
#pragma unroll
  for (int b = 0; b < YSI_PERIM_BINS; ++b) pc[b] = 0;
  for (int lr = warp; lr < TR; lr += 8) {
    const int r = R0 + lr;
    if (r >= H) break;
    uint8_t outv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int lc = lane * 4 + k;
      const int c = C0 + lc;
      const bool inimg = c < W;
      const int mr = lr + 2, mc = lc + 2;      // mask-tile coords
      const uint8_t v = inimg ? sm[mr][mc] : 0;
      outv[k] = v;
      if (v) {
        area += 1; sum_r += r; sum_c += c;
        mnr = min(mnr, r); mxr = max(mxr, r); mnc = min(mnc, c); mxc = max(mxc, c);
        if (want_hist) atomicAdd(&s_hist[warp][__ldg(sum3 + static_cast<size_t>(r) * W + c) / 3], 1u);
      }
      // mixed 2x2 cell with this pixel as upper-left (find_contours visits cells in raster order)
      if (inimg && r < H - 1 && c < W - 1) {
        const int sum4 = v + sm[mr][mc + 1] + sm[mr + 1][mc] + sm[mr + 1][mc + 1];
        if (sum4 != 0 && sum4 != 4) first = min(first, static_cast<unsigned int>(r) * W + c);
      }
      int bin = -1;
      const int br = lr + 1, bc = lc + 1;      // border-tile coords
      if (inimg && sb[br][bc]) {
        const int n4 = sb[br - 1][bc] + sb[br + 1][bc] + sb[br][bc - 1] + sb[br][bc + 1];
        const int nd = sb[br - 1][bc - 1] + sb[br - 1][bc + 1] + sb[br + 1][bc - 1] + sb[br + 1][bc + 1];
        bin = c_code_bin[1 + 2 * n4 + 10 * nd];
      }
#pragma unroll
      for (int b = 0; b < YSI_PERIM_BINS; ++b) pc[b] += __popc(__ballot_sync(0xFFFFFFFFu, bin == b));
    }
    if (MODE == 0 && masks_out) {
      const int c = C0 + lane * 4;
      uint8_t* dst = masks_out + m * plane + static_cast<size_t>(r) * W + c;
      if (c + 3 < W && (W & 3) == 0) {
        *reinterpret_cast<uchar4*>(dst) = make_uchar4(outv[0], outv[1], outv[2], outv[3]);
      } else {
        for (int k = 0; k < 4; ++k)
          if (c + k < W) dst[k] = outv[k];
      }
    }
  }
  // ---- reduce: warp -> block -> global
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    area += __shfl_xor_sync(0xFFFFFFFFu, area, o);
    sum_r += __shfl_xor_sync(0xFFFFFFFFu, sum_r, o);
    sum_c += __shfl_xor_sync(0xFFFFFFFFu, sum_c, o);
    mnr = min(mnr, __shfl_xor_sync(0xFFFFFFFFu, mnr, o));
    mnc = min(mnc, __shfl_xor_sync(0xFFFFFFFFu, mnc, o));
    mxr = max(mxr, __shfl_xor_sync(0xFFFFFFFFu, mxr, o));
    mxc = max(mxc, __shfl_xor_sync(0xFFFFFFFFu, mxc, o));
    first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, o));
  }
  if (lane == 0) {
    atomicAdd(&s_acc[0], area); atomicAdd(&s_acc[1], sum_r); atomicAdd(&s_acc[2], sum_c);
    atomicMin(&s_mm[0], mnr); atomicMin(&s_mm[1], mnc); atomicMax(&s_mm[2], mxr); atomicMax(&s_mm[3], mxc);
    atomicMin(&s_first, first);
#pragma unroll
    for (int b = 0; b < YSI_PERIM_BINS; ++b)
      if (pc[b]) atomicAdd(&s_perim[b], pc[b]);
  }
  __syncthreads();
  MaskStatsDev* st = stats + m;
  if (tid == 0) {
    if (s_acc[0]) {
      atomicAdd(&st->area, s_acc[0]); atomicAdd(&st->sum_r, s_acc[1]); atomicAdd(&st->sum_c, s_acc[2]);
      atomicMin(&st->min_r, s_mm[0]); atomicMin(&st->min_c, s_mm[1]);
      atomicMax(&st->max_r, s_mm[2]); atomicMax(&st->max_c, s_mm[3]);
    }
    if (s_first != 0xFFFFFFFFu) atomicMin(&st->first_cell, s_first);
  }
  if (tid < YSI_PERIM_BINS && s_perim[tid]) atomicAdd(&st->perim_hist[tid], s_perim[tid]);
  if (want_hist) {
    unsigned int h = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) h += s_hist[w][tid];
    if (h) atomicAdd(&st->mask_hist[tid], h);
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast path of K1 for the dominant geometry (H = W = 1024: the second resize of post_process_masks is the identity,
// the first is an exact x4 upsample): bit-parallel, ~10 instructions per output pixel instead of ~300.
//
//  * x4 bilinear with align_corners=False has a fixed pattern: output rows 4G-2 .. 4G+1 ("row group" G) blend low
//    rows G-1 and G with l1 = 1/8, 3/8, 5/8, 7/8 (rows 0,1 and 1022,1023 degenerate to one low row through the
//    index clamps); columns likewise.  The taps, their order and the fma contraction are those of lin_index /
//    lerp_torch above, so the result is bit-identical to the generic kernel (and to ATen's CPU kernel).
//  * a lane owns 32 consecutive output columns: it reads 8 floats of a low row (the warp reads the whole 1 KB row,
//    coalesced) + one neighbour each side by shuffle, interpolates horizontally once per low row (kept in registers
//    for the next group) and vertically 4 rows per group, and thresholds straight into a 32-bit row word.
//  * everything downstream works on row words: area / centroid sums by POPC, bounding box by OR-accumulation,
//    border = W & ~(up & down & left & right), the neighbour counts of skimage's perimeter code as bit-sliced adders
//    (LOP3 full adders) and one POPC per code bin, the first mixed 2x2 cell by XORs; the mask bytes are expanded from
//    the row words and written 16 bytes per lane, 512 contiguous bytes per instruction.
//  * a warp walks down a band of rows with a two-row delay line (codes of row r need the border of rows r-1 .. r+1,
//    which needs the mask of rows r-2 .. r+2); one row group above and one below the band are recomputed as halo.
// ---------------------------------------------------------------------------------------------------
// 4 warps per CTA and <= 128 registers per thread: four CTAs = 16 warps per SM (round-2 ncu of the 8-warp, 135-register
// version: ONE resident CTA per SM, 12 % warps active, issue 33 % -- a latency-bound kernel starved of warps)
constexpr int FAST_WARPS = 4;

__device__ __forceinline__ unsigned int shl_in(unsigned int w, unsigned int prev_lane_word) {   // columns c-1 -> bit c
  return (w << 1) | (prev_lane_word >> 31);
}
__device__ __forceinline__ unsigned int shr_in(unsigned int w, unsigned int next_lane_word) {   // columns c+1 -> bit c
  return (w >> 1) | (next_lane_word << 31);
}

__global__ void __launch_bounds__(FAST_WARPS * 32, 4)
upsample_stats_fast_kernel(const float* __restrict__ low_all, const uint8_t* __restrict__ gray_all,
                           const int* __restrict__ mask_image, uint8_t* __restrict__ masks_out,
                           uint8_t* __restrict__ packed_out, MaskStatsDev* __restrict__ stats, int groups_per_warp) {
  constexpr int H = 1024, W = 1024;
  __shared__ unsigned int s_hist[FAST_WARPS][2][256];
  const int m = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* low = low_all + static_cast<size_t>(m) * 65536;
  const bool want_hist = gray_all != nullptr;
  const uint8_t* gray = want_hist ? gray_all + static_cast<size_t>(mask_image ? mask_image[m] : 0) * H * W : nullptr;
  uint8_t* mout = masks_out ? masks_out + static_cast<size_t>(m) * H * W : nullptr;
  uint32_t* pout = reinterpret_cast<uint32_t*>(packed_out + static_cast<size_t>(m) * (H * W / 8));
  if (want_hist)
    for (int i = threadIdx.x; i < FAST_WARPS * 2 * 256; i += FAST_WARPS * 32) (&s_hist[0][0][0])[i] = 0;
  __syncthreads();
  unsigned int* my_hist = &s_hist[warp][lane & 1][0];

  // band of this warp: row groups [ga, gb) -> output rows [4 ga - 2, 4 gb - 2), clipped to the image
  const int ga = (blockIdx.x * FAST_WARPS + warp) * groups_per_warp;
  const int gb = min(ga + groups_per_warp, 257);
  const int R0 = max(4 * ga - 2, 0), R1 = min(4 * gb - 2, H);

  unsigned long long area = 0, sum_r = 0, sum_c = 0;
  unsigned int colmask = 0;                 // OR of all finalised row words: min / max column of this lane
  int mnr = INT_MAX, mxr = -1;
  unsigned int first = 0xFFFFFFFFu;
  unsigned int pc[YSI_PERIM_BINS];
#pragma unroll
  for (int b = 0; b < YSI_PERIM_BINS; ++b) pc[b] = 0;

  if (ga < gb) {
    // horizontal interpolation of one low row into this lane's 32 output columns
    auto hrow = [&](int lr, float (&h)[32]) {
      const float4* src = reinterpret_cast<const float4*>(low + lr * 256 + lane * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      float p[10];
      p[1] = a.x; p[2] = a.y; p[3] = a.z; p[4] = a.w; p[5] = b.x; p[6] = b.y; p[7] = b.z; p[8] = b.w;
      p[0] = __shfl_up_sync(0xFFFFFFFFu, p[8], 1);
      p[9] = __shfl_down_sync(0xFFFFFFFFu, p[1], 1);
      if (lane == 31) p[9] = p[8];                      // i1 clamps to the last low column
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int j0 = (k + 2) >> 2;                    // p index of the left tap (low column 8 lane - 1 + j0)
        const int i = (k + 2) & 3;
        const float l1 = 0.125f + 0.25f * i, l0 = 1.0f - l1;
        h[k] = lerp_torch(l0, p[j0], l1, p[j0 + 1]);
      }
      if (lane == 0) {                                  // output columns 0, 1: src clamps to 0 -> weights (1, 0)
        h[0] = lerp_torch(1.0f, p[1], 0.0f, p[2]);
        h[1] = h[0];
      }
    };
    float hA[32], hB[32];
    // delay line: mask words of the two previous rows, border words (and their column-shifted copies) of rows -3, -2
    unsigned int Wp1 = 0, Wp2 = 0, Bp2 = 0, Bp3 = 0, BLp2 = 0, BRp2 = 0, BLp3 = 0, BRp3 = 0;
    const int g_first = ga - 1, g_last = gb;            // halo groups included
    {
      const int la = min(max(g_first - 1, 0), 255);
      hrow(la, hA);
    }
    for (int G = g_first; G <= g_last; ++G) {
      const int lb = min(max(G, 0), 255);
      hrow(lb, hB);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rn = 4 * G - 2 + i;                   // row generated in this step (may be outside the image)
        unsigned int Wn = 0;
        if (rn >= 0 && rn < H) {
          const float l1 = 0.125f + 0.25f * i, l0 = 1.0f - l1;
          const bool top = rn < 2;                      // rows 0, 1: weights (1, 0) on low rows (0, 1)
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = top ? lerp_torch(1.0f, hB[k], 0.0f, hB[k]) : lerp_torch(l0, hA[k], l1, hB[k]);
            Wn |= (x > 0.0f) ? (1u << k) : 0u;
          }
          if (rn >= R0 && rn < R1) {
            // np.packbits order (utils/mask_encoding.py:24): pixel 8 i + k of the row-major mask is bit 7 - k of byte i.
            // Bit k of Wn is column 32 lane + k: reverse the bits, then the bytes (little-endian store) -- 128 contiguous
            // bytes per warp and row.
            pout[rn * (W / 32) + lane] = __byte_perm(__brev(Wn), 0u, 0x0123u);
          }
          if (mout && rn >= R0 && rn < R1) {
            // bytes of this row (on request only): lane L writes columns 16 L .. +15 and 512 + 16 L .. +15
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const unsigned int ws = __shfl_sync(0xFFFFFFFFu, Wn, (lane >> 1) + 16 * hh);
              const unsigned int bits = (ws >> (16 * (lane & 1))) & 0xFFFFu;
              uint4 o;
              o.x = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
              o.y = (((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
              o.z = (((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
              o.w = (((bits >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
              *reinterpret_cast<uint4*>(mout + static_cast<size_t>(rn) * W + 512 * hh + 16 * lane) = o;
            }
          }
        }
        // ---- border of row rn-1 (mask rows rn-2, rn-1, rn)
        unsigned int wl = __shfl_up_sync(0xFFFFFFFFu, Wp1, 1), wr = __shfl_down_sync(0xFFFFFFFFu, Wp1, 1);
        if (lane == 0) wl = 0;
        if (lane == 31) wr = 0;
        const unsigned int Bp1 = Wp1 & ~(Wp2 & Wn & shl_in(Wp1, wl) & shr_in(Wp1, wr));
        unsigned int bl = __shfl_up_sync(0xFFFFFFFFu, Bp1, 1), br = __shfl_down_sync(0xFFFFFFFFu, Bp1, 1);
        if (lane == 0) bl = 0;
        if (lane == 31) br = 0;
        const unsigned int BLp1 = shl_in(Bp1, bl), BRp1 = shr_in(Bp1, br);
        // ---- finalise row f = rn-2: mask word Wp2, border rows f-1 (p3), f (p2), f+1 (p1)
        const int f = rn - 2;
        if (f >= R0 && f < R1) {
          const unsigned int Wf = Wp2;
          if (__any_sync(0xFFFFFFFFu, Wf != 0)) {
            const unsigned int n = __popc(Wf);
            area += n;
            sum_r += static_cast<unsigned long long>(n) * f;
            unsigned int si = __popc(Wf & 0xAAAAAAAAu) + 2 * __popc(Wf & 0xCCCCCCCCu) + 4 * __popc(Wf & 0xF0F0F0F0u) +
                              8 * __popc(Wf & 0xFF00FF00u) + 16 * __popc(Wf & 0xFFFF0000u);
            sum_c += static_cast<unsigned long long>(n) * (32 * lane) + si;
            colmask |= Wf;
            mnr = min(mnr, f); mxr = max(mxr, f);     // warp-uniform: some lane of the warp has a pixel in row f
            // perimeter code = 1 + 2 n4 + 10 nd over border neighbours; bit-sliced counts n4, nd in {0..4}
            const unsigned int B = Bp2;
            if (__any_sync(0xFFFFFFFFu, B != 0)) {
              const unsigned int a0 = Bp3, a1 = Bp1, a2 = BLp2, a3 = BRp2;          // up, down, left, right
              const unsigned int s1 = a0 ^ a1 ^ a2, c1 = (a0 & a1) | (a2 & (a0 ^ a1));
              const unsigned int n4_0 = s1 ^ a3, c2 = s1 & a3;
              const unsigned int n4_1 = c1 ^ c2, n4_2 = c1 & c2;
              const unsigned int d0 = BLp3, d1 = BRp3, d2 = BLp1, d3 = BRp1;        // diagonals
              const unsigned int t1 = d0 ^ d1 ^ d2, e1 = (d0 & d1) | (d2 & (d0 ^ d1));
              const unsigned int nd_0 = t1 ^ d3, e2 = t1 & d3;
              const unsigned int nd_1 = e1 ^ e2, nd_2 = e1 & e2;
              // selectors n4 == v / nd == v for the values that occur in weighted codes (5,7,13,15,17,21,23,25,27,33)
              const unsigned int n4e0 = ~(n4_0 | n4_1 | n4_2), n4e1 = n4_0 & ~n4_1 & ~n4_2, n4e2 = ~n4_0 & n4_1 & ~n4_2,
                                 n4e3 = n4_0 & n4_1 & ~n4_2;
              const unsigned int nde0 = ~(nd_0 | nd_1 | nd_2), nde1 = nd_0 & ~nd_1 & ~nd_2, nde2 = ~nd_0 & nd_1 & ~nd_2,
                                 nde3 = nd_0 & nd_1 & ~nd_2;
              pc[0] += __popc(B & n4e2 & nde0);   // code 5
              pc[1] += __popc(B & n4e3 & nde0);   // 7
              pc[2] += __popc(B & n4e1 & nde1);   // 13
              pc[3] += __popc(B & n4e2 & nde1);   // 15
              pc[4] += __popc(B & n4e3 & nde1);   // 17
              pc[5] += __popc(B & n4e0 & nde2);   // 21
              pc[6] += __popc(B & n4e1 & nde2);   // 23
              pc[7] += __popc(B & n4e2 & nde2);   // 25
              pc[8] += __popc(B & n4e3 & nde2);   // 27
              pc[9] += __popc(B & n4e1 & nde3);   // 33
            }
            // intensity histogram over the mask pixels of this row word
            if (want_hist && Wf != 0) {
              const uint4* gp = reinterpret_cast<const uint4*>(gray + static_cast<size_t>(f) * W + 32 * lane);
              const uint4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
              const unsigned int gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (Wf & (1u << k)) atomicAdd(&my_hist[(gw[k >> 2] >> (8 * (k & 3))) & 0xFFu], 1u);
            }
          }
          // first mixed 2x2 cell with (f, c) as its upper-left pixel, c < W-1, f < H-1 (rows f, f+1 = Wp2, Wp1)
          if (first == 0xFFFFFFFFu && f < H - 1) {
            unsigned int nf = __shfl_down_sync(0xFFFFFFFFu, Wf, 1), n1 = __shfl_down_sync(0xFFFFFFFFu, Wp1, 1);
            if (lane == 31) { nf = Wf >> 31 << 0; n1 = Wp1 >> 31; }   // column 1024 does not exist: replicate -> never mixed
            const unsigned int Wfr = shr_in(Wf, lane == 31 ? (Wf >> 31) : nf), W1r = shr_in(Wp1, lane == 31 ? (Wp1 >> 31) : n1);
            unsigned int mixed = (Wf ^ Wfr) | (Wf ^ Wp1) | (Wf ^ W1r);
            if (lane == 31) mixed &= 0x7FFFFFFFu;     // c = 1023 is not an upper-left corner
            const unsigned int ball = __ballot_sync(0xFFFFFFFFu, mixed != 0);
            if (ball) {
              const int l0 = __ffs(ball) - 1;
              const unsigned int mw = __shfl_sync(0xFFFFFFFFu, mixed, l0);
              first = static_cast<unsigned int>(f) * W + 32 * l0 + (__ffs(mw) - 1);
            }
          }
        }
        Wp2 = Wp1; Wp1 = Wn;
        Bp3 = Bp2; Bp2 = Bp1; BLp3 = BLp2; BRp3 = BRp2; BLp2 = BLp1; BRp2 = BRp1;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) hA[k] = hB[k];
    }
  }
  // ---- reduce: lanes -> warp -> global
  int mnc = INT_MAX, mxc = -1;
  if (colmask) { mnc = 32 * lane + __ffs(colmask) - 1; mxc = 32 * lane + 31 - __clz(colmask); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    area += __shfl_xor_sync(0xFFFFFFFFu, area, o);
    sum_r += __shfl_xor_sync(0xFFFFFFFFu, sum_r, o);
    sum_c += __shfl_xor_sync(0xFFFFFFFFu, sum_c, o);
    mnc = min(mnc, __shfl_xor_sync(0xFFFFFFFFu, mnc, o));
    mxc = max(mxc, __shfl_xor_sync(0xFFFFFFFFu, mxc, o));
#pragma unroll
    for (int b = 0; b < YSI_PERIM_BINS; ++b) pc[b] += __shfl_xor_sync(0xFFFFFFFFu, pc[b], o);
  }
  MaskStatsDev* st = stats + m;
  if (lane == 0) {
    if (area) {
      atomicAdd(&st->area, area); atomicAdd(&st->sum_r, sum_r); atomicAdd(&st->sum_c, sum_c);
      atomicMin(&st->min_r, mnr); atomicMin(&st->min_c, mnc); atomicMax(&st->max_r, mxr); atomicMax(&st->max_c, mxc);
    }
    if (first != 0xFFFFFFFFu) atomicMin(&st->first_cell, first);
#pragma unroll
    for (int b = 0; b < YSI_PERIM_BINS; ++b)
      if (pc[b]) atomicAdd(&st->perim_hist[b], pc[b]);
  }
  if (want_hist) {
    __syncthreads();
    for (int bin = threadIdx.x; bin < 256; bin += FAST_WARPS * 32) {
      unsigned int hsum = 0;
#pragma unroll
      for (int w = 0; w < FAST_WARPS; ++w) hsum += s_hist[w][0][bin] + s_hist[w][1][bin];
      if (hsum) atomicAdd(&st->mask_hist[bin], hsum);
    }
  }
}

void launch_upsample_stats(const float* low, int nmask, PostGeom g, const uint16_t* sum3, const uint8_t* gray,
                           const int* mask_image, uint8_t* masks, uint8_t* packed, bool want_bytes, float* up_logits,
                           MaskStatsDev* stats, cudaStream_t s) {
  if (nmask <= 0) return;
  YSI_CHECK(masks && packed, "upsample: mask scratch and packed output are required");
  if (g.H == 1024 && g.W == 1024 && g.rh == 1024 && g.rw == 1024 && !up_logits && (sum3 == nullptr || gray != nullptr)) {
    // enough warps to fill the GPU for few masks, long bands (little halo recomputation) for many
    // 257 row groups per mask. Many masks: 33 groups per warp = 8 warps = 2 CTAs per mask, so 256 masks are one wave of 512 CTAs
    // (592 slots) with 6 % halo recomputation; few masks: short bands, more warps
    const int gpw = nmask >= 64 ? 33 : (nmask >= 16 ? 8 : 4);       // row groups (4 rows each) per warp
    const int ctas = ceil_div(257, FAST_WARPS * gpw);
    upsample_stats_fast_kernel<<<dim3(ctas, nmask), FAST_WARPS * 32, 0, s>>>(low, sum3 ? gray : nullptr, mask_image,
                                                                             want_bytes ? masks : nullptr, packed, stats, gpw);
    YSI_CUDA(cudaGetLastError());
    return;
  }
  PostParams p;
  p.g = g;
  p.sh = static_cast<float>(g.rh) / static_cast<float>(g.H);
  p.sw = static_cast<float>(g.rw) / static_cast<float>(g.W);
  dim3 grid(ceil_div(g.W, TC), ceil_div(g.H, TR), nmask);
  upsample_stats_kernel<0><<<grid, 256, 0, s>>>(low, nullptr, p, sum3, mask_image, masks, up_logits, stats);
  YSI_CUDA(cudaGetLastError());
  // generic geometry: the byte masks are scratch, the packed rows (what the contour kernel and the wire format read)
  // are derived from them
  launch_packbits(masks, packed, nmask, static_cast<long long>(g.H) * g.W, s);
}

void launch_mask_stats(const uint8_t* masks_in, int nmask, int H, int W, const uint16_t* sum3, const int* mask_image,
                       MaskStatsDev* stats, cudaStream_t s) {
  if (nmask <= 0) return;
  PostParams p;
  p.g.H = H; p.g.W = W; p.g.rh = H; p.g.rw = W;
  p.sh = p.sw = 1.0f;
  dim3 grid(ceil_div(W, TC), ceil_div(H, TR), nmask);
  upsample_stats_kernel<1><<<grid, 256, 0, s>>>(nullptr, masks_in, p, sum3, mask_image, nullptr, nullptr, stats);
  YSI_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
// K2: contour 0 -> hull -> hull raster; centre disk
// ---------------------------------------------------------------------------------------------------
// Contour vertices live on doubled coordinates: pixel (r,c) -> (2r,2c); the crossing between two
// 4-adjacent pixels is their midpoint.  A 2x2 cell (r0,c0) has edges T=(2r0,2c0+1) B=(2r0+2,2c0+1)
// L=(2r0+1,2c0) R=(2r0+1,2c0+2).
enum { E_T = 0, E_B = 1, E_L = 2, E_R = 3, E_NONE = 4 };

struct CellWalk {
  const uint8_t* mk;      // np.packbits rows of the mask: pixel i = r W + c is bit 7 - (i & 7) of byte i >> 3
  int H, W;
  __device__ __forceinline__ int px(int r, int c) const {
    const unsigned int i = static_cast<unsigned int>(r) * W + c;
    return (__ldg(mk + (i >> 3)) >> (7u - (i & 7u))) & 1u;       // read-only path: the walk revisits the same few lines
  }
  // marching-squares case of cell (r0,c0): 1*ul + 2*ur + 4*ll + 8*lr   (_find_contours_cy.pyx)
  __device__ __forceinline__ int cell_case(int r0, int c0) const {
    return px(r0, c0) + 2 * px(r0, c0 + 1) + 4 * px(r0 + 1, c0) + 8 * px(r0 + 1, c0 + 1);
  }
};

// the edge paired with `e` by the segment(s) of a cell with marching-squares case `cs` (low connectivity)
__device__ __forceinline__ int paired_edge(int cs, int e) {
  const int ul = cs & 1, ur = (cs >> 1) & 1, ll = (cs >> 2) & 1, lr = (cs >> 3) & 1;
  if (cs == 6) {          // segments (right,top) and (left,bottom)
    return e == E_T ? E_R : e == E_R ? E_T : e == E_L ? E_B : E_L;
  }
  if (cs == 9) {          // segments (top,left) and (bottom,right)
    return e == E_T ? E_L : e == E_L ? E_T : e == E_B ? E_R : E_B;
  }
  const bool ct = ul != ur, cb = ll != lr, cl = ul != ll, cr = ur != lr;
  if (ct && e != E_T) return E_T;
  if (cb && e != E_B) return E_B;
  if (cl && e != E_L) return E_L;
  if (cr && e != E_R) return E_R;
  return E_NONE;
}

// first segment emitted by a mixed cell, as (from_edge, to_edge), per the case table of _find_contours_cy.pyx
__device__ __forceinline__ void first_segment(int cs, int& e_from, int& e_to) {
  switch (cs) {
    case 1: e_from = E_T; e_to = E_L; break;
    case 2: e_from = E_R; e_to = E_T; break;
    case 3: e_from = E_R; e_to = E_L; break;
    case 4: e_from = E_L; e_to = E_B; break;
    case 5: e_from = E_T; e_to = E_B; break;
    case 6: e_from = E_R; e_to = E_T; break;
    case 7: e_from = E_R; e_to = E_B; break;
    case 8: e_from = E_B; e_to = E_R; break;
    case 9: e_from = E_T; e_to = E_L; break;
    case 10: e_from = E_B; e_to = E_T; break;
    case 11: e_from = E_B; e_to = E_L; break;
    case 12: e_from = E_L; e_to = E_R; break;
    case 13: e_from = E_T; e_to = E_R; break;
    default: e_from = E_L; e_to = E_T; break;   // 14
  }
}

__device__ __forceinline__ void edge_point(int r0, int c0, int e, int& y2, int& x2) {
  y2 = 2 * r0 + (e == E_T ? 0 : e == E_B ? 2 : 1);
  x2 = 2 * c0 + (e == E_L ? 0 : e == E_R ? 2 : 1);
}

struct RowSpan {   // per doubled row: min / max doubled column among contour vertices
  int2* span;      // [2H+1]
  int ymin, ymax, npts;
  __device__ __forceinline__ void add(int y2, int x2) {
    int2 s = span[y2];
    if (x2 < s.x) s.x = x2;
    if (x2 > s.y) s.y = x2;
    span[y2] = s;
    ymin = min(ymin, y2);
    ymax = max(ymax, y2);
    ++npts;
  }
};

// walk from the vertex on edge `e` of cell (r0,c0) away from that cell until the contour closes at
// (stop_y, stop_x) or leaves the cell grid.  Returns true if it closed.
__device__ bool walk_contour(const CellWalk& cw, int r0, int c0, int e, int stop_y, int stop_x, RowSpan& rs,
                             long long budget, bool& truncated) {
  while (budget-- > 0) {
    // cross the edge into the neighbouring cell
    int nr = r0, nc = c0, ne;
    if (e == E_T) { nr = r0 - 1; ne = E_B; }
    else if (e == E_B) { nr = r0 + 1; ne = E_T; }
    else if (e == E_L) { nc = c0 - 1; ne = E_R; }
    else { nc = c0 + 1; ne = E_L; }
    if (nr < 0 || nc < 0 || nr > cw.H - 2 || nc > cw.W - 2) return false;   // open end at the image border
    const int cs = cw.cell_case(nr, nc);
    const int ex = paired_edge(cs, ne);
    int y2, x2;
    edge_point(nr, nc, ex, y2, x2);
    if (y2 == stop_y && x2 == stop_x) return true;
    rs.add(y2, x2);
    r0 = nr; c0 = nc; e = ex;
  }
  truncated = true;
  return false;
}

constexpr int K2_THREADS = 256;

__global__ void __launch_bounds__(K2_THREADS)
contour_hull_disk_kernel(const uint8_t* __restrict__ packed, int nmask, int H, int W,
                         const uint16_t* __restrict__ sum3_all, const int* __restrict__ mask_image,
                         const MaskStatsDev* __restrict__ stats, ysi_mask_metrics* __restrict__ out) {
  const int m = blockIdx.x;
  const int tid = threadIdx.x;
  const size_t plane = static_cast<size_t>(H) * W;
  const MaskStatsDev* stp = stats + m;
  struct { unsigned long long area, sum_r, sum_c; int min_r, min_c, max_r, max_c; unsigned int first_cell; } st;
  st.area = stp->area; st.sum_r = stp->sum_r; st.sum_c = stp->sum_c;
  st.min_r = stp->min_r; st.min_c = stp->min_c; st.max_r = stp->max_r; st.max_c = stp->max_c;
  st.first_cell = stp->first_cell;
  ysi_mask_metrics* o = out + m;
  // dynamic smem: row spans int2[2H+1] | left chain (y,x)[cap] | right chain (y,x)[cap] | lo/hi per pixel row
  extern __shared__ int smem_all[];
  int2* span = reinterpret_cast<int2*>(smem_all);
  int* smem_dyn = smem_all + 2 * (2 * H + 1);
  __shared__ int s_info[8];       // ymin, ymax, npts, closed, truncated, nleft, nright, degenerate
  __shared__ unsigned long long s_disk[3];
  __shared__ unsigned long long s_harea;
  __shared__ unsigned int s_hperim[YSI_PERIM_BINS];

  if (tid < 3) s_disk[tid] = 0;
  if (tid == 0) s_harea = 0;
  if (tid < YSI_PERIM_BINS) s_hperim[tid] = 0;
  for (int i = tid; i < 2 * H + 1; i += K2_THREADS) span[i] = make_int2(INT_MAX, INT_MIN);
  __syncthreads();

  const bool empty = st.area == 0;
  // ---- phase 1a (thread 0): trace contours[0]
  if (tid == 0) {
    RowSpan rs;
    rs.span = span; rs.ymin = INT_MAX; rs.ymax = INT_MIN; rs.npts = 0;
    bool truncated = false;
    if (st.first_cell != 0xFFFFFFFFu) {
      CellWalk cw; cw.mk = packed + static_cast<size_t>(m) * ((plane + 7) / 8); cw.H = H; cw.W = W;
      const int r0 = st.first_cell / W, c0 = st.first_cell % W;
      const int cs = cw.cell_case(r0, c0);
      int ef, et;
      first_segment(cs, ef, et);
      int fy, fx, ty, tx;
      edge_point(r0, c0, ef, fy, fx);
      edge_point(r0, c0, et, ty, tx);
      rs.add(fy, fx);
      rs.add(ty, tx);
      const long long budget = 4ll * H * W;
      const bool closed = walk_contour(cw, r0, c0, et, fy, fx, rs, budget, truncated);
      if (!closed) walk_contour(cw, r0, c0, ef, ty, tx, rs, budget, truncated);
    }
    s_info[0] = rs.ymin; s_info[1] = rs.ymax; s_info[2] = rs.npts; s_info[4] = truncated ? 1 : 0;
  }
  // ---- phase 1b (other warps): centre-disk sums, utils/metrics.py:81-94
  if (tid >= 32 && !empty) {
    const uint16_t* sum3 = sum3_all + static_cast<size_t>(mask_image ? mask_image[m] : 0) * plane;
    const double cx = static_cast<double>(st.sum_r) / static_cast<double>(st.area);   // centroid row
    const double cy = static_cast<double>(st.sum_c) / static_cast<double>(st.area);   // centroid col
    const int radius = static_cast<int>((H < W ? H : W) * 0.1);
    const double r2 = static_cast<double>(radius) * radius;
    int rlo = static_cast<int>(floor(cx)) - radius - 1, rhi = static_cast<int>(ceil(cx)) + radius + 1;
    int clo = static_cast<int>(floor(cy)) - radius - 1, chi = static_cast<int>(ceil(cy)) + radius + 1;
    rlo = max(rlo, 0); clo = max(clo, 0); rhi = min(rhi, H - 1); chi = min(chi, W - 1);
    const int bw = chi - clo + 1, bh = rhi - rlo + 1;
    unsigned long long n = 0, s1 = 0, s2 = 0;
    for (int i = tid - 32; i < bw * bh; i += K2_THREADS - 32) {
      const int r = rlo + i / bw, c = clo + i % bw;
      const double dr = __dsub_rn(static_cast<double>(r), cx), dc = __dsub_rn(static_cast<double>(c), cy);
      const double d2 = __dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc));
      if (d2 <= r2) {
        const unsigned long long v = sum3[static_cast<size_t>(r) * W + c];
        n += 1; s1 += v; s2 += v * v;
      }
    }
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      n += __shfl_xor_sync(0xFFFFFFFFu, n, ofs);
      s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, ofs);
      s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, ofs);
    }
    if ((tid & 31) == 0) { atomicAdd(&s_disk[0], n); atomicAdd(&s_disk[1], s1); atomicAdd(&s_disk[2], s2); }
  }
  __syncthreads();

  // ---- phase 2 (thread 0): monotone-chain hull over (row, min col) / (row, max col)
  const int ymin = s_info[0], ymax = s_info[1], npts = s_info[2];
  const int cap = (ymax >= ymin) ? (ymax - ymin + 1) : 0;
  int* lch = smem_dyn;                 // left chain  : y at [2i], x at [2i+1]
  int* rch = smem_dyn + 2 * cap;       // right chain
  if (tid == 0) {
    int nl = 0, nr = 0;
    for (int y = ymin; y <= ymax && npts > 0; ++y) {
      const int2 sp = span[y];
      if (sp.x > sp.y) continue;       // no vertex on this doubled row
      // left chain keeps vertices strictly left of the chord; pop while the middle one is on/right of it
      while (nl >= 2) {
        const long long ay = lch[2 * (nl - 2)], ax = lch[2 * (nl - 2) + 1];
        const long long by = lch[2 * (nl - 1)], bx = lch[2 * (nl - 1) + 1];
        if ((bx - ax) * (y - ay) - (sp.x - ax) * (by - ay) >= 0) --nl; else break;
      }
      lch[2 * nl] = y; lch[2 * nl + 1] = sp.x; ++nl;
      while (nr >= 2) {
        const long long ay = rch[2 * (nr - 2)], ax = rch[2 * (nr - 2) + 1];
        const long long by = rch[2 * (nr - 1)], bx = rch[2 * (nr - 1) + 1];
        if ((bx - ax) * (y - ay) - (sp.y - ax) * (by - ay) <= 0) --nr; else break;
      }
      rch[2 * nr] = y; rch[2 * nr + 1] = sp.y; ++nr;
    }
    // twice the polygon area (shoelace) decides degeneracy (< 3 non-collinear points => QhullError)
    long long area2 = 0;
    for (int i = 0; i + 1 < nl; ++i) {   // down the left chain
      area2 += static_cast<long long>(lch[2 * i + 1]) * lch[2 * (i + 1)] - static_cast<long long>(lch[2 * (i + 1) + 1]) * lch[2 * i];
    }
    if (nl > 0 && nr > 0)                // bottom: left end -> right end
      area2 += static_cast<long long>(lch[2 * (nl - 1) + 1]) * rch[2 * (nr - 1)] - static_cast<long long>(rch[2 * (nr - 1) + 1]) * lch[2 * (nl - 1)];
    for (int i = nr - 1; i > 0; --i) {   // up the right chain
      area2 += static_cast<long long>(rch[2 * i + 1]) * rch[2 * (i - 1)] - static_cast<long long>(rch[2 * (i - 1) + 1]) * rch[2 * i];
    }
    if (nl > 0 && nr > 0)                // top: right start -> left start
      area2 += static_cast<long long>(rch[1]) * lch[0] - static_cast<long long>(lch[1]) * rch[0];
    s_info[5] = nl; s_info[6] = nr;
    s_info[7] = (npts < 3 || area2 == 0) ? 1 : 0;
  }
  __syncthreads();
  const int nl = s_info[5], nr = s_info[6];
  const bool degenerate = s_info[7] != 0;

  // ---- phase 3: inclusive raster of the hull: per pixel row r the column interval [lo, hi]
  // pixel rows covered: ceil(ymin/2) .. floor(ymax/2); intervals padded by 2 empty rows on both sides
  const int pr0 = (ymin + 1) >> 1, pr1 = ymax >> 1;
  const int nrows = (!degenerate && pr1 >= pr0) ? (pr1 - pr0 + 1) : 0;
  int* lo = smem_dyn + 4 * cap;                 // [nrows + 4]
  int* hi = lo + (nrows + 4);
  for (int i = tid; i < nrows + 4; i += K2_THREADS) {
    int l = 1, h = 0;   // empty
    const int r = pr0 - 2 + i;
    if (i >= 2 && i < nrows + 2) {
      const int y = 2 * r;
      // left boundary: segment of the left chain containing y
      int k = 0;
      while (k + 1 < nl && lch[2 * (k + 1)] < y) ++k;
      {
        const long long ay = lch[2 * k], ax = lch[2 * k + 1];
        if (k + 1 < nl && ay != y) {
          const long long by = lch[2 * (k + 1)], bx = lch[2 * (k + 1) + 1];
          // x(y) = ax + (bx-ax)(y-ay)/(by-ay);  need 2c >= x  =>  c >= num / (2 den), ceil
          const long long den = 2 * (by - ay), num = ax * (by - ay) + (bx - ax) * (y - ay);
          l = static_cast<int>(num >= 0 ? (num + den - 1) / den : -((-num) / den));
        } else {
          l = static_cast<int>((ax + 1) >> 1);   // ceil(ax / 2), ax >= 0
        }
      }
      k = 0;
      while (k + 1 < nr && rch[2 * (k + 1)] < y) ++k;
      {
        const long long ay = rch[2 * k], ax = rch[2 * k + 1];
        if (k + 1 < nr && ay != y) {
          const long long by = rch[2 * (k + 1)], bx = rch[2 * (k + 1) + 1];
          const long long den = 2 * (by - ay), num = ax * (by - ay) + (bx - ax) * (y - ay);
          h = static_cast<int>(num >= 0 ? num / den : -((-num + den - 1) / den));   // floor
        } else {
          h = static_cast<int>(ax >> 1);         // floor(ax / 2)
        }
      }
      if (l < 0) l = 0;
      if (h > W - 1) h = W - 1;
    }
    lo[i] = l; hi[i] = h;
  }
  __syncthreads();
  if (nrows > 0) {
    // hull area + perimeter histogram (border pixels only are classified)
    auto in_hull = [&](int i, int c) -> int { return (c >= lo[i] && c <= hi[i]) ? 1 : 0; };   // i = row index + 2
    auto is_border = [&](int i, int c) -> int {
      if (!in_hull(i, c)) return 0;
      return (in_hull(i - 1, c) & in_hull(i + 1, c) & in_hull(i, c - 1) & in_hull(i, c + 1)) ? 0 : 1;
    };
    unsigned long long harea = 0;
    unsigned int pc[YSI_PERIM_BINS];
#pragma unroll
    for (int b = 0; b < YSI_PERIM_BINS; ++b) pc[b] = 0;
    // work items: (row, column run). Interior pixels of a row are skipped with an interval test.
    for (int i = 2 + (tid >> 5); i < nrows + 2; i += K2_THREADS / 32) {
      const int l = lo[i], h = hi[i];
      if (tid % 32 == 0 && h >= l) harea += static_cast<unsigned long long>(h - l + 1);
      // interior columns: strictly inside this row's interval and inside both neighbour rows' intervals
      const int il = max(l + 1, max(lo[i - 1], lo[i + 1])), ih = min(h - 1, min(hi[i - 1], hi[i + 1]));
      for (int c = l + (tid & 31); c <= h; c += 32) {
        if (c >= il && c <= ih) {           // jump over the interior run
          continue;
        }
        if (!is_border(i, c)) continue;
        const int n4 = is_border(i - 1, c) + is_border(i + 1, c) + is_border(i, c - 1) + is_border(i, c + 1);
        const int nd = is_border(i - 1, c - 1) + is_border(i - 1, c + 1) + is_border(i + 1, c - 1) + is_border(i + 1, c + 1);
        const int bin = c_code_bin[1 + 2 * n4 + 10 * nd];
        if (bin >= 0) pc[bin] += 1;
      }
    }
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
      harea += __shfl_xor_sync(0xFFFFFFFFu, harea, ofs);
#pragma unroll
      for (int b = 0; b < YSI_PERIM_BINS; ++b) pc[b] += __shfl_xor_sync(0xFFFFFFFFu, pc[b], ofs);
    }
    if ((tid & 31) == 0) {
      atomicAdd(&s_harea, harea);
#pragma unroll
      for (int b = 0; b < YSI_PERIM_BINS; ++b)
        if (pc[b]) atomicAdd(&s_hperim[b], pc[b]);
    }
  }
  __syncthreads();
  // ---- final row
  if (tid == 0) {
    o->area = static_cast<int64_t>(st.area);
    o->sum_r = static_cast<int64_t>(st.sum_r);
    o->sum_c = static_cast<int64_t>(st.sum_c);
    o->min_r = empty ? 0 : st.min_r; o->min_c = empty ? 0 : st.min_c;
    o->max_r = empty ? 0 : st.max_r + 1; o->max_c = empty ? 0 : st.max_c + 1;
    o->hull_area = static_cast<int64_t>(s_harea);
    o->disk_n = static_cast<int64_t>(s_disk[0]);
    o->disk_sum = s_disk[1];
    o->disk_sumsq = s_disk[2];
    // an empty hull raster also lands in the reference's except branch (regionprops(...)[0] raises)
    o->flags = (empty ? YSI_FLAG_EMPTY_MASK : 0u) | ((degenerate || s_harea == 0) ? YSI_FLAG_HULL_DEGENERATE : 0u) |
               (s_info[4] ? YSI_FLAG_CONTOUR_TRUNCATED : 0u);
    o->contour_points = npts;
    o->hull_vertices = degenerate ? 0 : nl + nr;
    o->reserved = 0;
  }
  if (tid < YSI_PERIM_BINS) {
    o->perim_hist[tid] = stp->perim_hist[tid];
    o->hull_perim_hist[tid] = s_hperim[tid];
  }
  o->mask_hist[tid] = stp->mask_hist[tid];
}

void launch_contour_hull_disk(const uint8_t* packed, int nmask, int H, int W, const uint16_t* sum3,
                              const int* mask_image, const MaskStatsDev* stats, ysi_mask_metrics* out,
                              cudaStream_t s) {
  if (nmask <= 0) return;
  // dynamic smem: row spans + two chains of (2H+1) (y,x) pairs + lo/hi for (H+4) rows
  const size_t smem = sizeof(int) * (6 * static_cast<size_t>(2 * H + 1) + 2 * static_cast<size_t>(H + 4));
  YSI_CHECK(smem <= 226 * 1024, "image too tall for the hull kernel's shared memory (H <= 4096)");
  ensure_dyn_smem(reinterpret_cast<const void*>(contour_hull_disk_kernel), static_cast<int>(smem));
  contour_hull_disk_kernel<<<nmask, K2_THREADS, smem, s>>>(packed, nmask, H, W, sum3, mask_image, stats, out);
  YSI_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
__global__ void packbits_kernel(const uint8_t* __restrict__ masks, uint8_t* __restrict__ packed, long long npix,
                                long long nbytes) {
  const int m = blockIdx.y;
  const uint8_t* src = masks + m * npix;
  uint8_t* dst = packed + m * nbytes;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nbytes;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p0 = i * 8;
    unsigned int v = 0;
    if (p0 + 8 <= npix && ((reinterpret_cast<uintptr_t>(src + p0) & 7) == 0)) {
      const unsigned long long w = *reinterpret_cast<const unsigned long long*>(src + p0);
#pragma unroll
      for (int k = 0; k < 8; ++k) v |= (((w >> (8 * k)) & 0xFF) ? 1u : 0u) << (7 - k);
    } else {
      for (int k = 0; k < 8 && p0 + k < npix; ++k) v |= (src[p0 + k] ? 1u : 0u) << (7 - k);
    }
    dst[i] = static_cast<uint8_t>(v);
  }
}

void launch_packbits(const uint8_t* masks, uint8_t* packed, int nmask, long long npix, cudaStream_t s) {
  if (nmask <= 0) return;
  const long long nbytes = (npix + 7) / 8;
  dim3 grid(static_cast<unsigned>(std::min<long long>((nbytes + 255) / 256, 1184)), nmask);
  packbits_kernel<<<grid, 256, 0, s>>>(masks, packed, npix, nbytes);
  YSI_CUDA(cudaGetLastError());
}

// R+G+B per pixel (uint16, the centre-disk sums square it) and floor((R+G+B)/3) (uint8, the mask histogram's bin)
__global__ void sum3_kernel(const uint8_t* __restrict__ rgb, int H, int W, int row_stride, uint16_t* __restrict__ out,
                            uint8_t* __restrict__ gray) {
  const int n = blockIdx.z;
  const int r = blockIdx.y;
  const uint8_t* row = rgb + (static_cast<size_t>(n) * H + r) * row_stride;
  uint16_t* orow = out + (static_cast<size_t>(n) * H + r) * W;
  uint8_t* grow = gray ? gray + (static_cast<size_t>(n) * H + r) * W : nullptr;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < W; c += gridDim.x * blockDim.x) {
    const unsigned int v = row[3 * c] + row[3 * c + 1] + row[3 * c + 2];
    orow[c] = static_cast<uint16_t>(v);
    if (grow) grow[c] = static_cast<uint8_t>(v / 3);
  }
}

void launch_sum3(const uint8_t* rgb, int n, int H, int W, int row_stride, uint16_t* sum3, uint8_t* gray, cudaStream_t s) {
  dim3 grid(ceil_div(W, 256), H, n);
  sum3_kernel<<<grid, 256, 0, s>>>(rgb, H, W, row_stride, sum3, gray);
  YSI_CUDA(cudaGetLastError());
}

// f3: raw grey pixels (uint8, or uint16 -> v >> 8 as cv2.imread's default flags reduce 16-bit files) -> uint8 RGB (grey
// replicated) + R+G+B (uint16) + floor((R+G+B)/3) = grey (uint8). One thread = 4 consecutive pixels of the flat array.
template <int BPP>
__global__ void __launch_bounds__(256)
gray_ingest_kernel(const uint8_t* __restrict__ src, long long npix, uint8_t* __restrict__ rgb, uint16_t* __restrict__ sum3,
                   uint8_t* __restrict__ gray) {
  const long long ngrp = (npix + 3) / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < ngrp;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p0 = i * 4;
    unsigned int g[4];
    if (p0 + 4 <= npix) {
      if (BPP == 1) {
        const unsigned int w = *reinterpret_cast<const unsigned int*>(src + p0);
        g[0] = w & 0xFFu; g[1] = (w >> 8) & 0xFFu; g[2] = (w >> 16) & 0xFFu; g[3] = w >> 24;
      } else {
        const uint2 w = *reinterpret_cast<const uint2*>(src + 2 * p0);      // little-endian uint16: high byte = v >> 8
        g[0] = (w.x >> 8) & 0xFFu; g[1] = w.x >> 24; g[2] = (w.y >> 8) & 0xFFu; g[3] = w.y >> 24;
      }
      // 12 RGB bytes: g0 g0 g0 g1 | g1 g1 g2 g2 | g2 g3 g3 g3
      unsigned int* o = reinterpret_cast<unsigned int*>(rgb + 3 * p0);
      o[0] = g[0] * 0x010101u | (g[1] << 24);
      o[1] = g[1] * 0x0101u | (g[2] << 16) | (g[2] << 24);
      o[2] = g[2] | (g[3] << 8) | (g[3] << 16) | (g[3] << 24);
      *reinterpret_cast<uint2*>(sum3 + p0) = make_uint2((3u * g[0]) | ((3u * g[1]) << 16), (3u * g[2]) | ((3u * g[3]) << 16));
      *reinterpret_cast<unsigned int*>(gray + p0) = g[0] | (g[1] << 8) | (g[2] << 16) | (g[3] << 24);
    } else {
      for (long long p = p0; p < npix; ++p) {
        const unsigned int v = BPP == 1 ? src[p] : src[2 * p + 1];
        rgb[3 * p] = rgb[3 * p + 1] = rgb[3 * p + 2] = static_cast<uint8_t>(v);
        sum3[p] = static_cast<uint16_t>(3u * v);
        gray[p] = static_cast<uint8_t>(v);
      }
    }
  }
}

void launch_gray_ingest(const void* src, int bytes_per_px, int n, int H, int W, uint8_t* rgb, uint16_t* sum3, uint8_t* gray,
                        cudaStream_t s) {
  YSI_CHECK(bytes_per_px == 1 || bytes_per_px == 2, "grey ingest takes 8- or 16-bit pixels");
  const long long npix = static_cast<long long>(n) * H * W;
  const int grid = static_cast<int>(std::min<long long>((npix / 4 + 255) / 256, 148 * 16));
  if (bytes_per_px == 1) gray_ingest_kernel<1><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(src), npix, rgb, sum3, gray);
  else gray_ingest_kernel<2><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(src), npix, rgb, sum3, gray);
  YSI_CUDA(cudaGetLastError());
}

}  // namespace ysi
