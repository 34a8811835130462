// Fused flash-style attention for the SAM ViT encoder (modeling_sam.py:761-801, 843-882), sm_100a.
//
//   out = softmax(q k^T * hd^-0.5 + q.Rh[qh-kh] + q.Rw[qw-kw]) v      (decomposed rel-pos bias, unscaled q)
//
// One CTA = 128 queries of one (sequence, head); keys/values stream through in tiles of 64.
//   warp 4 (one lane) : TMA loads (Q, rel-pos tables, K/V double buffer) + all tcgen05.mma issue
//   warps 0-3         : one query row per thread: S from TMEM, bias, online softmax (fp32, exp2),
//                       P -> bf16 -> 128B-swizzled smem, O accumulated in registers from TMEM partials
// Tensor-core work per tile: S = Q K^T (128x64x64, both K-major) and PV = P V (128x64x64, V is the
// MN-major B operand straight out of the qkv activation, no transpose pass). The rel-pos terms are
// two extra 128x128x64 MMAs per CTA (Q . table^T) whose TMEM result is re-indexed per query:
//   global   (S=64): rel_h stays in TMEM (column qh-kh+63 is warp-uniform), rel_w goes to smem [kw][q]
//   windowed (S=14): both go to smem [k][q]; 64->70 zero-padded tokens are ordinary keys (they carry
//                    the qkv bias, modeling_sam.py:913-916), keys >= 196 of the tile are masked.
// 2 CTAs/SM (96 KB smem, 256 TMEM columns each) overlap one CTA's softmax with the other's MMAs.
#include "kernels.h"
#include "ptx.cuh"

namespace ysi {

namespace attn {
constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int THREADS = 160;
constexpr int Q_BYTES = BQ * HD * 2;        // 16 KB
constexpr int KV_BYTES = BKV * HD * 2;      // 8 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;            // 2 stages (also the rel_w table during setup)
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;       // 2 stages
constexpr int OFF_P = OFF_V + 2 * KV_BYTES;       // 16 KB (also the rel_h table during setup)
constexpr int OFF_REL = OFF_P + Q_BYTES;          // fp32 bias tables [k][128]
constexpr int REL_BYTES = 64 * 128 * 4;           // 32 KB (global: rel_w; windowed: rel_h | rel_w, 14 rows each)
constexpr int OFF_BAR = OFF_REL + REL_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // + alignment slack
constexpr int TMEM_COLS = 256;
constexpr int COL_TH = 0, COL_S = 128, COL_PV = 192, COL_TW = 128;
}  // namespace attn

__constant__ unsigned char c_div14[256];

struct AttnParams {
  int T;          // sequence length: 196 (window) or 4096 (global)
  int D;          // heads * 64
  float scale_log2e;   // hd^-0.5 * log2(e)
  bf16* out;      // [n_seq * T, D]
};

template <bool GLOBAL>
__global__ void __launch_bounds__(attn::THREADS, 2)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const __grid_constant__ CUtensorMap tmRel, AttnParams p) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  float* rel_s = reinterpret_cast<float*>(sgen + OFF_REL);
  const uint32_t bar = sbase + OFF_BAR;
  const uint32_t bar_q = bar, bar_tab = bar + 8, bar_rel = bar + 16, bar_s = bar + 24, bar_p = bar + 32,
                 bar_o = bar + 40, bar_kvfull0 = bar + 48, bar_kvempty0 = bar + 64;
  const uint32_t tmem_ptr_smem = bar + 80;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int row0 = seq * p.T;                 // first row of this sequence in the qkv matrix
  const int ntiles = (p.T + BKV - 1) / BKV;
  const int cq = head * HD, ck = p.D + head * HD, cv = 2 * p.D + head * HD;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_tab, 1); mbar_init(bar_rel, 128); mbar_init(bar_s, 1);
    mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    mbar_init(bar_kvfull0, 1); mbar_init(bar_kvfull0 + 8, 1);
    mbar_init(bar_kvempty0, 1); mbar_init(bar_kvempty0 + 8, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc_n128 = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);   // B (= V) is MN-major
      // ---- setup: Q tile + both rel-pos tables
      mbar_arrive_expect_tx(bar_q, 3 * Q_BYTES);
      tma_load_2d(sbase + OFF_Q, &tmQ, bar_q, cq, row0 + qt * BQ);
      tma_load_2d(sbase + OFF_P, &tmRel, bar_q, 0, 0);      // rel_pos_h, 128 rows (zero padded)
      tma_load_2d(sbase + OFF_K, &tmRel, bar_q, 0, 128);    // rel_pos_w
      mbar_wait(bar_q, 0);
      tc_fence_after();
      const uint64_t qdesc = umma_desc_sw128(sbase + OFF_Q, 16, 1024);
      {
        const uint64_t hdesc = umma_desc_sw128(sbase + OFF_P, 16, 1024);
        const uint64_t wdesc = umma_desc_sw128(sbase + OFF_K, 16, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + COL_TH, qdesc + 2u * k, hdesc + 2u * k, idesc_n128, k);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + COL_TW, qdesc + 2u * k, wdesc + 2u * k, idesc_n128, k);
        umma_commit(bar_tab);
      }
      mbar_wait(bar_rel, 0);     // tables copied out of TMEM; K stages and S/PV columns are free
      tc_fence_after();
      // ---- K/V prologue
      for (int j = 0; j < 2 && j < ntiles; ++j) {
        mbar_arrive_expect_tx(bar_kvfull0 + 8 * j, 2 * KV_BYTES);
        tma_load_2d(sbase + OFF_K + j * KV_BYTES, &tmKV, bar_kvfull0 + 8 * j, ck, row0 + j * BKV);
        tma_load_2d(sbase + OFF_V + j * KV_BYTES, &tmKV, bar_kvfull0 + 8 * j, cv, row0 + j * BKV);
      }
      mbar_wait(bar_kvfull0, 0);
      tc_fence_after();
      {
        const uint64_t kdesc = umma_desc_sw128(sbase + OFF_K, 16, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + COL_S, qdesc + 2u * k, kdesc + 2u * k, idesc_s, k);
        umma_commit(bar_s);
      }
      const uint64_t pdesc = umma_desc_sw128(sbase + OFF_P, 16, 1024);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar_p, j & 1);
        tc_fence_after();
        // PV_j : A = P (K-major, K = 64 keys), B = V_j (MN-major: 64 key rows x 64 hd); 16 keys = 2048 B
        const uint64_t vdesc = umma_desc_sw128(sbase + OFF_V + st * KV_BYTES, 1024, 1024);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) umma_bf16_ss(tmem_base + COL_PV, pdesc + 2u * k, vdesc + 128u * k, idesc_pv, k);
        umma_commit(bar_o);
        umma_commit(bar_kvempty0 + 8 * st);
        if (j + 1 < ntiles) {
          const int sn = (j + 1) & 1;
          mbar_wait(bar_kvfull0 + 8 * sn, ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t kdesc = umma_desc_sw128(sbase + OFF_K + sn * KV_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + COL_S, qdesc + 2u * k, kdesc + 2u * k, idesc_s, k);
          umma_commit(bar_s);
        }
        if (j + 2 < ntiles) {
          mbar_wait(bar_kvempty0 + 8 * st, (j >> 1) & 1);
          mbar_arrive_expect_tx(bar_kvfull0 + 8 * st, 2 * KV_BYTES);
          tma_load_2d(sbase + OFF_K + st * KV_BYTES, &tmKV, bar_kvfull0 + 8 * st, ck, row0 + (j + 2) * BKV);
          tma_load_2d(sbase + OFF_V + st * KV_BYTES, &tmKV, bar_kvfull0 + 8 * st, cv, row0 + (j + 2) * BKV);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread = query row
    const int t = threadIdx.x;                       // 0..127, TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int lq = qt * BQ + t;                      // query index inside the sequence
    const bool q_valid = lq < p.T;
    int qh, qw;
    if (GLOBAL) { qh = lq >> 6; qw = lq & 63; } else { qh = (lq < 196 ? lq : 195) / 14; qw = (lq < 196 ? lq : 195) - qh * 14; }
    constexpr int S = GLOBAL ? 64 : 14;

    mbar_wait(bar_tab, 0);
    tc_fence_after();
    // rel_w (and rel_h when windowed) -> smem [k][q], picking column (q_pos - k_pos + S-1) of Q.table^T
    {
      const int ncol = 2 * S - 1;
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(tlane + COL_TW + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kw = qw + (S - 1) - (c0 + i);
          if (kw >= 0 && kw < S) rel_s[(GLOBAL ? 0 : 14 * 128) + kw * 128 + t] = __uint_as_float(r[i]);
        }
      }
      if (!GLOBAL) {
        uint32_t r[32];
        tmem_ld_x32(tlane + COL_TH, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kh = qh + (S - 1) - i;
          if (kh >= 0 && kh < S) rel_s[kh * 128 + t] = __uint_as_float(r[i]);
        }
      }
    }
    tc_fence_before();
    mbar_arrive(bar_rel);

    const float LOG2E = 1.4426950408889634f;
    float m_run = -1.0e30f, l_run = 0.0f;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.0f;
    const uint32_t p_row = sbase + OFF_P + static_cast<uint32_t>(t) * 128u;
    const uint32_t swz = static_cast<uint32_t>(t & 7);

    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      float x[BKV];
      {
        uint32_t r0[32], r1[32];
        tmem_ld_x32(tlane + COL_S, r0);
        tmem_ld_x32(tlane + COL_S + 32, r1);
        float bh = 0.0f;
        if (GLOBAL) {
          uint32_t rb;
          tmem_ld_x1(tlane + COL_TH + static_cast<uint32_t>(qh + 63 - j), rb);   // warp-uniform column
          tmem_ld_wait();
          bh = __uint_as_float(rb) * LOG2E;
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int i = 0; i < BKV; ++i) {
          const float s = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]);
          if (GLOBAL) {
            x[i] = fmaf(s, p.scale_log2e, fmaf(rel_s[i * 128 + t], LOG2E, bh));
          } else {
            const int k = j * BKV + i;
            if (k < 196) {
              const int kh = c_div14[k], kw = k - 14 * kh;
              x[i] = fmaf(s, p.scale_log2e, (rel_s[kh * 128 + t] + rel_s[14 * 128 + kw * 128 + t]) * LOG2E);
            } else {
              x[i] = -INFINITY;
            }
          }
        }
      }
      float tmax = x[0];
#pragma unroll
      for (int i = 1; i < BKV; ++i) tmax = fmaxf(tmax, x[i]);
      const float m_new = fmaxf(m_run, tmax);
      const float alpha = ex2_approx(m_run - m_new);
      float lsum = 0.0f;
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int i = 0; i < BKV; i += 2) {
        const float p0 = ex2_approx(x[i] - m_new), p1 = ex2_approx(x[i + 1] - m_new);
        lsum += p0 + p1;
        pk[i / 2] = pack_bf16x2(p0, p1);
      }
      l_run = fmaf(l_run, alpha, lsum);
      m_run = m_new;
      if (j > 0) {
        // PV_{j-1} (relative to the previous max) is complete; fold it in, then rescale to the new max
        mbar_wait(bar_o, (j - 1) & 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_x32(tlane + COL_PV, r0);
        tmem_ld_x32(tlane + COL_PV + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[i] = (o[i] + __uint_as_float(r0[i])) * alpha;
          o[i + 32] = (o[i + 32] + __uint_as_float(r1[i])) * alpha;
        }
      }
      // P_j -> smem (K-major, 128B swizzle: 16-byte chunk c of row t lands at chunk c ^ (t & 7))
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t addr = p_row + ((static_cast<uint32_t>(c) ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                     "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (ntiles - 1) & 1);
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld_x32(tlane + COL_PV, r0);
      tmem_ld_x32(tlane + COL_PV + 32, r1);
      tmem_ld_wait();
      const float inv = 1.0f / l_run;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = (o[i] + __uint_as_float(r0[i])) * inv;
        o[i + 32] = (o[i + 32] + __uint_as_float(r1[i])) * inv;
      }
    }
    if (q_valid) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row0 + lq) * p.D + head * HD);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        v.x = pack_bf16x2(o[8 * c], o[8 * c + 1]);
        v.y = pack_bf16x2(o[8 * c + 2], o[8 * c + 3]);
        v.z = pack_bf16x2(o[8 * c + 4], o[8 * c + 5]);
        v.w = pack_bf16x2(o[8 * c + 6], o[8 * c + 7]);
        dst[c] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, attn::TMEM_COLS);
}

// qkv: bf16 [n_seq*T, 3D]; rel_tab: bf16 [256, 64] (rows 0..127 rel_pos_h zero-padded, 128..255 rel_pos_w)
void launch_encoder_attention(const bf16* qkv, const bf16* rel_tab, bf16* out, int n_seq, int T, int heads,
                              bool is_global, cudaStream_t stream) {
  using namespace attn;
  static bool init = false;
  if (!init) {
    unsigned char h[256];
    for (int i = 0; i < 256; ++i) h[i] = static_cast<unsigned char>(i / 14);
    YSI_CUDA(cudaMemcpyToSymbol(c_div14, h, sizeof(h)));
    YSI_CUDA(cudaFuncSetAttribute(encoder_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    YSI_CUDA(cudaFuncSetAttribute(encoder_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    init = true;
  }
  const int D = heads * HD;
  const long long rows = static_cast<long long>(n_seq) * T;
  YSI_CHECK(is_global ? T == 4096 : T == 196, "attention kernel supports T = 4096 (global) or 196 (window)");
  const CUtensorMap tmQ = make_tmap_bf16_2d(qkv, rows, 3 * D, 3 * D, BQ);
  const CUtensorMap tmKV = make_tmap_bf16_2d(qkv, rows, 3 * D, 3 * D, BKV);
  const CUtensorMap tmRel = make_tmap_bf16_2d(rel_tab, 256, HD, HD, 128);
  AttnParams p;
  p.T = T; p.D = D; p.out = out;
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  dim3 grid(ceil_div(T, BQ), heads, n_seq);
  if (is_global)
    encoder_attention_kernel<true><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmRel, p);
  else
    encoder_attention_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(tmQ, tmKV, tmRel, p);
  YSI_CUDA(cudaGetLastError());
}

}  // namespace ysi
