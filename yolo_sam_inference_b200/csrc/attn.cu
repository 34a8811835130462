// Fused flash-style attention for the SAM ViT encoder (modeling_sam.py:761-801, 843-882), sm_100a.
//
//   out = softmax(q k^T * hd^-0.5 + q.Rh[qh-kh] + q.Rw[qw-kw]) v      (decomposed rel-pos bias, unscaled q)
//
// Inputs are prepared by the qkv GEMM epilogue / weight loader so that the kernel works in the log2 domain:
// the K columns arrive pre-multiplied by hd^-0.5 * log2(e) and the rel-pos tables by log2(e); p = 2^(x - m).
//
// One CTA = 128 queries of one (sequence, head); keys/values stream through in tiles of 64.
//   warp 4 (one lane) : TMA loads (Q, rel-pos tables, K and V rings) + every tcgen05.mma
//   warps 0-3         : one query row per thread; S from TMEM, bias add (FADD2), max (FMNMX3), exp2 (MUFU),
//                       P -> op16 -> 128B-swizzled smem (double buffered)
// Tensor-core work per tile: S = Q K^T (128x64x64, both K-major) and O += P V (128x64x64, V is the MN-major B
// operand straight out of the qkv activation). O stays in TMEM for the whole CTA: it is rescaled in place
// (tcgen05.ld / tcgen05.st) only when the running row maximum grows by more than 2^8 ("lazy rescale"), so the
// softmax threads never wait on the PV MMA in steady state. S is double buffered in TMEM and each buffer is
// released as soon as it is in registers, so S_{j+1} is complete before the softmax of tile j ends.
// The rel-pos terms:
//   global   (S=64): rel_w = one extra MMA per CTA (Q . Rw^T) whose TMEM result is re-indexed per query into
//                    64 registers per thread (same for every tile). rel_h of tile j (key row kh = j) only
//                    depends on the query row: the two rel_pos_h rows the CTA's two query rows need are
//                    appended to the K tile as extra "keys" (S is 128 x 80), so the bias arrives as column
//                    64 + (qh - qh0) of S and is folded into the exponent offset
//   windowed (S=14): 14+14 registers per thread; 64->70 zero-padded tokens are ordinary keys (they carry the
//                    qkv bias, modeling_sam.py:913-916); the window's 196 keys are resident (tiles 64,64,64,16),
//                    keys >= 196 of the last tile are masked.
// 2 CTAs/SM overlap one CTA's softmax with the other's MMAs; the kernel is MUFU(ex2)-bound by design.
#include "kernels.h"
#include <type_traits>

#include "ptx.cuh"

namespace ysi {

namespace attn {
constexpr int BQ = 128, BKV = 64;
constexpr int THREADS = 160;
constexpr int CH_Q = BQ * 128;              // one 64-column chunk of the Q tile: 128 rows x 128 B = 16 KB
constexpr int KV_BYTES = BKV * 128;         // one 64-column chunk of a 64-key tile: 8 KB
constexpr int P_BYTES = BQ * BKV * 2;       // 16 KB
constexpr float LAZY_LOG2 = 8.0f;           // rescale O only when the row max grows by more than 2^8

// HD = 64 (ViT-B/L) or 80 (ViT-H). Operands are staged in 64-column, 128B-swizzled chunks; for HD = 80 a second
// chunk holds columns 64..127 of which only the first 16 (one k-step) are used -- whatever follows them in the
// qkv row (next head / next section / zero fill) is never touched by an MMA.
template <bool GLOBAL, int HD>
struct Cfg {
  static constexpr int NCH = HD > 64 ? 2 : 1;                       // 64-column chunks per operand
  static constexpr int NKS = HD / 16;                               // k-steps of Q K^T and of the table MMAs
  static constexpr int NST = GLOBAL ? 3 : 4;                        // K / V ring depth (window: whole window)
  static constexpr int K_CHUNK = GLOBAL ? 80 * 128 : 208 * 128;     // global: 64 keys + 16 rows for the rel_pos_h "keys"
  static constexpr int K_STAGE = GLOBAL ? NCH * K_CHUNK : KV_BYTES; // global: a stage holds its chunks back to back;
                                                                    // window: chunk c lives at OFF_K + c*K_CHUNK, tiles 8 KB apart
  static constexpr int K_TOTAL = GLOBAL ? NST * K_STAGE : NCH * K_CHUNK;
  static constexpr int V_CHUNK = GLOBAL ? KV_BYTES : 208 * 128;
  static constexpr int V_STAGE = GLOBAL ? NCH * V_CHUNK : KV_BYTES;
  static constexpr int V_TOTAL = GLOBAL ? NST * V_STAGE : NCH * V_CHUNK;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + NCH * CH_Q;
  static constexpr int OFF_V = OFF_K + K_TOTAL;
  static constexpr int OFF_P = OFF_V + V_TOTAL;      // 2 x 16 KB; fp32 bias scratch [k][128] during setup
  // rel-pos tables as TMA'd for the table MMA: global -> rel_pos_w in the V area (16 KB per chunk, the V loads wait
  // for the MMA); windowed -> second P buffer (4 KB per table chunk), so the whole window's K / V can be requested
  // up front with Q
  static constexpr int TAB_ROWS = GLOBAL ? 128 : 32;        // rows of each rel-pos table fed to the table MMA
  static constexpr int TAB_CHUNK = TAB_ROWS * 128;
  static constexpr int OFF_TABH = GLOBAL ? OFF_K : OFF_P + P_BYTES;     // (unused when GLOBAL)
  static constexpr int OFF_TABW = GLOBAL ? OFF_V : OFF_P + P_BYTES + NCH * TAB_CHUNK;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_N = GLOBAL ? 80 : 64;              // columns of one S buffer
  static constexpr int COL_S = 0;                           // two S buffers (alias the setup tables)
  static constexpr int COL_O = 2 * S_N;
  static constexpr int COL_TH = 0;                          // windowed setup only
  static constexpr int COL_TW = GLOBAL ? 0 : 32;
  static constexpr int CTAS_PER_SM = HD > 64 ? 1 : 2;
  static_assert(HD == 64 || HD == 80, "head_dim 64 or 80");
  static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && OFF_P % 1024 == 0 && K_CHUNK % 1024 == 0, "swizzle atoms need 1 KB alignment");
  static_assert(GLOBAL || 2 * NCH * TAB_CHUNK <= P_BYTES, "window tables must fit the second P buffer");
  static_assert(COL_O + HD <= TMEM_COLS, "TMEM budget");
};
}  // namespace attn

struct AttnParams {
  int T;          // sequence length: 196 (window) or 4096 (global)
  int D;          // heads * head_dim
  int unwindow;   // windowed only: write rows in token order [img*4096 + y*64 + x] and drop the pad tokens
  op16* out;      // [n_seq * T, D]  (or [n_img * 4096, D] when unwindow)
};

template <bool GLOBAL, int HD>
__global__ void __launch_bounds__(attn::THREADS, attn::Cfg<GLOBAL, HD>::CTAS_PER_SM)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const __grid_constant__ CUtensorMap tmKVtail, const __grid_constant__ CUtensorMap tmRel,
                         const __grid_constant__ CUtensorMap tmRel8, AttnParams p) {
  using namespace attn;
  using C = Cfg<GLOBAL, HD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  float* rel_s = reinterpret_cast<float*>(sgen + C::OFF_P);
  const uint32_t bar = sbase + C::OFF_BAR;
  const uint32_t bar_q = bar, bar_tab = bar + 8, bar_rel = bar + 16;
  const uint32_t bar_s_full = bar + 24;    // [2]
  const uint32_t bar_s_free = bar + 40;    // [2]
  const uint32_t bar_p_full = bar + 56;    // [2]
  const uint32_t bar_p_free = bar + 72;    // [2]
  const uint32_t bar_kfull = bar + 88;     // [4]
  const uint32_t bar_kempty = bar + 120;   // [4]
  const uint32_t bar_vfull = bar + 152;    // [4]
  const uint32_t bar_vempty = bar + 184;   // [4]
  const uint32_t tmem_ptr_smem = bar + 216;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int row0 = seq * p.T;                 // first row of this sequence in the qkv matrix
  const int ntiles = GLOBAL ? p.T / BKV : 4;
  const int cq = head * HD, ck = p.D + head * HD, cv = 2 * p.D + head * HD;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_tab, 1); mbar_init(bar_rel, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_s_free + 8 * i, 128);
      mbar_init(bar_p_full + 8 * i, 128); mbar_init(bar_p_free + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_kfull + 8 * i, 1); mbar_init(bar_kempty + 8 * i, 1);
      mbar_init(bar_vfull + 8 * i, 1); mbar_init(bar_vempty + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc_tab = umma_idesc_op16(128, C::TAB_ROWS, 0, 0);
      constexpr uint32_t idesc_s = umma_idesc_op16(128, C::S_N, 0, 0);
      constexpr uint32_t idesc_s16 = umma_idesc_op16(128, 16, 0, 0);
      constexpr uint32_t idesc_pv64 = umma_idesc_op16(128, 64, 0, 1);   // B (= V) is MN-major
      constexpr uint32_t idesc_pv16 = umma_idesc_op16(128, 16, 0, 1);
      constexpr int NCH = C::NCH, NKS = C::NKS;
      // K-major operand, k-step k: chunk k/4 (64 columns each), +32 B per 16 columns inside the swizzle atom
      auto kdesc_at = [&](uint32_t base, int chunk_bytes, int k) {
        return umma_desc_sw128(base + static_cast<uint32_t>((k >> 2) * chunk_bytes), 16, 1024) + 2u * static_cast<uint32_t>(k & 3);
      };
      // ---- setup: Q tile + rel-pos table(s)
      mbar_arrive_expect_tx(bar_q, NCH * (CH_Q + (GLOBAL ? 1 : 2) * C::TAB_CHUNK));
      for (int c = 0; c < NCH; ++c) {
        tma_load_2d(sbase + C::OFF_Q + c * CH_Q, &tmQ, bar_q, cq + 64 * c, row0 + qt * BQ);
        if (!GLOBAL) tma_load_2d(sbase + C::OFF_TABH + c * C::TAB_CHUNK, &tmRel, bar_q, 64 * c, 0);   // rel_pos_h rows (zero padded)
        tma_load_2d(sbase + C::OFF_TABW + c * C::TAB_CHUNK, &tmRel, bar_q, 64 * c, 128);              // rel_pos_w rows
      }
      auto k_tile_addr = [&](int st, int c) {
        return sbase + C::OFF_K + (GLOBAL ? st * C::K_STAGE + c * C::K_CHUNK : c * C::K_CHUNK + st * KV_BYTES);
      };
      auto v_tile_addr = [&](int st, int c) {
        return sbase + C::OFF_V + (GLOBAL ? st * C::V_STAGE + c * C::V_CHUNK : c * C::V_CHUNK + st * KV_BYTES);
      };
      auto load_k = [&](int tile, int st) {
        const bool tail = !GLOBAL && tile == 3;
        mbar_arrive_expect_tx(bar_kfull + 8 * st, NCH * (GLOBAL ? KV_BYTES + 8 * 128 : (tail ? 16 * 128 : KV_BYTES)));
        for (int c = 0; c < NCH; ++c) {
          tma_load_2d(k_tile_addr(st, c), tail ? &tmKVtail : &tmKV, bar_kfull + 8 * st, ck + 64 * c, row0 + tile * BKV);
          // global: rows 64.. of the K tile = rel_pos_h[qh - kh + 63] for the CTA's two query rows (qh0 = 2 qt, kh = tile)
          if (GLOBAL) tma_load_2d(k_tile_addr(st, c) + KV_BYTES, &tmRel8, bar_kfull + 8 * st, 64 * c, 2 * qt - tile + 63);
        }
      };
      auto load_v = [&](int tile, int st) {
        const bool tail = !GLOBAL && tile == 3;
        mbar_arrive_expect_tx(bar_vfull + 8 * st, NCH * (tail ? 16 * 128 : KV_BYTES));
        for (int c = 0; c < NCH; ++c)
          tma_load_2d(v_tile_addr(st, c), tail ? &tmKVtail : &tmKV, bar_vfull + 8 * st, cv + 64 * c, row0 + tile * BKV);
      };
      if (!GLOBAL) {             // whole window: K / V do not alias the tables, request them right away
        for (int j = 0; j < 4; ++j) load_k(j, j);
        for (int j = 0; j < 4; ++j) load_v(j, j);
      } else {                   // the K ring does not alias the rel_pos_w table
        for (int j = 0; j < C::NST && j < ntiles; ++j) load_k(j, j);
      }
      mbar_wait(bar_q, 0);
      tc_fence_after();
      {
        if (!GLOBAL) {
#pragma unroll
          for (int k = 0; k < NKS; ++k)
            umma_op16_ss(tmem_base + C::COL_TH, kdesc_at(sbase + C::OFF_Q, CH_Q, k), kdesc_at(sbase + C::OFF_TABH, C::TAB_CHUNK, k), idesc_tab, k);
        }
#pragma unroll
        for (int k = 0; k < NKS; ++k)
          umma_op16_ss(tmem_base + C::COL_TW, kdesc_at(sbase + C::OFF_Q, CH_Q, k), kdesc_at(sbase + C::OFF_TABW, C::TAB_CHUNK, k), idesc_tab, k);
        umma_commit(bar_tab);
      }
      if (GLOBAL) {
        mbar_wait(bar_tab, 0);     // rel_pos_w table consumed: the V area is free
        for (int j = 0; j < C::NST && j < ntiles; ++j) load_v(j, j);
      }
      // S_t -> S buffer t & 1 (free once the softmax threads hold S_{t-2} in registers); issued two tiles ahead, as
      // soon as S_{t-2} has been read, so it never queues behind a PV in the in-order tensor pipe.
      auto issue_s = [&](int t) {
        const int st = t % C::NST, buf = t & 1;
        mbar_wait(bar_kfull + 8 * st, (t / C::NST) & 1);
        if (t >= 2) mbar_wait(bar_s_free + 8 * buf, ((t >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t idesc = (!GLOBAL && t == 3) ? idesc_s16 : idesc_s;
        const uint32_t d = tmem_base + C::COL_S + buf * C::S_N;
#pragma unroll
        for (int k = 0; k < NKS; ++k)
          umma_op16_ss(d, kdesc_at(sbase + C::OFF_Q, CH_Q, k), kdesc_at(k_tile_addr(st, 0), C::K_CHUNK, k), idesc, k);
        umma_commit(bar_kempty + 8 * st);    // (the barrier somebody always waits on is committed last)
        umma_commit(bar_s_full + 8 * buf);
      };
      mbar_wait(bar_rel, 0);     // bias tables copied out of TMEM: the S / O columns are free
      tc_fence_after();
      issue_s(0);
      if (ntiles > 1) issue_s(1);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 2 < ntiles) issue_s(j + 2);
        const int st = j % C::NST, pb = j & 1;
        // refill the K stage of tile j right away (S_j was issued a whole tile ago): tile j+NST is needed by
        // issue_s two iterations from now, so its TMA latency hides behind two softmax tiles
        if (GLOBAL && j + C::NST < ntiles) {
          mbar_wait(bar_kempty + 8 * st, (j / C::NST) & 1);
          load_k(j + C::NST, st);
        }
        mbar_wait(bar_vfull + 8 * st, (j / C::NST) & 1);
        mbar_wait(bar_p_full + 8 * pb, (j >> 1) & 1);
        tc_fence_after();
        {
          // O (+)= P_j V_j : A = P (K-major, K = keys), B = V_j (MN-major: key rows x 64 hd); 16 keys = 2048 B
          const uint64_t pdesc = umma_desc_sw128(sbase + C::OFF_P + pb * P_BYTES, 16, 1024);
          const uint64_t vdesc = umma_desc_sw128(v_tile_addr(st, 0), 1024, 1024);
          const int ksteps = (!GLOBAL && j == 3) ? 1 : BKV / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_op16_ss(tmem_base + C::COL_O, pdesc + 2u * k, vdesc + 128u * k, idesc_pv64, (j | k) != 0 ? 1u : 0u);
          if (NCH == 2) {            // head columns 64..79: a second, 16-wide MMA from the second V chunk
            const uint64_t vdesc1 = umma_desc_sw128(v_tile_addr(st, 1), 1024, 1024);
            for (int k = 0; k < ksteps; ++k)
              umma_op16_ss(tmem_base + C::COL_O + 64, pdesc + 2u * k, vdesc1 + 128u * k, idesc_pv16, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(bar_vempty + 8 * st);
          umma_commit(bar_p_free + 8 * pb);
        }
        if (GLOBAL) {
          if (j >= 1 && j - 1 + C::NST < ntiles) { // V stage of tile j-1: PV_{j-1} was issued one iteration ago
            const int sv = (j - 1) % C::NST;
            mbar_wait(bar_vempty + 8 * sv, ((j - 1) / C::NST) & 1);
            load_v(j - 1 + C::NST, sv);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread = query row
    const int t = threadIdx.x;                       // 0..127, TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int lq = qt * BQ + t;                      // query index inside the sequence
    const bool q_valid = lq < p.T;
    const bool warp_active = GLOBAL || (qt * BQ + warp * 32) < p.T;   // warps with no valid query only keep the barriers moving
    int qh, qw;
    if (GLOBAL) { qh = lq >> 6; qw = lq & 63; } else { const int l = lq < 196 ? lq : 195; qh = l / 14; qw = l - qh * 14; }
    constexpr int S = GLOBAL ? 64 : 14;
    constexpr int NB = GLOBAL ? 64 : 28;             // bias registers: rel_w[64] | rel_h[14] + rel_w[14]
    float bias[NB];

    mbar_wait(bar_tab, 0);
    tc_fence_after();
    // Q.table^T sits in TMEM as [query][table row]; thread t needs column (q_pos - k_pos + S-1) for every key
    // position: scatter through smem scratch [k][t] (the column is thread dependent), then keep it in registers.
    if (GLOBAL) {
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32p(tlane + C::COL_TW + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kw = qw + (S - 1) - (c0 + i);
          if (kw >= 0 && kw < S) rel_s[kw * 128 + t] = __uint_as_float(r[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) bias[i] = rel_s[i * 128 + t];
    } else {
      uint32_t r[32];
      tmem_ld_x32p(tlane + C::COL_TH, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kh = qh + (S - 1) - i;
        if (kh >= 0 && kh < S) rel_s[kh * 128 + t] = __uint_as_float(r[i]);
      }
      tmem_ld_x32p(tlane + C::COL_TW, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kw = qw + (S - 1) - i;
        if (kw >= 0 && kw < S) rel_s[(14 + kw) * 128 + t] = __uint_as_float(r[i]);
      }
#pragma unroll
      for (int i = 0; i < 28; ++i) bias[i] = rel_s[i * 128 + t];
    }
    tc_fence_before();
    mbar_arrive(bar_rel);

    float m_used = -INFINITY;
    float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
    const uint32_t p_row0 = sbase + C::OFF_P + static_cast<uint32_t>(t) * 128u;
    const uint32_t swz = static_cast<uint32_t>(t & 7);

    // one KV tile: NW columns of S, the first NV of them valid keys
    auto do_tile = [&](auto nw_c, auto nv_c, int j, const float* b /* NW bias values (log2 units) */) {
      constexpr int NW = decltype(nw_c)::value, NV = decltype(nv_c)::value;
      const int pb = j & 1;                       // S buffer and P buffer of this tile
      mbar_wait(bar_s_full + 8 * pb, (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[NW];
      float bh = 0.f;
      if (warp_active) {
        const uint32_t scol = tlane + C::COL_S + static_cast<uint32_t>(pb * C::S_N);
        if constexpr (NW == 64) { tmem_ld_x32p(scol, r); tmem_ld_x32p(scol + 32, r + 32); }
        else tmem_ld_x16p(scol, r);
        if (GLOBAL) {
          uint32_t rb;
          tmem_ld_x1(scol + 64u + static_cast<uint32_t>(warp >> 1), rb);   // q . rel_pos_h[qh - j + 63]: warp-uniform column
          tmem_ld_wait();
          bh = __uint_as_float(rb);
        } else {
          tmem_ld_wait();
        }
      }
      tc_fence_before();
      mbar_arrive(bar_s_free + 8 * pb);
      // P buffer pb was consumed by PV_{j-2}. Every thread waits (also the idle warps of a ragged window tile):
      // S runs two tiles ahead, so without this an idle warp could arrive on bar_p_full for tile j while the
      // barrier's phase of tile j-2 is still open.
      if (j >= 2) mbar_wait(bar_p_free + 8 * pb, ((j - 2) >> 1) & 1);
      if (warp_active) {
        float2 y[NW / 2];
#pragma unroll
        for (int i = 0; i < NW / 2; ++i)
          y[i] = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(b[2 * i], b[2 * i + 1]));
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) mx[i & 3] = max3(mx[i & 3], y[i].x, y[i].y);
        const float m_cand = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) + bh;
        if (__any_sync(0xFFFFFFFFu, m_cand > m_used + LAZY_LOG2)) {
          const float m_new = fmaxf(m_used, m_cand);
          if (j > 0) {
            // fold the new maximum into O (TMEM) and l; PV_{j-1} must have landed, PV_j has not been issued
            const float f = ex2_approx(m_used - m_new);
            mbar_wait(bar_p_free + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int qr = 0; qr < HD / 16; ++qr) {
              uint32_t o[16];
              tmem_ld_x16p(tlane + C::COL_O + 16 * qr, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_x16p(tlane + C::COL_O + 16 * qr, o);
            }
            tmem_st_wait();
            l2a.x *= f; l2a.y *= f; l2b.x *= f; l2b.y *= f;
          }
          m_used = m_new;
        }
        const float c = bh - m_used;
        const float2 c2 = make_float2(c, c);
        const uint32_t p_row = p_row0 + pb * P_BYTES;
#pragma unroll
        for (int ch = 0; ch < NW / 8; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int i = ch * 4 + h;                 // pair index
            float2 e = add2(y[i], c2);
            e.x = (2 * i < NV) ? ex2_approx(e.x) : 0.f;
            e.y = (2 * i + 1 < NV) ? ex2_approx(e.y) : 0.f;
            if (h & 1) l2b = add2(l2b, e); else l2a = add2(l2a, e);
            pk[h] = pack_op16x2(e.x, e.y);
          }
          // K-major, 128B swizzle: 16-byte chunk ch of row t lands at chunk ch ^ (t & 7)
          const uint32_t addr = p_row + ((static_cast<uint32_t>(ch) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
        fence_proxy_async_smem();
        tc_fence_before();
      }
      mbar_arrive(bar_p_full + 8 * pb);
    };

    if (GLOBAL) {
      for (int j = 0; j < ntiles; ++j) do_tile(std::integral_constant<int, 64>{}, std::integral_constant<int, 64>{}, j, bias);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // key k = 64 j + i of the window: kh = k / 14, kw = k % 14 (compile-time after unrolling)
        if (j < 3) {
          float b[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) b[i] = bias[(64 * j + i) / 14] + bias[14 + (64 * j + i) % 14];
          do_tile(std::integral_constant<int, 64>{}, std::integral_constant<int, 64>{}, j, b);
        } else {
          float b[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) b[i] = i < 4 ? bias[(192 + i) / 14] + bias[14 + (192 + i) % 14] : 0.f;
          do_tile(std::integral_constant<int, 16>{}, std::integral_constant<int, 4>{}, j, b);
        }
      }
    }
    mbar_wait(bar_p_free + 8 * ((ntiles - 1) & 1), ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    if (warp_active) {
      uint32_t o[HD];
      tmem_ld_x32p(tlane + C::COL_O, o);
      tmem_ld_x32p(tlane + C::COL_O + 32, o + 32);
      if constexpr (HD > 64) tmem_ld_x16p(tlane + C::COL_O + 64, o + 64);
      tmem_ld_wait();
      const float inv = 1.0f / ((l2a.x + l2a.y) + (l2b.x + l2b.y));
      long long orow = -1;
      if (q_valid) {
        if (!GLOBAL && p.unwindow) {
          const int img = seq / 25, win = seq - img * 25;
          const int y = (win / 5) * 14 + qh, x = (win % 5) * 14 + qw;
          if (y < 64 && x < 64) orow = static_cast<long long>(img) * 4096 + y * 64 + x;
        } else {
          orow = static_cast<long long>(row0) + lq;
        }
      }
      if (orow >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(orow) * p.D + head * HD);
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          uint4 v;
          v.x = pack_op16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          v.y = pack_op16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          v.z = pack_op16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          v.w = pack_op16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
          dst[c] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, Cfg<GLOBAL, HD>::TMEM_COLS);
}

template <bool GLOBAL, int HD>
static void launch_attn_t(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const CUtensorMap& tmKVtail, const CUtensorMap& tmRel,
                          const CUtensorMap& tmRel8, const AttnParams& p, dim3 grid, cudaStream_t stream) {
  using C = attn::Cfg<GLOBAL, HD>;
  static bool init = false;
  if (!init) {
    YSI_CUDA(cudaFuncSetAttribute(encoder_attention_kernel<GLOBAL, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    init = true;
  }
  encoder_attention_kernel<GLOBAL, HD><<<grid, attn::THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, p);
}

// qkv: op16 [n_seq*T, 3D] with the K columns pre-scaled by hd^-0.5*log2(e); rel_tab: op16 [256, HDP] pre-scaled by
// log2(e) (rows 0..127 rel_pos_h zero-padded, 128..255 rel_pos_w; HDP = 64, or 128 with zero columns 80.. for
// head_dim 80). unwindow: see AttnParams.
void launch_encoder_attention(const op16* qkv, const op16* rel_tab, op16* out, int n_seq, int T, int heads, int head_dim,
                              bool is_global, bool unwindow, cudaStream_t stream) {
  using namespace attn;
  YSI_CHECK(head_dim == 64 || head_dim == 80, "attention kernel supports head_dim 64 (ViT-B/L) and 80 (ViT-H)");
  const int D = heads * head_dim, HDP = head_dim == 64 ? 64 : 128;
  const long long rows = static_cast<long long>(n_seq) * T;
  YSI_CHECK(is_global ? T == 4096 : T == 196, "attention kernel supports T = 4096 (global) or 196 (window)");
  YSI_CHECK(!unwindow || (!is_global && n_seq % 25 == 0), "unwindow needs whole images of 25 windows");
  const CUtensorMap tmQ = make_tmap_op16_2d(qkv, rows, 3 * D, 3 * D, BQ);
  const CUtensorMap tmKV = make_tmap_op16_2d(qkv, rows, 3 * D, 3 * D, BKV);
  const CUtensorMap tmKVtail = make_tmap_op16_2d(qkv, rows, 3 * D, 3 * D, 16);
  const CUtensorMap tmRel = make_tmap_op16_2d(rel_tab, 256, HDP, HDP, is_global ? 128 : 32);
  const CUtensorMap tmRel8 = make_tmap_op16_2d(rel_tab, 256, HDP, HDP, 8);
  AttnParams p;
  p.T = T; p.D = D; p.out = out; p.unwindow = unwindow ? 1 : 0;
  dim3 grid(ceil_div(T, BQ), heads, n_seq);
  if (head_dim == 64) {
    if (is_global) launch_attn_t<true, 64>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, p, grid, stream);
    else launch_attn_t<false, 64>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, p, grid, stream);
  } else {
    if (is_global) launch_attn_t<true, 80>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, p, grid, stream);
    else launch_attn_t<false, 80>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, p, grid, stream);
  }
  YSI_CUDA(cudaGetLastError());
}

}  // namespace ysi
