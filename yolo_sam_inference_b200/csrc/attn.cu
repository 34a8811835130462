// Fused flash-style attention for the SAM ViT encoder (modeling_sam.py:761-801, 843-882), sm_100a.
//
//   out = softmax(q k^T * hd^-0.5 + q.Rh[qh-kh] + q.Rw[qw-kw]) v      (decomposed rel-pos bias, unscaled q)
//
// Inputs are prepared by the qkv GEMM epilogue / weight loader so that the kernel works in the log2 domain:
// the K columns arrive pre-multiplied by hd^-0.5 * log2(e) and the rel-pos tables by log2(e); p = 2^(x - m).
//
// One CTA = 128 queries of one (sequence, head); keys/values stream through in tiles of 64.
//   3 producer warps (one elected lane each): TMA loads (Q, rel-pos tables, K and V rings) | table + S MMAs | P.V MMAs
//   softmax warps     : warp w owns rows 32 (w & 3) .. +31 (its TMEM lane quarter): S from TMEM, bias add (FADD2), max
//                       (FMNMX3), exp2 (MUFU), P -> op16 -> TENSOR MEMORY (tcgen05.st), which P.V reads as its A operand.
//                       One thread per query row by default (SPLIT = 1, 4 warps); optionally two (SPLIT = 2, 8 warps,
//                       32 key columns each) -- see SPLIT below. The producer warpgroup hands its registers to the
//                       softmax warps (setmaxnreg), two CTAs per SM.
//   With SPLIT = 2 the two threads of a row must scale P by the same running maximum: every thread publishes its
//   half-row maximum (one flagged shared-memory word), computes its exponentials speculatively against the current
//   (lazily updated) maximum, and only then reads its partner's value; the tile is redone in the rare case (first
//   tiles of a row) that the maximum had to move.
// Tensor-core work per tile: S = Q K^T (128x64x64, both K-major) and O += P V (128x64x64, V is the MN-major B
// operand straight out of the qkv activation). O stays in TMEM for the whole CTA: it is rescaled in place
// (tcgen05.ld / tcgen05.st) only when the running row maximum grows by more than 2^8 ("lazy rescale"), so the
// softmax threads never wait on the PV MMA in steady state. S is double buffered in TMEM; P_j overwrites the first 32
// columns of S_j (it is the TMEM A operand of P.V_j), and S_{j+2} is issued into that buffer when P.V_j completes.
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
#include <cstdio>
#include <type_traits>

#include "ptx.cuh"

namespace ysi {

namespace attn {
constexpr int BQ = 128, BKV = 64;
#ifndef YSI_ATTN_SPLIT
#define YSI_ATTN_SPLIT 1
#endif
// Threads per query row. 1 (default): 4 softmax warps of 216 registers, a thread owns its row's 64 key columns of a tile
// and nothing has to be exchanged. 2 (-DYSI_ATTN_SPLIT=2): 8 softmax warps of 104 registers, 32 columns each, the two
// threads of a row agree on the running maximum through a flagged shared-memory word. Measured on one box, alternating
// builds (scripts/gpu_ab_split.sh): global layers equal (2.92 ms per 8-image step), windowed 1.05 vs 1.19 ms,
// head_dim 80 windowed 5.7 vs 6.9 ms -- the per-tile fixed costs are paid by half as many warps.
constexpr int SPLIT = YSI_ATTN_SPLIT;
static_assert(SPLIT == 1, "the two-threads-per-row variant (round 1, measured slower) was removed with the round-2 softmax pipeline");
// every POLY_EVERY-th pair of exponentials is evaluated with ex2_poly2 on the FMA pipe instead of MUFU.EX2 (0: none)
#ifndef YSI_ATTN_POLY_EVERY
#define YSI_ATTN_POLY_EVERY 8
#endif
constexpr int POLY_EVERY = YSI_ATTN_POLY_EVERY;
#ifndef YSI_ATTN_PSEP
#define YSI_ATTN_PSEP 0
#endif
// timing-only ablation builds (WRONG results; scripts/gpu_attn_ablate.sh): 1 no bias add, 2 no row sums, 4 no row max,
// 8 no exponentials, 16 no 16-bit packing, 32 no S load from tensor memory, 64 no P store, 128 no P.V MMAs, 256 no S MMAs
#ifndef YSI_ATTN_ABLATE
#define YSI_ATTN_ABLATE 0
#endif
constexpr int ABL = YSI_ATTN_ABLATE;
// Sum-checked exponentials (round 2): after the first tile the row maximum is NOT computed. The exponentials are taken against
// the current reference maximum m_used right away and their tile sum (needed anyway) is the overflow detector: a sum below
// NOMAX_LIMIT bounds every P value of the tile (16-bit operand range), anything else -- including inf / NaN -- sends the warp
// through the exact path (row maximum, fold the new reference into O and l, exponentials again). Steady state saves the
// FMNMX pass and its dependency chain between the S load and the first MUFU.
#ifndef YSI_ATTN_NOMAX
#define YSI_ATTN_NOMAX 1
#endif
// Measured (profiles/r02_attention_experiments.txt): global head_dim 64 657 -> 611 us, global 80 919 -> 880, windowed 64 124 -> 120;
// windowed head_dim 80 is 3 % slower with it (169 -> 174 us: one 80-column tile of three cannot skip the maximum) and keeps the
// max-first loop.
template <bool GLOBAL, int HD> constexpr bool nomax_v = (YSI_ATTN_NOMAX != 0) && (GLOBAL || HD == 64);
constexpr float NOMAX_LIMIT = 16384.0f;
// fetch the next S tile from tensor memory under the exponentials of the current one (global layers with PSEP)
// Measured (round 2, profiles/r02_attention_experiments.txt): 704 vs 657 us per ViT-B batch-8 global layer -- slower; the
// skeleton of the tile loop (tensor-memory round trips between the softmax warps and the two MMA warps), not the
// exponentials, bounds the kernel, and splitting the S load lengthens it. Kept as a build option, off by default.
#ifndef YSI_ATTN_PREFETCH
#define YSI_ATTN_PREFETCH 0
#endif
constexpr bool PREFETCH = YSI_ATTN_PREFETCH != 0;
// Global layers, head_dim 64: softmax steps of 128 keys (two key rows of the 64 x 64 grid) on a single S buffer of 144
// columns instead of 64-key tiles on two buffers of 80. The per-step latencies that bound the 64-key loop (tensor-pipe and
// barrier round trips, tensor-memory load / store latencies: profiles/r02_attention_experiments.txt) are paid half as
// often; the two CTAs of an SM alternate between their MMA and their softmax phases.
// Measured (round 2, same file): correct (attention / ViT-B parity / reproducibility tests green) but SLOWER, 2.92 vs 2.45 ms
// per ViT-B batch-8 step: with a single S buffer a CTA's softmax and its MMAs no longer overlap, so each SM sub-partition
// has one runnable softmax warp at a time and nothing hides its tensor-memory and MUFU latencies. Build option, off by default.
#ifndef YSI_ATTN_BIG
#define YSI_ATTN_BIG 0
#endif
constexpr int TW = BKV / SPLIT;              // key columns of a tile owned by one thread
constexpr int SM_WARPS = 4 * SPLIT;          // softmax warps
constexpr int THREADS = 32 * (SM_WARPS + 4); // + one producer warpgroup: TMA warp, S-MMA warp, PV-MMA warp, and a register donor
// Register budget per SM sub-partition (16384 registers = 512 per lane): two CTAs put 4 softmax warps and 2 producer
// warps on each. The kernel launches with <= 80 registers per thread (6 x 80 = 480); the producer warpgroup then
// shrinks to REG_PRODUCER and the softmax warpgroups grow to REG_SOFTMAX with setmaxnreg. The pool is per CTA:
// 8 x 104 + 4 x 32 = 960 = 12 warps x 80 (asking for more than the CTA released deadlocks the allocation).
constexpr int REG_SOFTMAX = SPLIT == 2 ? 104 : 216, REG_PRODUCER = 32;     // SPLIT 1: 8 warps x 128 = 4 x 216 + 4 x 32 + slack
constexpr int CH_Q = BQ * 128;              // one 64-column chunk of the Q tile: 128 rows x 128 B = 16 KB
constexpr int KV_BYTES = BKV * 128;         // one 64-column chunk of a 64-key tile: 8 KB
constexpr int P_BYTES = BQ * BKV * 2;       // 16 KB
constexpr float LAZY_LOG2 = 8.0f;           // rescale O only when the row max grows by more than 2^8

// HD = 64 (ViT-B/L) or 80 (ViT-H). Operands are staged as a 64-column, 128B-swizzled chunk; for HD = 80 the trailing
// 16 columns (one k-step) are a second chunk with 32-byte rows (32B swizzle), a quarter of the first one's size.
template <bool GLOBAL, int HD>
struct Cfg {
  static constexpr bool HAS1 = HD > 64;                             // head_dim 80: a second, 16-column chunk per operand
  static constexpr int NKS = HD / 16;                               // k-steps of Q K^T and of the table MMAs
  // K / V ring depths (window: whole window resident). A K tile is requested well before the S MMA that consumes it
  // and a V tile NSTV - 1 tiles ahead: one tile of lead time (~1 us) does not cover the L2 round trip of the TMA.
  static constexpr bool BIG = GLOBAL && HD == 64 && (YSI_ATTN_BIG != 0);
  static constexpr int STEP = BIG ? 128 : BKV;                      // keys per softmax step
  static constexpr int NSTK = BIG ? 3 : (GLOBAL ? (HAS1 ? 3 : 5) : 4);
  static constexpr int NSTV = BIG ? 2 : 4;
  // chunk 0: columns 0..63 (128-byte rows, SWIZZLE_128B); chunk 1 (HAS1): columns 64..79 (32-byte rows, SWIZZLE_32B)
  static constexpr int Q1_BYTES = HAS1 ? BQ * 32 : 0;
  static constexpr int K_CHUNK = BIG ? 144 * 128 : (GLOBAL ? 80 * 128 : 208 * 128);     // global: keys + 16 rows for the rel_pos_h "keys"
  static constexpr int K1_CHUNK = HAS1 ? K_CHUNK / 4 : 0;
  static constexpr int K_TOTAL = GLOBAL ? NSTK * K_CHUNK : K_CHUNK; // window: tiles 8 KB (chunk 1: 2 KB) apart
  static constexpr int K1_TOTAL = ((GLOBAL ? NSTK * K1_CHUNK : K1_CHUNK) + 1023) / 1024 * 1024;
  static constexpr int V_CHUNK = BIG ? 2 * KV_BYTES : (GLOBAL ? KV_BYTES : 208 * 128);
  static constexpr int V1_CHUNK = HAS1 ? V_CHUNK / 4 : 0;
  static constexpr int V_TOTAL = GLOBAL ? NSTV * V_CHUNK : V_CHUNK;
  static constexpr int V1_TOTAL = ((GLOBAL ? NSTV * V1_CHUNK : V1_CHUNK) + 1023) / 1024 * 1024;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_Q1 = OFF_Q + CH_Q;
  static constexpr int OFF_K = OFF_Q1 + Q1_BYTES;
  static constexpr int OFF_K1 = OFF_K + K_TOTAL;
  static constexpr int OFF_V = OFF_K1 + K1_TOTAL;
  static constexpr int OFF_V1 = OFF_V + V_TOTAL;
  // 32 KB of setup scratch: fp32 bias values [k][128] (P itself lives in TMEM). Global layers and head_dim 80: aliases
  // the V area, whose first loads wait until every softmax thread has its bias in registers (bar_rel). Windowed
  // head_dim 64: its own region, so the whole window's K / V can be requested up front with Q.
  static constexpr int SCRATCH = 32768;
  static constexpr bool ALIAS_V = GLOBAL || HAS1;
  static constexpr int OFF_P = ALIAS_V ? OFF_V : OFF_V1 + V1_TOTAL;
  // rel-pos tables as TMA'd for the table MMA (consumed before the scratch is written): global -> rel_pos_w at the
  // start of the V area; windowed -> 16 KB into the scratch
  static constexpr int TAB_ROWS = GLOBAL ? 128 : 32;        // rows of each rel-pos table fed to the table MMA
  static constexpr int TAB_CHUNK = TAB_ROWS * 128;
  static constexpr int TAB1_CHUNK = HAS1 ? TAB_ROWS * 32 : 0;
  static constexpr int OFF_TABH = GLOBAL ? OFF_K : OFF_P + P_BYTES;     // (unused when GLOBAL)
  static constexpr int OFF_TABW = GLOBAL ? OFF_V : OFF_TABH + TAB_CHUNK;
  static constexpr int OFF_TABH1 = OFF_TABW + TAB_CHUNK;                 // (unused when GLOBAL)
  static constexpr int OFF_TABW1 = GLOBAL ? OFF_TABW + TAB_CHUNK : OFF_TABH1 + TAB1_CHUNK;
  static constexpr int OFF_BAR = ALIAS_V ? OFF_V1 + V1_TOTAL : OFF_P + SCRATCH;       // 512 B of mbarriers
  static constexpr int OFF_XM = OFF_BAR + 512;              // fp32 [2][128][2]: half-row maxima of tile parity 0 / 1
  static constexpr int OFF_XL = OFF_XM + 2048;              // fp32 [128][2]: half-row sums at the end
  static constexpr int SMEM_BYTES = OFF_XL + 1024 + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = 256;                     // 2 S buffers (P_j overwrites the first 32 columns of S_j) + O
  // windowed, one thread per row: the window's 196 keys are three tiles (64, 64, 68 of 80 columns) instead of four
  // (64, 64, 64, 4 of 16): one pass through the per-tile fixed costs less
  static constexpr bool W3 = !GLOBAL && SPLIT == 1;
  static constexpr int S_N = BIG ? 144 : ((GLOBAL || W3) ? 80 : 64);      // columns of one S buffer
  static constexpr int COL_S = 0;                           // two S buffers -- one with BIG -- (alias the setup tables)
  static constexpr int COL_O = BIG ? S_N : 2 * S_N;
  // Global layers at head_dim 64 have 32 tensor-memory columns to spare (2 x 80 + 64 = 224 of 256): P gets its own columns
  // instead of overwriting the S buffer it came from. S_{j+2} then only has to wait until the softmax warps have READ S_j
  // (bar_s_free, early in tile j) instead of until P.V_j has completed, which takes the tensor-pipe round trip
  // (p_full -> P.V_j -> commit -> S_{j+2} -> commit) and the slowest warp of the CTA out of the per-tile critical path.
  static constexpr bool PSEP = GLOBAL && HD == 64 && (YSI_ATTN_PSEP != 0) && !BIG;
  // YSI_ATTN_PSEP=2: P alternates between the dedicated columns (even tiles) and the first columns of S buffer 1 (odd tiles),
  // i.e. P is double buffered inside the same 256 columns: storing P_j then waits for P.V_{j-2} (even j) or for nothing
  // (odd j) instead of for P.V_{j-1}; S_t waits for "S_{t-2} read" (even t) or for P.V_{t-2} (odd t).
  static constexpr bool PALT = PSEP && (YSI_ATTN_PSEP == 2);
  static constexpr int COL_P = COL_O + HD;                  // PSEP only: one P buffer (32 columns = 64 keys)
  // setup tables (Q . R^T): global -> 128 columns over the S buffers (S_0 waits until they have been read); windowed -> 2 x 32
  // columns over O, which is first written by P.V_0 -- long after every softmax thread has its bias values -- so S_0 / S_1
  // are issued while the bias is still being re-indexed. (Round 1 reverted this layout when ViT-H became non-deterministic;
  // the cause was the missing proxy fence on the bias scratch, fixed in round 2, not the layout.)
  static constexpr int COL_TH = GLOBAL ? 0 : COL_O;         // windowed setup only
  static constexpr int COL_TW = GLOBAL ? 0 : COL_O + 32;
  static constexpr int CTAS_PER_SM = 2;
  static_assert(HD == 64 || HD == 80, "head_dim 64 or 80");
  static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && OFF_P % 1024 == 0 && K_CHUNK % 1024 == 0, "swizzle atoms need 1 KB alignment");
  static_assert(OFF_K1 % 256 == 0 && OFF_V1 % 256 == 0 && OFF_Q1 % 256 == 0 && K1_CHUNK % 256 == 0 && V1_CHUNK % 256 == 0, "32B-swizzle atoms need 256 B alignment");
  static_assert(OFF_TABW1 + TAB1_CHUNK <= OFF_P + (ALIAS_V ? V_TOTAL + V1_TOTAL : SCRATCH), "tables must fit the scratch / V area");
  static_assert(!ALIAS_V || V_TOTAL + V1_TOTAL >= (GLOBAL ? SCRATCH : 2 * 28 * 128 * 4), "the bias scratch aliases the V area");
  static_assert(2 * (OFF_BAR + 512 + 2048 + 1024 + 1024 + 1024) <= 232448, "two CTAs per SM");
  static_assert(NSTK <= 8 && NSTV <= 8, "barrier arrays");
  static_assert(COL_O + HD + (PSEP ? 32 : 0) <= TMEM_COLS, "TMEM budget");
};
}  // namespace attn

// Optional timing trace of one CTA (build with -DYSI_ATTN_TRACE, run scripts/gpu_attn_trace.sh): clock64 at the phase
// boundaries of every softmax warp and producer role, printed by the CTA at exit. Not compiled into the product library.
#ifdef YSI_ATTN_TRACE
__device__ long long g_attn_trace[12][64][8];
#define ATTN_TRACE(w, tile, ev) do { if (trace && lane == 0) g_attn_trace[w][(tile) & 63][ev] = clock64(); } while (0)
#else
#define ATTN_TRACE(w, tile, ev) do { } while (0)
#endif

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
                         const __grid_constant__ CUtensorMap tmRel8, const __grid_constant__ CUtensorMap tmQ1,
                         const __grid_constant__ CUtensorMap tmKV1, const __grid_constant__ CUtensorMap tmKVtail1,
                         const __grid_constant__ CUtensorMap tmRel1, const __grid_constant__ CUtensorMap tmRel8_1,
                         AttnParams p) {
  using namespace attn;
  using C = Cfg<GLOBAL, HD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  float* rel_s = reinterpret_cast<float*>(sgen + C::OFF_P);
  const uint32_t bar = sbase + C::OFF_BAR;
  const uint32_t bar_q = bar, bar_tab = bar + 8, bar_rel = bar + 16;
  const uint32_t bar_s_full = bar + 24;    // [2]
  const uint32_t bar_p_full = bar + 40;    // [2] P_j is in TMEM (in S buffer j & 1)
  const uint32_t bar_p_free = bar + 56;    // [2] P.V_j has completed: S buffer j & 1 may be overwritten by S_{j+2}, O is up to date
  const uint32_t bar_kfull = bar + 72;     // [8]
  const uint32_t bar_kempty = bar + 136;   // [8]
  const uint32_t bar_vfull = bar + 200;    // [8]
  const uint32_t bar_vempty = bar + 264;   // [8]
  const uint32_t bar_fin = bar + 328;      // [4 row quarters]: the two warps sharing 32 rows exchange their row sums
  const uint32_t tmem_ptr_smem = bar + 360;
  const uint32_t bar_s_free = bar + 376;   // [2] PSEP: every softmax warp has S buffer j & 1 in registers
  float* xm = reinterpret_cast<float*>(sgen + C::OFF_XM);
  float* xl = reinterpret_cast<float*>(sgen + C::OFF_XL);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int row0 = seq * p.T;                 // first row of this sequence in the qkv matrix
  const int ntiles = GLOBAL ? p.T / C::STEP : (C::W3 ? 3 : 4);      // softmax steps (BIG: 128 keys each)
  const int cq = head * HD, ck = p.D + head * HD, cv = 2 * p.D + head * HD;
  constexpr int NSM = 32 * SM_WARPS;          // softmax threads
#ifdef YSI_ATTN_TRACE
  const bool trace = GLOBAL ? (blockIdx.x == 7 && blockIdx.y == 3 && blockIdx.z == 0) : (blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == 150);
  const long long t_entry = clock64();
#endif

  if (threadIdx.x == 0) {
    // the softmax warps arrive ONCE PER WARP (lane 0 after __syncwarp): per-thread arrivals -- ~800 per tile and CTA --
    // serialise in the SM's barrier unit and were the bottleneck of the whole kernel
    mbar_init(bar_q, 1); mbar_init(bar_tab, 1); mbar_init(bar_rel, SM_WARPS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_s_full + 8 * i, 1);
      mbar_init(bar_p_full + 8 * i, SM_WARPS); mbar_init(bar_p_free + 8 * i, 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(bar_kfull + 8 * i, 1); mbar_init(bar_kempty + 8 * i, 1);
      mbar_init(bar_vfull + 8 * i, 1); mbar_init(bar_vempty + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(bar_fin + 8 * i, 2);
    for (int i = 0; i < 2; ++i) mbar_init(bar_s_free + 8 * i, SM_WARPS);
    fence_mbar_init();
    // Q tile + rel-pos table(s) are requested right here, before the tensor-memory allocation and the CTA barrier: their
    // round trip (~1800 cycles in the windowed kernel's clock trace) is the head of every CTA's latency chain
    mbar_arrive_expect_tx(bar_q, CH_Q + C::Q1_BYTES + (GLOBAL ? 1 : 2) * (C::TAB_CHUNK + C::TAB1_CHUNK));
    tma_load_2d(sbase + C::OFF_Q, &tmQ, bar_q, cq, row0 + qt * BQ);
    if (!GLOBAL) tma_load_2d(sbase + C::OFF_TABH, &tmRel, bar_q, 0, 0);        // rel_pos_h rows (zero padded)
    tma_load_2d(sbase + C::OFF_TABW, &tmRel, bar_q, 0, 128);                   // rel_pos_w rows
    if (C::HAS1) {
      tma_load_2d(sbase + C::OFF_Q1, &tmQ1, bar_q, cq + 64, row0 + qt * BQ);
      if (!GLOBAL) tma_load_2d(sbase + C::OFF_TABH1, &tmRel1, bar_q, 64, 0);
      tma_load_2d(sbase + C::OFF_TABW1, &tmRel1, bar_q, 64, 128);
    }
  }
  for (int i = threadIdx.x; i < 512; i += THREADS) reinterpret_cast<uint32_t*>(xm)[i] = 0xFFFFFFFFu;   // tag 0xFF: nothing published yet
  if (warp == SM_WARPS) {
    tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp >= SM_WARPS) {
    // ------------------------------------------------------------------ producer warpgroup
    // Three single-lane roles, one per warp, so that no role's serial instruction stream (barrier polls, descriptor
    // arithmetic, uniform-datapath issue sequences: ~80 SASS instructions per tile each) paces the softmax:
    //   warp 8: TMA loads   warp 9: table + S = Q K^T MMAs   warp 10: O += P V MMAs   (warp 11 only donates registers)
    if (C::CTAS_PER_SM == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_PRODUCER));
    constexpr int NKS = C::NKS;
    constexpr bool HAS1 = C::HAS1;
    // chunk 0 / chunk 1 of a K or V tile (global: ring stage st; window: tile st of the resident window)
    auto k_tile_addr = [&](int st) { return sbase + C::OFF_K + st * (GLOBAL ? C::K_CHUNK : KV_BYTES); };
    auto k1_tile_addr = [&](int st) { return sbase + C::OFF_K1 + st * (GLOBAL ? C::K1_CHUNK : KV_BYTES / 4); };
    auto v_tile_addr = [&](int st) { return sbase + C::OFF_V + st * (GLOBAL ? C::V_CHUNK : KV_BYTES); };
    auto v1_tile_addr = [&](int st) { return sbase + C::OFF_V1 + st * (GLOBAL ? C::V1_CHUNK : KV_BYTES / 4); };
    // K-major operand, k-step k: steps 0..3 = +32 B per 16 columns inside the 128B swizzle atom of chunk 0, step 4 = chunk 1
    auto kdesc_at = [&](uint32_t base0, uint32_t base1, int k) {
      return k < 4 ? umma_desc_sw128(base0, 16, 1024) + 2u * static_cast<uint32_t>(k) : umma_desc_sw32(base1, 16, 256);
    };
    // Each role runs with the whole warp converged (all lanes poll the barriers) and one elected lane issuing.
    if (warp == SM_WARPS) {
      const bool lead = elect_one();
      // ---- TMA: the K / V rings (Q tile + rel-pos tables were requested by thread 0 before the CTA barrier)
      auto load_k = [&](int tile, int st) {
        if (!lead) return;
        if constexpr (C::BIG) {
          // step `tile` = key rows kh = 2 tile, 2 tile + 1: 128 keys, then 8 rel_pos_h rows starting at (qh0 - kh - 1 + 63) with
          // qh0 = 2 qt, kh = 2 tile: the three rows the CTA's two query rows need for the two key rows are rows 0..2 of them
          mbar_arrive_expect_tx(bar_kfull + 8 * st, 2 * KV_BYTES + 8 * 128);
          tma_load_2d(k_tile_addr(st), &tmKV, bar_kfull + 8 * st, ck, row0 + tile * 128);
          tma_load_2d(k_tile_addr(st) + KV_BYTES, &tmKV, bar_kfull + 8 * st, ck, row0 + tile * 128 + 64);
          tma_load_2d(k_tile_addr(st) + 2 * KV_BYTES, &tmRel8, bar_kfull + 8 * st, 0, 2 * qt - 2 * tile + 62);
          return;
        }
        const bool tail = !GLOBAL && tile == 3;
        const int bytes0 = GLOBAL ? KV_BYTES + 8 * 128 : (tail ? 16 * 128 : KV_BYTES);
        mbar_arrive_expect_tx(bar_kfull + 8 * st, bytes0 + (HAS1 ? bytes0 / 4 : 0));
        tma_load_2d(k_tile_addr(st), tail ? &tmKVtail : &tmKV, bar_kfull + 8 * st, ck, row0 + tile * BKV);
        // global: rows 64.. of the K tile = rel_pos_h[qh - kh + 63] for the CTA's two query rows (qh0 = 2 qt, kh = tile)
        if (GLOBAL) tma_load_2d(k_tile_addr(st) + KV_BYTES, &tmRel8, bar_kfull + 8 * st, 0, 2 * qt - tile + 63);
        if (HAS1) {
          tma_load_2d(k1_tile_addr(st), tail ? &tmKVtail1 : &tmKV1, bar_kfull + 8 * st, ck + 64, row0 + tile * BKV);
          if (GLOBAL) tma_load_2d(k1_tile_addr(st) + KV_BYTES / 4, &tmRel8_1, bar_kfull + 8 * st, 64, 2 * qt - tile + 63);
        }
      };
      auto load_v = [&](int tile, int st) {
        if (!lead) return;
        if constexpr (C::BIG) {
          mbar_arrive_expect_tx(bar_vfull + 8 * st, 2 * KV_BYTES);
          tma_load_2d(v_tile_addr(st), &tmKV, bar_vfull + 8 * st, cv, row0 + tile * 128);
          tma_load_2d(v_tile_addr(st) + KV_BYTES, &tmKV, bar_vfull + 8 * st, cv, row0 + tile * 128 + 64);
          return;
        }
        const bool tail = !GLOBAL && tile == 3;
        const int bytes0 = tail ? 16 * 128 : KV_BYTES;
        mbar_arrive_expect_tx(bar_vfull + 8 * st, bytes0 + (HAS1 ? bytes0 / 4 : 0));
        tma_load_2d(v_tile_addr(st), tail ? &tmKVtail : &tmKV, bar_vfull + 8 * st, cv, row0 + tile * BKV);
        if (HAS1) tma_load_2d(v1_tile_addr(st), tail ? &tmKVtail1 : &tmKV1, bar_vfull + 8 * st, cv + 64, row0 + tile * BKV);
      };
      if (!GLOBAL) {             // whole window resident: K does not alias the tables, request it right away
        for (int j = 0; j < 4; ++j) load_k(j, j);
        if (C::ALIAS_V) mbar_wait(bar_rel, 0);     // the bias scratch sits in the V area: wait until it has been read
        for (int j = 0; j < 4; ++j) load_v(j, j);
      } else {
        for (int j = 0; j < C::NSTK && j < ntiles; ++j) load_k(j, j);      // the K ring does not alias the rel_pos_w table
        mbar_wait(bar_rel, 0);   // every softmax thread has its bias values: the scratch (= V ring) is free
        for (int j = 0; j < C::NSTV && j < ntiles; ++j) load_v(j, j);
        // refill a K stage once its S MMA has completed and a V stage once its P.V has (S runs two tiles ahead of
        // P.V, so waiting in this order never delays a K tile that is needed soon)
        int sk = 0, sv = 0;
        uint32_t pk = 0, pv = 0;
        for (int j = 0; j < ntiles; ++j) {
          if (j + C::NSTK < ntiles) { mbar_wait(bar_kempty + 8 * sk, pk); ATTN_TRACE(8, j, 0); load_k(j + C::NSTK, sk); }
          if (j + C::NSTV < ntiles) { mbar_wait(bar_vempty + 8 * sv, pv); ATTN_TRACE(8, j, 1); load_v(j + C::NSTV, sv); }
          if (++sk == C::NSTK) { sk = 0; pk ^= 1u; }
          if (++sv == C::NSTV) { sv = 0; pv ^= 1u; }
        }
      }
    } else if (warp == SM_WARPS + 1) {
      const bool lead = elect_one();
      // ---- rel-pos table MMA(s), then S_t = Q K_t^T into S buffer t & 1 as soon as P.V_{t-2} -- whose A operand P_{t-2}
      // sits in the first columns of that buffer -- has completed: about one softmax tile before S_t is needed
      constexpr uint32_t idesc_tab = umma_idesc_op16(128, C::TAB_ROWS, 0, 0);
      constexpr uint32_t idesc_s = umma_idesc_op16(128, GLOBAL ? 80 : 64, 0, 0);
      constexpr uint32_t idesc_s80 = umma_idesc_op16(128, 80, 0, 0);
      constexpr uint32_t idesc_s16 = umma_idesc_op16(128, 16, 0, 0);
      uint64_t qdesc[NKS];
#pragma unroll
      for (int k = 0; k < NKS; ++k) qdesc[k] = kdesc_at(sbase + C::OFF_Q, sbase + C::OFF_Q1, k);
      mbar_wait(bar_q, 0);
      tc_fence_after();
      if (lead) {
        if (!GLOBAL) {
#pragma unroll
          for (int k = 0; k < NKS; ++k)
            umma_op16_ss(tmem_base + C::COL_TH, qdesc[k], kdesc_at(sbase + C::OFF_TABH, sbase + C::OFF_TABH1, k), idesc_tab, k);
        }
#pragma unroll
        for (int k = 0; k < NKS; ++k)
          umma_op16_ss(tmem_base + C::COL_TW, qdesc[k], kdesc_at(sbase + C::OFF_TABW, sbase + C::OFF_TABW1, k), idesc_tab, k);
        umma_commit(bar_tab);
      }
      if (GLOBAL) mbar_wait(bar_rel, 0);     // bias tables copied out of TMEM: the S columns are free
      if constexpr (C::BIG) {
        // One thread issues S and P.V of every step in program order: P_j overwrites the first 64 columns of the single S
        // buffer, so S_{j+1} must follow P.V_j in the (in-order) tensor pipe -- no barrier round trip between the two.
        constexpr uint32_t idesc_s144 = umma_idesc_op16(128, 144, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_op16(128, 64, 0, 1);   // B (= V) is MN-major
        const uint32_t d = tmem_base + C::COL_S, otm = tmem_base + C::COL_O;
        auto issue_s = [&](int stage) {
          const uint32_t kbase = k_tile_addr(stage);
#pragma unroll
          for (int k = 0; k < NKS; ++k) umma_op16_ss(d, qdesc[k], umma_desc_sw128(kbase, 16, 1024) + 2u * static_cast<uint32_t>(k), idesc_s144, k);
          umma_commit(bar_kempty + 8 * stage);
          umma_commit(bar_s_full);
        };
        int sk = 0, sv = 0;
        uint32_t pkf = 0, pvf = 0;
        mbar_wait(bar_kfull + 8 * sk, pkf);
        tc_fence_after();
        if (lead) issue_s(sk);
        __syncwarp();
        if (++sk == C::NSTK) { sk = 0; pkf ^= 1u; }
        for (int j = 0; j < ntiles; ++j) {
          mbar_wait(bar_vfull + 8 * sv, pvf);
          mbar_wait(bar_p_full, static_cast<uint32_t>(j & 1));
          tc_fence_after();
          if (lead) {
            const uint64_t vdesc = umma_desc_sw128(v_tile_addr(sv), 1024, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_op16_ts(otm, d + 8u * k, vdesc + 128u * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
            umma_commit(bar_vempty + 8 * sv);
            umma_commit(bar_p_free);
          }
          __syncwarp();
          if (++sv == C::NSTV) { sv = 0; pvf ^= 1u; }
          if (j + 1 < ntiles) {
            mbar_wait(bar_kfull + 8 * sk, pkf);
            tc_fence_after();
            if (lead) issue_s(sk);
            __syncwarp();
            if (++sk == C::NSTK) { sk = 0; pkf ^= 1u; }
          }
        }
      }
      int st = 0;
      uint32_t ph = 0;
      for (int t = 0; !C::BIG && t < ntiles; ++t) {
        const int buf = t & 1;
        mbar_wait(bar_kfull + 8 * st, ph);
        if (C::W3 && t == 2) mbar_wait(bar_kfull + 8 * 3, 0);      // the last window tile spans K tiles 2 and 3
        ATTN_TRACE(9, t, 0);
        if (t >= 2) {
          if (C::PSEP && !(C::PALT && buf)) mbar_wait(bar_s_free + 8 * buf, ((t >> 1) - 1) & 1);   // S_{t-2} has been read out of buffer t & 1
          else mbar_wait(bar_p_free + 8 * buf, ((t >> 1) - 1) & 1);           // P.V_{t-2} done: S / P buffer t & 1 is free
        }
        ATTN_TRACE(9, t, 1);
        tc_fence_after();
        const uint32_t idesc = (!GLOBAL && t == 3) ? idesc_s16 : ((C::W3 && t == 2) ? idesc_s80 : idesc_s);
        const uint32_t d = tmem_base + C::COL_S + buf * C::S_N;
        const uint32_t kbase = k_tile_addr(st), kbase1 = k1_tile_addr(st);
        if (lead) {
#pragma unroll
          for (int k = 0; k < ((ABL & 256) ? 0 : NKS); ++k) umma_op16_ss(d, qdesc[k], kdesc_at(kbase, kbase1, k), idesc, k);
          umma_commit(bar_kempty + 8 * st);
          umma_commit(bar_s_full + 8 * buf);
        }
        __syncwarp();
        ATTN_TRACE(9, t, 2);
        if (++st == C::NSTK) { st = 0; ph ^= 1u; }
      }
    } else if (warp == SM_WARPS + 2) {
      const bool lead = elect_one();
      // ---- O (+)= P_j V_j : A = P straight out of TENSOR MEMORY (row = lane, two op16 keys per 32-bit column, 8 columns
      // per 16-key k-step) -- no shared-memory round trip for P; B = V_j (MN-major: key rows x 64 hd), 16 keys = 2048 B
      constexpr uint32_t idesc_pv64 = umma_idesc_op16(128, 64, 0, 1);   // B (= V) is MN-major
      constexpr uint32_t idesc_pv16 = umma_idesc_op16(128, 16, 0, 1);
      const uint32_t otm = tmem_base + C::COL_O;
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; !C::BIG && j < ntiles; ++j) {
        mbar_wait(bar_vfull + 8 * st, ph);
        if (C::W3 && j == 2) mbar_wait(bar_vfull + 8 * 3, 0);
        ATTN_TRACE(10, j, 0);
        mbar_wait(bar_p_full + 8 * (j & 1), (j >> 1) & 1);
        ATTN_TRACE(10, j, 1);
        tc_fence_after();
        const uint32_t ptm = (C::PSEP && !(C::PALT && (j & 1))) ? tmem_base + C::COL_P
                                                                : tmem_base + C::COL_S + static_cast<uint32_t>((j & 1) * C::S_N);
        if (lead) {
          const uint64_t vdesc = umma_desc_sw128(v_tile_addr(st), 1024, 1024);
          if (ABL & 128) {
          } else if (GLOBAL || j < 3) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) umma_op16_ts(otm, ptm + 8u * k, vdesc + 128u * k, idesc_pv64, (j | k) != 0 ? 1u : 0u);
            if (C::W3 && j == 2) umma_op16_ts(otm, ptm + 8u * 4, vdesc + 128u * 4, idesc_pv64, 1u);     // keys 192..207 (196.. are zero in P)
          } else {
            umma_op16_ts(otm, ptm, vdesc, idesc_pv64, 1u);
          }
          if (HAS1) {                // head columns 64..79: a second, 16-wide MMA from the V chunk with 32-byte rows
            const uint64_t vdesc1 = umma_desc_sw32(v1_tile_addr(st), 256, 256);      // MN-major: 8 key rows per 256 B group
            const int ksteps = (!GLOBAL && j == 3) ? 1 : ((C::W3 && j == 2) ? 5 : BKV / 16);
            for (int k = 0; k < ksteps; ++k) umma_op16_ts(otm + 64, ptm + 8u * k, vdesc1 + 32u * k, idesc_pv16, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(bar_vempty + 8 * st);
          umma_commit(bar_p_free + 8 * (j & 1));
        }
        __syncwarp();
        ATTN_TRACE(10, j, 2);
        if (++st == C::NSTV) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: two threads per query row
    if (C::CTAS_PER_SM == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_SOFTMAX));
    const int rq = warp & 3;                         // row quarter = TMEM lane quarter of this warp
    const int half = SPLIT == 2 ? warp >> 2 : 0;     // which TW key columns of every 64-key tile
    const int t = rq * 32 + lane;                    // query row inside the tile = TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(rq * 32) << 16);
    const int lq = qt * BQ + t;                      // query index inside the sequence
    const bool q_valid = lq < p.T;
    const bool warp_active = GLOBAL || (qt * BQ + rq * 32) < p.T;   // warps with no valid query only keep the barriers moving
    int qh, qw;
    if (GLOBAL) { qh = lq >> 6; qw = lq & 63; } else { const int l = lq < 196 ? lq : 195; qh = l / 14; qw = l - qh * 14; }
    constexpr int S = GLOBAL ? 64 : 14;
    constexpr int NB = GLOBAL ? TW : 28;             // bias registers: rel_w of this thread's TW keys | rel_h[14] + rel_w[14]
    float bias[NB];

    ATTN_TRACE(warp, 60, 0);
    mbar_wait(bar_tab, 0);
    ATTN_TRACE(warp, 60, 1);
    tc_fence_after();
    // Q.table^T sits in TMEM as [query][table row]; thread t needs column (q_pos - k_pos + S-1) for every key
    // position it owns: scatter through smem scratch [k][t] (the column is thread dependent), then keep it in registers.
    if (GLOBAL) {
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32p(tlane + C::COL_TW + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kw = qw + (S - 1) - (c0 + i);
          if (kw >= 0 && kw < 64 && kw / TW == half) rel_s[kw * 128 + t] = __uint_as_float(r[i]);     // in this thread's columns
        }
      }
#pragma unroll
      for (int i = 0; i < TW; ++i) bias[i] = rel_s[(TW * half + i) * 128 + t];
    } else {
      // both threads of a row keep all 28 values; each writes / reads its own scratch column block
      float* my = rel_s + half * (28 * 128);
      uint32_t r[32];
      tmem_ld_x32p(tlane + C::COL_TH, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kh = qh + (S - 1) - i;
        if (kh >= 0 && kh < S) my[kh * 128 + t] = __uint_as_float(r[i]);
      }
      tmem_ld_x32p(tlane + C::COL_TW, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kw = qw + (S - 1) - i;
        if (kw >= 0 && kw < S) my[(14 + kw) * 128 + t] = __uint_as_float(r[i]);
      }
#pragma unroll
      for (int i = 0; i < 28; ++i) bias[i] = my[i * 128 + t];
    }
    // The scratch this thread just wrote and read back through the generic proxy is about to be overwritten by the TMA
    // (async proxy) with V tiles, as soon as the last softmax warp has arrived on bar_rel. Accesses of the two proxies
    // to the same shared memory are only ordered by a proxy fence: without it V bytes could land before this warp's
    // reads had been performed (a warp's bias registers then held V data) or a scratch store could land after them
    // (corrupt V for the whole CTA). Found in round 2 with the run-to-run bit-equality test at ViT-H, batch 8
    // (profiles/r02_race_diag_before_fix.txt): windowed head_dim-80 attention differed in whole 32-row blocks.
    ATTN_TRACE(warp, 60, 2);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_rel);

    bool s_ready = false;                            // S of the next tile already seen complete (probed a tile early)
    bool p_pending = false;                          // P of the previous tile is stored (tcgen05.st issued) but not yet handed to the PV warp
    uint32_t p_pending_bar = 0;
    float m_used = -INFINITY;
    float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
    constexpr int OH = HD / SPLIT;                   // O columns owned by this thread
    const uint32_t ocol = tlane + C::COL_O + static_cast<uint32_t>(half * OH);

    // hand P of the previous tile to the P.V warp: its tcgen05.st was issued at the end of that tile and has had the whole
    // S load / bias / max phase of the current tile to complete, so this wait is (almost) free
    auto finalize_p = [&]() {
      if (p_pending) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_pending_bar);
        p_pending = false;
      }
    };

    // one KV tile: this thread owns NW columns of S starting at column c0 (NW = 0: nothing, only the barriers),
    // the first NV of them valid keys; bfn(i) = bias of column i (log2 units; i is a compile-time constant after unrolling)
    auto do_tile = [&](auto nw_c, auto nv_c, int j, auto bfn, int c0) {
      constexpr int NW = decltype(nw_c)::value, NV = decltype(nv_c)::value;
      constexpr int NR = NW > 0 ? NW : 2;
      const int pb = j & 1;                       // S buffer and P buffer of this tile
      const uint32_t par = static_cast<uint32_t>((j >> 1) & 1);
      ATTN_TRACE(warp, j, 0);
      if (!s_ready) mbar_wait(bar_s_full + 8 * pb, par);
      ATTN_TRACE(warp, j, 1);
      tc_fence_after();
      uint32_t r[NR];
      float bh = 0.f;
      const bool act = warp_active && NW > 0;
      if (act) {
        const uint32_t scol = tlane + C::COL_S + static_cast<uint32_t>(pb * C::S_N);
        if constexpr (NW == 80) { tmem_ld_x32p(scol, r); tmem_ld_x32p(scol + 32, r + 32); tmem_ld_x16p(scol + 64, r + 64); }
        else if constexpr (NW == 64) {
          if (ABL & 32) {
#pragma unroll
            for (int i = 0; i < 64; ++i) r[i] = 0x3c000000u + i;
          } else { tmem_ld_x32p(scol + c0, r); tmem_ld_x32p(scol + c0 + 32, r + 32); }
        }
        else if constexpr (NW == 32) tmem_ld_x32p(scol + c0, r);
        else if constexpr (NW == 16) tmem_ld_x16p(scol + c0, r);
        if (GLOBAL) {
          uint32_t rb;
          tmem_ld_x1(scol + 64u + static_cast<uint32_t>(rq >> 1), rb);   // q . rel_pos_h[qh - j + 63]: warp-uniform column
          tmem_ld_wait();
          bh = __uint_as_float(rb);
        } else {
          tmem_ld_wait();
        }
      }
      if (C::PSEP) {               // S_j is in registers: the S warp may overwrite this buffer with S_{j+2}
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_s_free + 8 * pb);
      }
      ATTN_TRACE(warp, j, 2);
      float2 y[NR / 2];
      if (act) {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i)
          y[i] = (ABL & 1) ? make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]))
                           : add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(bfn(2 * i), bfn(2 * i + 1)));
      }
      // row maximum of this tile (log2 units, rel_pos_h term included)
      auto row_max = [&]() -> float {
        if (!act) return -INFINITY;
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < ((ABL & 4) ? 1 : NV / 2); ++i) mx[i & 3] = max3(mx[i & 3], y[i].x, y[i].y);
        return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) + bh;
      };
      // move the reference maximum: fold the factor into O (TMEM) and l. PV_{j-1} must have landed (its P was handed over by
      // finalize_p above, by every warp before its own check), PV_j has not been issued.
      auto move_reference = [&](float m_new) {
        if (j > 0) {
          const float f = ex2_approx(m_used - m_new);
          mbar_wait(bar_p_free + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int qr = 0; qr < OH / 8; ++qr) {
            uint32_t o[8];
            tmem_ld_x8p(ocol + 8 * qr, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_x8p(ocol + 8 * qr, o);
          }
          tmem_st_wait();
          l2a.x *= f; l2a.y *= f; l2b.x *= f; l2b.y *= f;
        }
        m_used = m_new;
      };
      uint32_t pk[NR / 2];
      float2 ta = make_float2(0.f, 0.f), tb = make_float2(0.f, 0.f);
      // exponentials against the current reference maximum -> packed P values + their sums
      auto exp_pass = [&]() {
        const float c = bh - m_used;
        const float2 c2 = make_float2(c, c);
        ta = make_float2(0.f, 0.f); tb = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) {
          float2 e = add2(y[i], c2);
          // a fixed share of the pairs takes the polynomial on the FMA pipe instead of the MUFU
          if (ABL & 8) { }
          else if (POLY_EVERY > 0 && (i % (POLY_EVERY > 0 ? POLY_EVERY : 1)) == (POLY_EVERY - 1) && 2 * i + 1 < NV) e = ex2_poly2(e);
          else {
            e.x = (2 * i < NV) ? ex2_approx(e.x) : 0.f;
            e.y = (2 * i + 1 < NV) ? ex2_approx(e.y) : 0.f;
          }
          if (!(ABL & 2)) { if (i & 1) tb = add2(tb, e); else ta = add2(ta, e); }
          pk[i] = (ABL & 16) ? __float_as_uint(e.x) : pack_op16x2(e.x, e.y);
        }
      };
      finalize_p();
      ATTN_TRACE(warp, j, 3);
      if constexpr (nomax_v<GLOBAL, HD>) {
        // Tile 0 fixes the reference maximum; later tiles go straight to the exponentials and use their sum as the overflow
        // detector (see NOMAX above). The exact path is warp-uniform and rare after the first tiles of a row.
        if (j == 0) m_used = row_max();
        ATTN_TRACE(warp, j, 7);
        if (act) exp_pass();
        ATTN_TRACE(warp, j, 4);
        if (j > 0) {
          const float ts = (ta.x + ta.y) + (tb.x + tb.y);
          if (__any_sync(0xFFFFFFFFu, act && !(ts <= NOMAX_LIMIT))) {
            move_reference(fmaxf(m_used, row_max()));
            if (act) exp_pass();
          }
        }
      } else {
        // Lazy rescale: the reference maximum m_used only moves when a row's maximum has grown by more than 2^8 (P stays
        // below 2^8 in the 16-bit operand). Decided BEFORE the exponentials, which are therefore computed exactly once.
        const float m_tile = row_max();
        if (__any_sync(0xFFFFFFFFu, m_tile > m_used + LAZY_LOG2)) move_reference(fmaxf(m_used, m_tile));
        ATTN_TRACE(warp, j, 4);
        if (act) exp_pass();
      }
      if (act) { l2a = add2(l2a, ta); l2b = add2(l2b, tb); }
      // P_j goes into the first columns of S buffer pb. They are free: S_j (this buffer) was only issued after P.V_{j-2}
      // had completed and this thread has its own S columns in registers. No wait on the tensor pipe in steady state.
      ATTN_TRACE(warp, j, 5);
      // probe the next tile's S now: the barrier unit's round trip overlaps the P store below
      s_ready = mbar_test_wait(bar_s_full + 8 * (pb ^ 1), static_cast<uint32_t>(((j + 1) >> 1) & 1));
      if (C::PALT) {               // P alternates: even tiles own the dedicated columns (free once P.V_{j-2} has completed),
        if (!pb && j >= 2) mbar_wait(bar_p_free, ((j - 2) >> 1) & 1);     // odd tiles overwrite their own S buffer (no wait)
      } else if (C::PSEP && j > 0) {      // the single P buffer is free once P.V_{j-1} has completed (normally long ago)
        mbar_wait(bar_p_free + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
      }
      if (act) {
        tc_fence_after();
        const uint32_t pcol = (C::PSEP && !(C::PALT && pb)) ? tlane + C::COL_P : tlane + C::COL_S + static_cast<uint32_t>(pb * C::S_N + c0 / 2);
        if constexpr (NW == 80) { tmem_st_x32p(pcol, pk); tmem_st_x8p(pcol + 32, pk + 32); }
        else if constexpr (NW == 64) { if (!(ABL & 64)) tmem_st_x32p(pcol, pk); }
        else if constexpr (NW == 32) tmem_st_x16p(pcol, pk);
        else if constexpr (NW == 16) tmem_st_x8p(pcol, pk);
      }
      // the hand-over (tcgen05.wait::st, fence, arrive on p_full) is deferred to the next tile, after its S load
      p_pending = true;
      p_pending_bar = bar_p_full + 8 * pb;
      ATTN_TRACE(warp, j, 6);
    };

    using ITW = std::integral_constant<int, TW>;
    using I32 = std::integral_constant<int, 32>;
    using I16 = std::integral_constant<int, 16>;
    using I4 = std::integral_constant<int, 4>;
    using I0 = std::integral_constant<int, 0>;
    if constexpr (C::BIG) {
      // 128 keys (key rows kh = 2 j and 2 j + 1) per step. Two passes over the S buffer so that only 64 scores are live at a
      // time: pass 1 = row maximum, pass 2 = exponentials + P. P half h (32 columns) overwrites S columns 32 h .. 32 h + 31,
      // which this thread has consumed by then (it only ever touches its own lane).
      const uint32_t scol = tlane + C::COL_S;
      const uint32_t hcol = scol + 129u + static_cast<uint32_t>(rq >> 1);       // rel_pos_h term of key row half h: column hcol - h
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(bar_s_full, static_cast<uint32_t>(j & 1));
        tc_fence_after();
        float bh[2];
        float m_tile = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[64], rb;
          tmem_ld_x32p(scol + 64u * h, r);
          tmem_ld_x32p(scol + 64u * h + 32u, r + 32);
          tmem_ld_x1(hcol - static_cast<uint32_t>(h), rb);
          tmem_ld_wait();
          bh[h] = __uint_as_float(rb);
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float2 y = add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(bias[2 * i], bias[2 * i + 1]));
            mx[i & 3] = max3(mx[i & 3], y.x, y.y);
          }
          m_tile = fmaxf(m_tile, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) + bh[h]);
        }
        if (__any_sync(0xFFFFFFFFu, m_tile > m_used + LAZY_LOG2)) {          // lazy rescale, see do_tile
          const float m_new = fmaxf(m_used, m_tile);
          if (j > 0) {
            const float f = ex2_approx(m_used - m_new);
            mbar_wait(bar_p_free, static_cast<uint32_t>((j - 1) & 1));       // P.V_{j-1}: complete since S_j is (same pipe, in order)
            tc_fence_after();
#pragma unroll
            for (int qr = 0; qr < OH / 8; ++qr) {
              uint32_t o[8];
              tmem_ld_x8p(ocol + 8 * qr, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_x8p(ocol + 8 * qr, o);
            }
            l2a.x *= f; l2a.y *= f; l2b.x *= f; l2b.y *= f;
          }
          m_used = m_new;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[64];
          tmem_ld_x32p(scol + 64u * h, r);
          tmem_ld_x32p(scol + 64u * h + 32u, r + 32);
          tmem_ld_wait();
          const float c = bh[h] - m_used;
          const float2 c2 = make_float2(c, c);
          float2 ta = make_float2(0.f, 0.f), tb = make_float2(0.f, 0.f);
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float2 e = add2(add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(bias[2 * i], bias[2 * i + 1])), c2);
            if (POLY_EVERY > 0 && (i % (POLY_EVERY > 0 ? POLY_EVERY : 1)) == (POLY_EVERY - 1)) e = ex2_poly2(e);
            else { e.x = ex2_approx(e.x); e.y = ex2_approx(e.y); }
            if (i & 1) tb = add2(tb, e); else ta = add2(ta, e);
            pk[i] = pack_op16x2(e.x, e.y);
          }
          l2a = add2(l2a, ta); l2b = add2(l2b, tb);
          tmem_st_x32p(scol + 32u * h, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_full);
      }
    } else if constexpr (C::PSEP && PREFETCH) {
      // Global layers, head_dim 64: the S tile of step j + 1 is fetched from tensor memory WHILE the exponentials of step j
      // run. The kernel's skeleton is bound by the tensor-memory read port (128 x 65 fp32 per tile and CTA at 64 B / clk / SM
      // -- as long as the MUFU work of the same tile), so the two must overlap instead of alternating. S_{j+1} is
      // available that early because P has its own columns (PSEP): the S warp only waits for bar_s_free.
      // One register set r[64]: its first half is dead as soon as the first 32 scores have been biased, and is refilled with
      // the first half of S_{j+1} at the start of the exponentials; the second half of S_j is fetched at the start of tile j
      // and lands under the bias / max work of the first half.
      uint32_t r[64], bh_bits = 0;
      auto fetch_lo = [&](int jj) {
        mbar_wait(bar_s_full + 8 * (jj & 1), static_cast<uint32_t>((jj >> 1) & 1));
        tc_fence_after();
        tmem_ld_x32p(tlane + C::COL_S + static_cast<uint32_t>((jj & 1) * C::S_N), r);
      };
      fetch_lo(0);
      for (int j = 0; j < ntiles; ++j) {
        const int pb = j & 1;
        const uint32_t scol = tlane + C::COL_S + static_cast<uint32_t>(pb * C::S_N);
        tmem_ld_wait();                                   // first half of S_j (fetched during tile j - 1)
        tmem_ld_x32p(scol + 32, r + 32);
        tmem_ld_x1(scol + 64u + static_cast<uint32_t>(rq >> 1), bh_bits);      // q . rel_pos_h[qh - j + 63]
        float2 y[32];
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          y[i] = (ABL & 1) ? make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]))
                           : add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(bias[2 * i], bias[2 * i + 1]));
          if (!(ABL & 4) || i == 0) mx[i & 3] = max3(mx[i & 3], y[i].x, y[i].y);
        }
        tmem_ld_wait();                                   // second half + the rel_pos_h term
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_s_free + 8 * pb);  // S_j is in registers: the S warp may overwrite this buffer with S_{j+2}
        const float bh = __uint_as_float(bh_bits);
#pragma unroll
        for (int i = 16; i < 32; ++i) {
          y[i] = (ABL & 1) ? make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]))
                           : add2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(bias[2 * i], bias[2 * i + 1]));
          if (!(ABL & 4)) mx[i & 3] = max3(mx[i & 3], y[i].x, y[i].y);
        }
        const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) + bh;
        finalize_p();
        if (__any_sync(0xFFFFFFFFu, m_tile > m_used + LAZY_LOG2)) {          // lazy rescale, see do_tile
          const float m_new = fmaxf(m_used, m_tile);
          if (j > 0) {
            const float f = ex2_approx(m_used - m_new);
            mbar_wait(bar_p_free + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int qr = 0; qr < OH / 8; ++qr) {
              uint32_t o[8];
              tmem_ld_x8p(ocol + 8 * qr, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_x8p(ocol + 8 * qr, o);
            }
            tmem_st_wait();
            l2a.x *= f; l2a.y *= f; l2b.x *= f; l2b.y *= f;
          }
          m_used = m_new;
        }
        if (j + 1 < ntiles) fetch_lo(j + 1);              // lands under the exponentials below
        const float c = bh - m_used;
        const float2 c2 = make_float2(c, c);
        float2 ta = make_float2(0.f, 0.f), tb = make_float2(0.f, 0.f);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float2 e = add2(y[i], c2);
          if (ABL & 8) { }
          else if (POLY_EVERY > 0 && (i % (POLY_EVERY > 0 ? POLY_EVERY : 1)) == (POLY_EVERY - 1)) e = ex2_poly2(e);
          else { e.x = ex2_approx(e.x); e.y = ex2_approx(e.y); }
          if (!(ABL & 2)) { if (i & 1) tb = add2(tb, e); else ta = add2(ta, e); }
          pk[i] = (ABL & 16) ? __float_as_uint(e.x) : pack_op16x2(e.x, e.y);
        }
        l2a = add2(l2a, ta); l2b = add2(l2b, tb);
        if (j > 0) mbar_wait(bar_p_free + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);    // the single P buffer is free again
        tc_fence_after();
        tmem_st_x32p(tlane + C::COL_P, pk);
        p_pending = true;
        p_pending_bar = bar_p_full + 8 * pb;
      }
    } else if (GLOBAL) {
      for (int j = 0; j < ntiles; ++j) do_tile(ITW{}, ITW{}, j, [&](int i) { return bias[i]; }, TW * half);
    } else {
      // key k = 64 j + 32 half + i of the window: kh = k / 14, kw = k % 14 (compile-time after unrolling)
      auto window_tiles = [&](auto half_c) {
        constexpr int HF = decltype(half_c)::value;
        auto tile = [&](auto j_c) {
          constexpr int K0 = 64 * decltype(j_c)::value + TW * HF;
          do_tile(ITW{}, ITW{}, decltype(j_c)::value, [&](int i) { return bias[(K0 + i) / 14] + bias[14 + (K0 + i) % 14]; }, TW * HF);
        };
        tile(std::integral_constant<int, 0>{}); tile(std::integral_constant<int, 1>{});
        if constexpr (C::W3) {   // keys 128..195 + 12 masked columns in one 80-column tile
          do_tile(std::integral_constant<int, 80>{}, std::integral_constant<int, 68>{}, 2,
                  [&](int i) { return i < 68 ? bias[(128 + i) / 14] + bias[14 + (128 + i) % 14] : 0.f; }, 0);
          return;
        }
        tile(std::integral_constant<int, 2>{});
        if (HF == 0)        // ragged last tile (keys 192..195 + 12 masked columns): the first half of the pair takes it
          do_tile(I16{}, I4{}, 3, [&](int i) { return i < 4 ? bias[(192 + i) / 14] + bias[14 + (192 + i) % 14] : 0.f; }, 0);
        else
          do_tile(I0{}, I0{}, 3, [&](int) { return 0.f; }, 0);
      };
      if (half == 0) window_tiles(std::integral_constant<int, 0>{}); else window_tiles(std::integral_constant<int, 1>{});
    }
    ATTN_TRACE(warp, 60, 3);
    finalize_p();
    // row sum = both halves
    xl[t * 2 + half] = (l2a.x + l2a.y) + (l2b.x + l2b.y);
    if (SPLIT == 1) xl[t * 2 + 1] = 0.f;
    __syncwarp();
    if (SPLIT == 2 && lane == 0) mbar_arrive(bar_fin + 8 * rq);
    if (C::BIG) mbar_wait(bar_p_free, static_cast<uint32_t>((ntiles - 1) & 1));
    else mbar_wait(bar_p_free + 8 * ((ntiles - 1) & 1), ((ntiles - 1) >> 1) & 1);
    tc_fence_after();
    ATTN_TRACE(warp, 60, 4);
    if (SPLIT == 2) mbar_wait(bar_fin + 8 * rq, 0);
    if (warp_active) {
      uint32_t o[OH];
#pragma unroll
      for (int c = 0; c < OH; c += 8) tmem_ld_x8p(ocol + c, o + c);
      tmem_ld_wait();
      ATTN_TRACE(warp, 60, 6);
      const float inv = 1.0f / (xl[t * 2] + xl[t * 2 + 1]);
      int orow = -1;
      if (q_valid) {
        if (!GLOBAL && p.unwindow) {
          const int img = seq / 25, win = seq - img * 25;
          const int y = (win / 5) * 14 + qh, x = (win % 5) * 14 + qw;
          if (y < 64 && x < 64) orow = img * 4096 + y * 64 + x;
        } else {
          orow = row0 + lq;
        }
      }
      // One row per lane would make every 16-byte store instruction touch 32 different lines (the windowed kernel's clock
      // trace showed 2400 cycles in this epilogue). The warp's 32 rows are staged through the idle Q tile instead (all S /
      // table MMAs that read it have completed) and leave as whole rows: consecutive lanes write consecutive 16-byte chunks
      // (1450 cycles; windowed head_dim 64: 115.6 -> 111.6 us per ViT-B layer). Chunk c of row r sits at slot (c + r) mod NCH
      // of its row, which spreads a store instruction's 32 chunks over the banks. Tried and rejected: one 128-byte
      // cp.async.bulk per lane out of padded rows -- the copy engine needs ~2200 cycles for a warp's 32 small copies (113 us).
      constexpr int NCH = HD / 8, ROWB = HD * 2;
      static_assert(SPLIT == 1 && 128 * ROWB <= CH_Q + C::Q1_BYTES, "the output stage aliases the Q tile");
      uint8_t* stage = sgen + C::OFF_Q + rq * 32 * ROWB;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint4 v;
        v.x = pack_op16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
        v.y = pack_op16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
        v.z = pack_op16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
        v.w = pack_op16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(stage + lane * ROWB + ((c + lane) % NCH) * 16) = v;
      }
      __syncwarp();
      ATTN_TRACE(warp, 60, 7);
#pragma unroll
      for (int it = 0; it < NCH; ++it) {
        const int f = it * 32 + lane, row = f / NCH, slot = f - row * NCH;
        const int c = (slot + NCH - row % NCH) % NCH;
        const int dst_row = __shfl_sync(0xFFFFFFFFu, orow, row);
        if (dst_row >= 0)
          *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(dst_row) * p.D + head * HD + c * 8) =
              *reinterpret_cast<const uint4*>(stage + f * 16);
      }
    }
  }
  ATTN_TRACE(warp, 60, 5);
  tc_fence_before();
  __syncthreads();
#ifdef YSI_ATTN_TRACE
  if (!GLOBAL && trace && threadIdx.x == 0) {
    const long long t0 = t_entry;
    for (int w = 0; w < 4; ++w) {
      printf("TW warp %d: setup %lld tab %lld bias %lld | loop_end %lld pv_done %lld o_loaded %lld staged %lld stored %lld | end %lld\n", w, g_attn_trace[w][60][0] - t0, g_attn_trace[w][60][1] - t0,
             g_attn_trace[w][60][2] - t0, g_attn_trace[w][60][3] - t0, g_attn_trace[w][60][4] - t0, g_attn_trace[w][60][6] - t0, g_attn_trace[w][60][7] - t0, g_attn_trace[w][60][5] - t0, clock64() - t0);
      for (int j = 0; j < 3; ++j)
        printf("TW   tile %d: start %lld s_full %lld ld %lld fin %lld max %lld exps %lld pre_st %lld done %lld\n", j, g_attn_trace[w][j][0] - t0, g_attn_trace[w][j][1] - t0,
               g_attn_trace[w][j][2] - t0, g_attn_trace[w][j][3] - t0, g_attn_trace[w][j][7] - t0, g_attn_trace[w][j][4] - t0, g_attn_trace[w][j][5] - t0, g_attn_trace[w][j][6] - t0);
    }
    for (int j = 0; j < 3; ++j)
      printf("TW tile %d S-issue: kfull %lld free %lld issued %lld | PV: vfull %lld p_full %lld issued %lld\n", j, g_attn_trace[9][j][0] - t0, g_attn_trace[9][j][1] - t0,
             g_attn_trace[9][j][2] - t0, g_attn_trace[10][j][0] - t0, g_attn_trace[10][j][1] - t0, g_attn_trace[10][j][2] - t0);
  }
  if (GLOBAL && trace && threadIdx.x == 0) {
    const long long t0 = g_attn_trace[0][20][0];
    for (int j = 20; j < 23; ++j) {
      for (int w = 0; w < 8; ++w)
        printf("TR tile %d warp %d: start %lld s_full %lld ld %lld exps %lld pair %lld p_free %lld done %lld\n", j, w, g_attn_trace[w][j][0] - t0,
               g_attn_trace[w][j][1] - t0, g_attn_trace[w][j][2] - t0, g_attn_trace[w][j][3] - t0, g_attn_trace[w][j][4] - t0,
               g_attn_trace[w][j][5] - t0, g_attn_trace[w][j][6] - t0);
      printf("TR tile %d S-issue: kfull %lld s_free %lld issued %lld | PV: vfull %lld p_full %lld issued %lld | TMA: kempty %lld vempty %lld\n", j,
             g_attn_trace[9][j][0] - t0, g_attn_trace[9][j][1] - t0, g_attn_trace[9][j][2] - t0, g_attn_trace[10][j][0] - t0,
             g_attn_trace[10][j][1] - t0, g_attn_trace[10][j][2] - t0, g_attn_trace[8][j][0] - t0, g_attn_trace[8][j][1] - t0);
    }
  }
#endif
  if (warp == attn::SM_WARPS) tmem_dealloc(tmem_base, Cfg<GLOBAL, HD>::TMEM_COLS);
}

template <bool GLOBAL, int HD>
static void launch_attn_t(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const CUtensorMap& tmKVtail, const CUtensorMap& tmRel,
                          const CUtensorMap& tmRel8, const CUtensorMap& tmQ1, const CUtensorMap& tmKV1, const CUtensorMap& tmKVtail1,
                          const CUtensorMap& tmRel1, const CUtensorMap& tmRel8_1, const AttnParams& p, dim3 grid, cudaStream_t stream) {
  using C = attn::Cfg<GLOBAL, HD>;
  ensure_dyn_smem(reinterpret_cast<const void*>(encoder_attention_kernel<GLOBAL, HD>), C::SMEM_BYTES);
  encoder_attention_kernel<GLOBAL, HD><<<grid, attn::THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, tmQ1, tmKV1, tmKVtail1,
                                                                                      tmRel1, tmRel8_1, p);
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
  // head_dim 80: columns 64..79 as 16-column boxes (32-byte rows, 32B swizzle); head_dim 64 never touches them
  const uint32_t c1 = head_dim == 64 ? 64 : 16;
  const CUtensorMap tmQ1 = make_tmap_op16_2d(qkv, rows, 3 * D, 3 * D, BQ, c1);
  const CUtensorMap tmKV1 = make_tmap_op16_2d(qkv, rows, 3 * D, 3 * D, BKV, c1);
  const CUtensorMap tmKVtail1 = make_tmap_op16_2d(qkv, rows, 3 * D, 3 * D, 16, c1);
  const CUtensorMap tmRel1 = make_tmap_op16_2d(rel_tab, 256, HDP, HDP, is_global ? 128 : 32, c1);
  const CUtensorMap tmRel8_1 = make_tmap_op16_2d(rel_tab, 256, HDP, HDP, 8, c1);
  AttnParams p;
  p.T = T; p.D = D; p.out = out; p.unwindow = unwindow ? 1 : 0;
  dim3 grid(ceil_div(T, BQ), heads, n_seq);
  if (head_dim == 64) {
    if (is_global) launch_attn_t<true, 64>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, tmQ1, tmKV1, tmKVtail1, tmRel1, tmRel8_1, p, grid, stream);
    else launch_attn_t<false, 64>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, tmQ1, tmKV1, tmKVtail1, tmRel1, tmRel8_1, p, grid, stream);
  } else {
    if (is_global) launch_attn_t<true, 80>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, tmQ1, tmKV1, tmKVtail1, tmRel1, tmRel8_1, p, grid, stream);
    else launch_attn_t<false, 80>(tmQ, tmKV, tmKVtail, tmRel, tmRel8, tmQ1, tmKV1, tmKVtail1, tmRel1, tmRel8_1, p, grid, stream);
  }
  YSI_CUDA(cudaGetLastError());
}

}  // namespace ysi
