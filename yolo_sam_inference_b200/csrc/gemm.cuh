// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T
//   op16 operands (K-major, TMA-loaded into 128B-swizzled shared memory), fp32 accumulators in TMEM.
//   warp 0   : TMA producer (one elected lane)          -- STAGES-deep smem ring, full/empty mbarriers
//   warp 1   : TMEM allocator + MMA issuer (one lane)    -- tcgen05.mma 128 x BN x 16, commit -> mbarriers
//   warps 2-9: epilogue (TMEM -> registers -> global)    -- double-buffered accumulator (2 x BN columns);
//              warp w owns TMEM lane quarter w%4 and column half (w-2)/4, so two warps per SM sub-partition
//              keep loads / MUFU / stores of the epilogue in flight while the next tile's MMAs run
// The epilogue is a functor so the same mainloop serves the ViT linears (bias / GELU / residual add),
// the decoder projections and the fused ConvTranspose upscaler epilogues.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace ysi {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf-GELU on a pair of values (packed FFMA2 / FMUL2), erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7):
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z = |x| / sqrt(2)
//   gelu(x) = x/2 + |x/2| * erf(z)
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 d = fma2(ax, make_float2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), make_float2(1.f, 1.f));
  const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  float2 q = fma2(t, make_float2(-1.061405429f, -1.061405429f), make_float2(1.453152027f, 1.453152027f));
  q = fma2(q, t, make_float2(-1.421413741f, -1.421413741f));
  q = fma2(q, t, make_float2(0.284496736f, 0.284496736f));
  q = fma2(q, t, make_float2(-0.254829592f, -0.254829592f));
  q = mul2(q, t);                                             // = -(a1 t + ... + a5 t^5)
  float2 s = mul2(x, x);
  s = mul2(s, make_float2(-0.72134752044448170f, -0.72134752044448170f));   // -z^2 * log2(e) = -x^2/2 * log2(e)
  const float2 e = make_float2(ex2_approx(s.x), ex2_approx(s.y));
  const float2 erfv = fma2(q, e, make_float2(1.f, 1.f));
  const float2 hx = mul2(x, make_float2(0.5f, 0.5f));
  return fma2(make_float2(fabsf(hx.x), fabsf(hx.y)), erfv, hx);
}

// what an epilogue may fetch per tile BEFORE the accumulator is ready (CTA-pair kernel: Epi::prefetch), so that the loads'
// latency hides under the wait: the row's LayerNorm scalars
struct EpiPre {
  float2 row;        // EpiStaged: (rstd, -rstd * mean) of the thread's row; (0, 0): zero (pad) row
  uint4 slab[8];     // EpiResidLN: the warp's first 32-column slab of the addend (slab_load_issue layout)
};
// per-warp epilogue context: staging shared memory (CTA-pair kernel only) and the warp's first output row
struct EpiCtx {
  uint32_t smem;     // 2 x 4 KB staging buffers of this warp (0: none)
  uint32_t nbuf;     // running buffer counter
  int row0;          // first row of the warp's 32-row slab
  int lane;
  EpiPre pre;        // CTA-pair kernel: what Epi::prefetch() returned, loaded before the wait for the accumulator
};

// Per-epilogue kernel options of gemm_op16_kernel (specialise for an epilogue type):
//   WARP_SMEM : bytes of staging shared memory per epilogue warp, handed over in EpiCtx::smem
//   FULL_ROW  : an accumulator tile is drained by FOUR warps that own whole rows (all BN columns) instead of eight that own
//               half a row each: the warps of column half h take the tiles that land in accumulator buffer h, so both
//               groups work at the same time on consecutive tiles and a row-wise reduction (LayerNorm over the tile's
//               columns) needs no exchange between warps
template <class Epi>
struct EpiTraits {
  static constexpr bool FULL_ROW = false;
  static constexpr int WARP_SMEM = 0;
};

// Warp-collective transposition through a 4 KB shared-memory slab of 32 rows x 128 bytes (16-byte pieces XOR-swizzled by
// the row, conflict free both ways).  An epilogue thread owns one accumulator ROW, so its own loads / stores touch 32
// different 128-byte lines per instruction (16 useful bytes each); through the slab every global instruction moves four
// whole lines instead.  g_row0 = address of the slab's first row (row r of the slab lives pitch_bytes * r further).
__device__ __forceinline__ void slab_put(uint32_t sm, int lane, int j, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  // 16-byte piece j (0..7) of the calling lane's row
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sm + static_cast<uint32_t>(lane) * 128u +
                                                                ((static_cast<uint32_t>(j) ^ static_cast<uint32_t>(lane & 7)) << 4)),
               "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void slab_flush(uint32_t sm, int lane, void* g_row0, size_t pitch_bytes) {
  __syncwarp();
  uint8_t* g = static_cast<uint8_t*>(g_row0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t r = static_cast<uint32_t>(4 * j + (lane >> 3)), pc = static_cast<uint32_t>(lane & 7);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sm + r * 128u + ((pc ^ (r & 7u)) << 4)) : "memory");
    *reinterpret_cast<uint4*>(g + pitch_bytes * r + pc * 16u) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void slab_store(uint32_t sm, int lane, const uint32_t (&pk)[32], void* g_row0, size_t pitch_bytes) {
#pragma unroll
  for (int j = 0; j < 8; ++j) slab_put(sm, lane, j, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  slab_flush(sm, lane, g_row0, pitch_bytes);
}
// Half-width variant: 32 rows x 64 bytes in 2 KB (four 16-byte pieces per row, swizzled by row >> 1), for outputs whose
// rows are 64-byte segments; two of them fit the warp's 4 KB.
__device__ __forceinline__ void slab64_put(uint32_t sm, int lane, int p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sm + static_cast<uint32_t>(lane) * 64u +
                                                                ((static_cast<uint32_t>(p) ^ static_cast<uint32_t>((lane >> 1) & 3)) << 4)),
               "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void slab64_flush(uint32_t sm, int lane, void* g_row0, size_t pitch_bytes) {
  __syncwarp();
  uint8_t* g = static_cast<uint8_t*>(g_row0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t r = static_cast<uint32_t>(8 * j + (lane >> 2)), pc = static_cast<uint32_t>(lane & 3);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sm + r * 64u + ((pc ^ ((r >> 1) & 3u)) << 4)) : "memory");
    *reinterpret_cast<uint4*>(g + pitch_bytes * r + pc * 16u) = v;
  }
  __syncwarp();
}
// Loads come in two halves so that a caller can have the next slab's global loads in flight while it works on the current one.
__device__ __forceinline__ void slab_load_issue(int lane, const void* g_row0, size_t pitch_bytes, uint4 (&v)[8]) {
  const uint8_t* g = static_cast<const uint8_t*>(g_row0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t r = static_cast<uint32_t>(4 * j + (lane >> 3)), pc = static_cast<uint32_t>(lane & 7);
    v[j] = __ldg(reinterpret_cast<const uint4*>(g + pitch_bytes * r + pc * 16u));
  }
}
// same, with coherent loads (for memory this kernel also writes -- the residual stream -- where ld.global.nc is not allowed)
__device__ __forceinline__ void slab_load_issue_rw(int lane, const void* g_row0, size_t pitch_bytes, uint4 (&v)[8]) {
  const uint8_t* g = static_cast<const uint8_t*>(g_row0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t r = static_cast<uint32_t>(4 * j + (lane >> 3)), pc = static_cast<uint32_t>(lane & 7);
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                 : "l"(g + pitch_bytes * r + pc * 16u) : "memory");
  }
}
// 64-byte slab written to mapped rows: slab row r goes to row dst_rows[r] of the destination (null: row0_dst + r)
__device__ __forceinline__ void slab64_flush_map(uint32_t sm, int lane, void* g_base, size_t pitch_bytes, const int* dst_rows, int row0_dst) {
  __syncwarp();
  uint8_t* g = static_cast<uint8_t*>(g_base);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t r = static_cast<uint32_t>(8 * j + (lane >> 2)), pc = static_cast<uint32_t>(lane & 3);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sm + r * 64u + ((pc ^ ((r >> 1) & 3u)) << 4)) : "memory");
    const size_t drow = dst_rows ? static_cast<size_t>(__ldg(dst_rows + r)) : static_cast<size_t>(row0_dst) + r;
    *reinterpret_cast<uint4*>(g + pitch_bytes * drow + pc * 16u) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void slab_load_finish(uint32_t sm, int lane, const uint4 (&v)[8], uint32_t (&out)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t r = static_cast<uint32_t>(4 * j + (lane >> 3)), pc = static_cast<uint32_t>(lane & 7);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sm + r * 128u + ((pc ^ (r & 7u)) << 4)), "r"(v[j].x), "r"(v[j].y),
                 "r"(v[j].z), "r"(v[j].w)
                 : "memory");
  }
  __syncwarp();
  const uint32_t rowp = sm + static_cast<uint32_t>(lane) * 128u, swz = static_cast<uint32_t>(lane & 7);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(out[4 * j]), "=r"(out[4 * j + 1]), "=r"(out[4 * j + 2]), "=r"(out[4 * j + 3])
                 : "r"(rowp + ((static_cast<uint32_t>(j) ^ swz) << 4)) : "memory");
  __syncwarp();
}
__device__ __forceinline__ void slab_load(uint32_t sm, int lane, const void* g_row0, size_t pitch_bytes, uint32_t (&out)[32]) {
  uint4 v[8];
  slab_load_issue(lane, g_row0, pitch_bytes, v);
  slab_load_finish(sm, lane, v, out);
}

// Generic epilogue: v = acc + bias[col]; act; + add_src[(row % add_mod), col]; -> out_f32 (= or +=) / out_op16.
struct EpiGeneric {
  GemmEpilogue p;
  int reverse_m = 0;   // CTA-pair kernel: walk the row tiles last-to-first (see GemmEpilogue::reverse_m)
  __device__ __forceinline__ void finish(EpiCtx&) const {}
  __device__ __forceinline__ EpiPre prefetch(int, int, int, int) const { return EpiPre{}; }
  // columns [c_begin, c_end) of the BN-wide accumulator tile belong to the calling warp
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx&) const {
    // every lane must execute the warp-collective tcgen05.ld convergently: predicate only the stores
    int drow = -1;
    if (row < M) drow = p.row_map ? p.row_map[row] : row;
    const bool active = drow >= 0;
    const float* addp = nullptr;
    if (active && p.add_src) {
      const int arow = p.add_group ? p.add_group[row / p.add_mod] * p.add_mod + row % p.add_mod : row % p.add_mod;
      addp = p.add_src + static_cast<size_t>(arow) * p.ld_add;
    }
    for (int c = c_begin; c < c_end; c += 32) {
      const int col0 = n0 + c;
      if (col0 >= N) break;          // uniform; N is a multiple of 32 (checked on the host)
      uint32_t r[32];
      tmem_ld_x32(taddr_row + c, r);
      tmem_ld_wait();
      if (!active) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 b = __ldg(b4 + i);
          v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
      }
      if (col0 >= p.scale_c0 && col0 < p.scale_c1) {     // uniform per 32-column chunk
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= p.col_scale;
      }
      if (p.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 g = gelu_erf2(make_float2(v[i], v[i + 1]));
          v[i] = g.x; v[i + 1] = g.y;
        }
      } else if (p.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
      }
      if (addp) {
        const float4* a4 = reinterpret_cast<const float4*>(addp + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 a = __ldg(a4 + i);
          v[4 * i] += a.x; v[4 * i + 1] += a.y; v[4 * i + 2] += a.z; v[4 * i + 3] += a.w;
        }
      }
      if (p.out_f32) {
        float4* o4 = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(drow) * p.ld_out + col0);
        if (p.accumulate == 2) {
          // fire-and-forget fp32 adds executed by the L2 (one add per element, so the result is the same
          // round-to-nearest sum a load/add/store would give) -- no read latency on the epilogue's critical path
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o4 + i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                         "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                         : "memory");
          continue;
        }
        if (p.accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 o = o4[i];
            v[4 * i] += o.x; v[4 * i + 1] += o.y; v[4 * i + 2] += o.z; v[4 * i + 3] += o.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      if (p.out_op16) {
        uint4* o4 = reinterpret_cast<uint4*>(p.out_op16 + static_cast<size_t>(drow) * p.ld_out_op16 + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_op16x2(v[8 * i], v[8 * i + 1]);
          o.y = pack_op16x2(v[8 * i + 2], v[8 * i + 3]);
          o.z = pack_op16x2(v[8 * i + 4], v[8 * i + 5]);
          o.w = pack_op16x2(v[8 * i + 6], v[8 * i + 7]);
          o4[i] = o;
        }
      }
    }
  }
};

// Staged epilogue (CTA-pair kernel): each warp converts a 32-row x 128-byte slab of its accumulator quarter,
// writes it to 128B-swizzled shared memory (conflict free) and one lane hands it to the TMA:
//   op16 activations -> cp.async.bulk.tensor store (64 columns per slab)
//   fp32 residual    -> cp.reduce.async.bulk.tensor .add (32 columns per slab): x += acc + bias, added in the L2
// Global writes are full 128-byte rows instead of 32 scattered 16-byte pieces per instruction, and rows past M
// are clipped by the tensor map.
struct alignas(64) EpiStaged {
  CUtensorMap tm_out;
  const float* bias;
  int act;
  float col_scale;
  int scale_c0, scale_c1;
  int f32_add;     // 0: op16 store, 1: fp32 reduce-add, 2: fp32 store
  int reverse_m = 0;   // walk the row tiles last-to-first (GemmEpilogue::reverse_m)
  // LayerNorm folded into this GEMM (op16 store path; GemmEpilogue::ln_stats): the A operand is op16(gamma * x) of the raw
  // residual stream and  y = rstd * acc - rstd * mean * cs[n] + wb[n] + bias[n]  with the row's mean / rstd from its partial
  // sums; a zero (pad) row of a padded window gets the plain bias
  const float2* ln_stats = nullptr;      // [token rows][ln_np] partial (mean, sum of squared deviations) written by EpiResidLN
  const int* ln_rowmap = nullptr;        // GEMM row -> token row (< 0: zero pad row); null: identity
  int ln_np = 0;
  float ln_inv_d = 0.f, ln_eps = 0.f, ln_cols_per_slot = 0.f;
  const float *ln_cs = nullptr, *ln_bw = nullptr, *ln_wb = nullptr;      // cs, bias + wb (real rows), wb (subtracted again on pad rows)
  __device__ __forceinline__ EpiPre prefetch(int row, int M, int col0, int lane) const {
    EpiPre p{};
    if (!ln_stats) return p;
    const int tok = row < M ? (ln_rowmap ? __ldg(ln_rowmap + row) : row) : -1;
    if (tok >= 0) {
      // ln_np is even and the row 16-byte aligned: two slots per load. (Loading them with L1::no_allocate, to keep the L1 for the
      // per-column vectors of run(), was measured slower: qkv 1.17 -> 1.38 ms per batch -- the eight loads of a row then each
      // go to the L2 instead of hitting the line the first one brought in.)
      // slot i = (mean, sum of squared deviations) of the row over its i-th group of D / ln_np columns; merged pairwise-stably
      // (Chan et al.), always in slot order: bitwise reproducible, and no E[x^2] - mean^2 cancellation for rows whose mean is
      // large against their spread
      const float4* sp = reinterpret_cast<const float4*>(ln_stats + static_cast<size_t>(tok) * ln_np);
      const float m = ln_cols_per_slot;
      float mean = 0.f, m2 = 0.f, k = 0.f;
      for (int i = 0; i < ln_np / 2; ++i) {
        const float4 t = __ldg(sp + i);
        float d = t.x - mean, r = __fdividef(1.f, k + 1.f);
        mean += d * r; m2 += t.y + d * d * (m * k * r); k += 1.f;
        d = t.z - mean; r = __fdividef(1.f, k + 1.f);
        mean += d * r; m2 += t.w + d * d * (m * k * r); k += 1.f;
      }
      const float rstd = rsqrtf(m2 * ln_inv_d + ln_eps);
      p.row = make_float2(rstd, -rstd * mean);
    }
    return p;
  }
  // bulk async-groups belong to the issuing thread: elect.sync picks the same lane for the same (full) mask every time
  __device__ __forceinline__ void finish(EpiCtx& ctx) const {
    if (elect_one()) bulk_wait_read<0>();
  }
  __device__ __forceinline__ void slab_out(EpiCtx& ctx, const uint32_t (&pk)[32], int col0) const {
    const uint32_t buf = ctx.smem + (ctx.nbuf & 1u) * 4096u;
    if (elect_one()) bulk_wait_read<1>();       // the slab stored from this buffer two steps ago has been read
    __syncwarp();
    const uint32_t rowp = buf + static_cast<uint32_t>(ctx.lane) * 128u;
    const uint32_t swz = static_cast<uint32_t>(ctx.lane & 7);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((static_cast<uint32_t>(j) ^ swz) << 4)), "r"(pk[4 * j]),
                   "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                   : "memory");
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      if (f32_add == 1) tma_reduce_add_2d(&tm_out, buf, col0, ctx.row0); else tma_store_2d(&tm_out, buf, col0, ctx.row0);
      bulk_commit();
    }
    ++ctx.nbuf;
  }
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx& ctx) const {
    if (f32_add) {
      for (int c = c_begin; c < c_end; c += 32) {
        const int col0 = n0 + c;
        uint32_t r[32];
        tmem_ld_x32(taddr_row + c, r);
        tmem_ld_wait();
        if (bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + b.x);
            r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + b.y);
            r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + b.z);
            r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + b.w);
          }
        }
        slab_out(ctx, r, col0);
      }
    } else {
      // folded LayerNorm: this thread's row scalars (fetched by prefetch() before the accumulator was ready)
      const float rs = ctx.pre.row.x, nrm = ctx.pre.row.y;
      const bool pad = ln_stats && rs == 0.f;
      const bool any_pad = ln_stats && __any_sync(0xFFFFFFFFu, pad);
      for (int c = c_begin; c < c_end; c += 64) {
        const int col0 = n0 + c;
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_x32(taddr_row + c + 32 * h, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          const int colh = col0 + 32 * h;
          if (ln_stats) {
            // y = rs * acc + (nrm * cs[n] + bw[n]), two packed FFMA2 per pair of columns (bw = bias + wb, folded on the host).
            // (Broadcasting cs / bw from registers spread over the lanes -- 256 shuffles per tile and warp -- measured slower
            // than these loads, although they miss the ~24 KB of L1 this kernel leaves: qkv 1.23 vs 1.17 ms per batch.)
            const float4* c4 = reinterpret_cast<const float4*>(ln_cs + colh);
            const float4* w4 = reinterpret_cast<const float4*>(ln_bw + colh);
            const float2 rs2 = make_float2(rs, rs), nrm2 = make_float2(nrm, nrm);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 cc = __ldg(c4 + i), ww = __ldg(w4 + i);
              const float2 a0 = fma2(rs2, make_float2(v[4 * i], v[4 * i + 1]), fma2(nrm2, make_float2(cc.x, cc.y), make_float2(ww.x, ww.y)));
              const float2 a1 = fma2(rs2, make_float2(v[4 * i + 2], v[4 * i + 3]), fma2(nrm2, make_float2(cc.z, cc.w), make_float2(ww.z, ww.w)));
              v[4 * i] = a0.x; v[4 * i + 1] = a0.y; v[4 * i + 2] = a1.x; v[4 * i + 3] = a1.y;
            }
            if (any_pad && pad) {          // zero row of a padded window: the plain bias
              const float4* p4 = reinterpret_cast<const float4*>(ln_wb + colh);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 ww = __ldg(p4 + i);
                v[4 * i] -= ww.x; v[4 * i + 1] -= ww.y; v[4 * i + 2] -= ww.z; v[4 * i + 3] -= ww.w;
              }
            }
          } else if (bias) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + colh);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (colh >= scale_c0 && colh < scale_c1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= col_scale;
          }
          if (act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 g = gelu_erf2(make_float2(v[i], v[i + 1]));
              v[i] = g.x; v[i + 1] = g.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[16 * h + i] = pack_op16x2(v[2 * i], v[2 * i + 1]);
        }
        slab_out(ctx, pk, col0);
      }
    }
  }
};

// Residual add with the NEXT LayerNorm folded in (CTA-pair kernel): x_new = (acc + bias) + x is formed on the SM (coalesced
// slab loads / stores of the fp32 stream instead of the L2 reduce-add), and while the values are in registers the epilogue
// also emits what the following qkv / fc1 GEMM needs in place of a LayerNorm pass over x:
//   x16   = op16(gamma * x_new), written to row rowmap[row] (window-partition order for windowed layers) or row itself
//   stats = per-row (mean, sum of squared deviations) of x_new over this warp's columns, one slot per (N tile, column half)
// The consumer GEMM (EpiStaged::prefetch) turns them into mean / rstd.  x_new is bit-identical to the reduce-add path.
struct alignas(64) EpiResidLN {
  float* x; int ld;
  const float* res = nullptr; int ld_res = 0, res_mod = 1;   // optional: the addend comes from res[row % res_mod] instead of x (patch embed: + pos_embed)
  const float* bias;
  const float* gamma;
  op16* x16; int ld16;
  const int* rowmap;
  float2* stats; int np;
  int reverse_m = 0;
  __device__ __forceinline__ void finish(EpiCtx&) const {}
  // the first slab of the addend is requested before the wait for the accumulator: its HBM round trip (which the L2
  // reduce-add of the plain residual epilogue never sees) hides under the tile's mainloop
  __device__ __forceinline__ EpiPre prefetch(int row, int, int col0, int lane) const {
    EpiPre p;
    p.row = make_float2(0.f, 0.f);
    const int row0 = row - lane;
    const float* lb = res ? res + static_cast<size_t>(row0 % res_mod) * ld_res : x + static_cast<size_t>(row0) * ld;
    slab_load_issue_rw(lane, lb + col0, static_cast<size_t>(res ? ld_res : ld) * sizeof(float), p.slab);
    return p;
  }
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx& ctx) const {
    const int lane = ctx.lane, row0 = ctx.row0;        // M is a multiple of 32: the warp's 32 rows all exist
    float* xb = x + static_cast<size_t>(row0) * ld + n0;
    const float* lb = res ? res + static_cast<size_t>(row0 % res_mod) * ld_res + n0 : xb;      // res_mod is a multiple of 32
    const size_t lpitch = static_cast<size_t>(res ? ld_res : ld) * sizeof(float);
    const uint32_t buf0 = ctx.smem, buf1 = ctx.smem + 4096u;
    float cnt = 0.f, mean = 0.f, m2 = 0.f;      // running (count, mean, sum of squared deviations) of this row over the warp's columns
    uint4 nxt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) nxt[j] = ctx.pre.slab[j];
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t a[32], r[32];
      uint4 cur[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      if (c + 32 < c_end) slab_load_issue_rw(lane, lb + c + 32, lpitch, nxt);
      tmem_ld_x32(taddr_row + c, a);
      slab_load_finish(buf0, lane, cur, r);
      tmem_ld_wait();
      const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = __ldg(b4 + i);
        const float v0 = (__uint_as_float(a[4 * i]) + bb.x) + __uint_as_float(r[4 * i]);
        const float v1 = (__uint_as_float(a[4 * i + 1]) + bb.y) + __uint_as_float(r[4 * i + 1]);
        const float v2 = (__uint_as_float(a[4 * i + 2]) + bb.z) + __uint_as_float(r[4 * i + 2]);
        const float v3 = (__uint_as_float(a[4 * i + 3]) + bb.w) + __uint_as_float(r[4 * i + 3]);
        a[4 * i] = __float_as_uint(v0); a[4 * i + 1] = __float_as_uint(v1);
        a[4 * i + 2] = __float_as_uint(v2); a[4 * i + 3] = __float_as_uint(v3);
      }
      {
        // statistics of these 32 values about their own mean, merged into the running ones (Chan et al.): stable whatever the
        // row's mean is
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          s0 += __uint_as_float(a[i]); s1 += __uint_as_float(a[i + 1]); s2 += __uint_as_float(a[i + 2]); s3 += __uint_as_float(a[i + 3]);
        }
        const float cm = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
        s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float d0 = __uint_as_float(a[i]) - cm, d1 = __uint_as_float(a[i + 1]) - cm;
          const float d2 = __uint_as_float(a[i + 2]) - cm, d3 = __uint_as_float(a[i + 3]) - cm;
          s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
        }
        const float d = cm - mean, tot = cnt + 32.f;
        mean += d * (32.f / tot);
        m2 += ((s0 + s1) + (s2 + s3)) + d * d * (cnt * 32.f / tot);
        cnt = tot;
      }
      slab_store(buf0, lane, a, xb + c, static_cast<size_t>(ld) * sizeof(float));
      if (x16) {
        const float4* g4 = reinterpret_cast<const float4*>(gamma + n0 + c);
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const float4 g0 = __ldg(g4 + 2 * p4), g1 = __ldg(g4 + 2 * p4 + 1);
          slab64_put(buf1, lane, p4, pack_op16x2(g0.x * __uint_as_float(a[8 * p4]), g0.y * __uint_as_float(a[8 * p4 + 1])),
                     pack_op16x2(g0.z * __uint_as_float(a[8 * p4 + 2]), g0.w * __uint_as_float(a[8 * p4 + 3])),
                     pack_op16x2(g1.x * __uint_as_float(a[8 * p4 + 4]), g1.y * __uint_as_float(a[8 * p4 + 5])),
                     pack_op16x2(g1.z * __uint_as_float(a[8 * p4 + 6]), g1.w * __uint_as_float(a[8 * p4 + 7])));
        }
        slab64_flush_map(buf1, lane, x16 + n0 + c, static_cast<size_t>(ld16) * sizeof(op16), rowmap ? rowmap + row0 : nullptr, row0);
      }
    }
    if (stats) stats[static_cast<size_t>(row0 + lane) * np + (n0 + c_begin) / (c_end - c_begin)] = make_float2(mean, m2);
  }
};

// measurement-only epilogue (ysi_gemm_bench): drains the accumulator from TMEM and stores nothing unless the value
// is an (impossible) sentinel, so the mainloop can be timed without the epilogue's global-memory traffic
struct EpiDrain {
  float* sink;
  int reverse_m = 0;
  __device__ __forceinline__ void finish(EpiCtx&) const {}
  __device__ __forceinline__ EpiPre prefetch(int, int, int, int) const { return EpiPre{}; }
  __device__ __forceinline__ void run(uint32_t taddr_row, int row, int M, int n0, int N, int c_begin, int c_end, EpiCtx&) const {
    float acc = 0.f;
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr_row + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
    }
    if (acc == 1.2345678e33f) sink[0] = acc;
  }
};

template <int BN, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_op16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 Epi epi) {
  using Cfg = GemmCfg<BN>;
  constexpr bool FULL_ROW = EpiTraits<Epi>::FULL_ROW;
  constexpr int WARP_SMEM = EpiTraits<Epi>::WARP_SMEM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_smem = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;       // GEMM_EPI_WARPS x WARP_SMEM of epilogue staging
  const uint32_t bar_base = epi_smem + GEMM_EPI_WARPS * WARP_SMEM;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), FULL_ROW ? GEMM_EPI_WARPS / 2 : GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // The two single-thread roles run with their whole warp converged and one ELECTED lane issuing: the compiler then
  // emits the uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR) back to back instead of wrapping each one in
  // a per-lane waterfall loop, which otherwise makes the issuing thread -- not the tensor pipe -- the pacemaker.
  if (warp == 0) {
    const bool lead = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * GEMM_BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (lead) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            tma_load_2d(sa, &tmA, full_bar(stage), kb * GEMM_BK, m0);
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, n0);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const bool lead = elect_one();
    {
      constexpr uint32_t idesc = umma_idesc_op16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (lead) {
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // +32 B per 16-element K step inside the 128 B swizzle atom (encoded >> 4 => +2)
              umma_op16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (lead) umma_commit(tfull_bar(acc));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    const int c_begin = FULL_ROW ? 0 : half * (BN / 2), c_end = FULL_ROW ? BN : c_begin + BN / 2;
    const uint32_t my_smem = WARP_SMEM ? epi_smem + static_cast<uint32_t>(warp - 2) * WARP_SMEM : 0u;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if (!FULL_ROW || acc == half) {      // FULL_ROW: this warp group owns the tiles of accumulator buffer `half`
        const int m0 = (tile / num_n) * GEMM_BM;
        const int n0 = (tile % num_n) * BN;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(q * 32) << 16);
        EpiCtx ctx{my_smem, 0u, m0 + q * 32, lane, EpiPre{}};
        epi.run(taddr, m0 + q * 32 + lane, M, n0, N, c_begin, c_end, ctx);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant: cluster of 2 CTAs, tcgen05.mma.cta_group::2 with M = 256 (128 rows per CTA), N = 256.
// Each CTA stages its own 128 x 64 slice of A and HALF of the 256 x 64 B tile per k-block (32 KB / stage
// instead of 48 KB), so the ring is 6 deep in the same shared memory -- enough bytes in flight to cover
// HBM latency at full MMA rate -- and the tensor core reads half as much shared memory per flop.
//   leader CTA (rank 0): arms the full barriers with the bytes of BOTH CTAs, issues every MMA, commits
//                        (multicast) to the empty / accumulator-full barriers of both CTAs
//   both CTAs          : TMA producer for their own slices (completion counted on the leader's barrier),
//                        8 epilogue warps on their own 128 accumulator lanes, which release the accumulator
//                        on the leader's barrier
// ------------------------------------------------------------------------------------------------
// BN_ = 256, or 192 where that fills the last wave better (N = 768 at 8 images: 384 tiles on 74 CTA pairs are 5.2 waves,
// 512 tiles of 192 columns are 6.9); 192 only with the fp32 epilogues (their 32-column slabs divide the 96-column half).
template <int BN_>
struct Gemm2Cfg {
  static constexpr int BN = BN_;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;        // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;       // 16 / 12 KB: this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 5;
  static constexpr int EPI_STAGE_BYTES = GEMM_EPI_WARPS * 2 * 4096;   // two 32-row x 128-byte slabs per epilogue warp
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
};

// AMODE 1: the A operand is the implicit im2col of a 3x3 / pad-1 convolution over a [images, 64, 64, 512] op16 map (tmA is
// the 4-D tensor map of it, box = 64 channels x 64 x x 2 y): k-block kb = tap (kb / 8) x 64-channel slice (kb % 8), the 128
// rows of a CTA are two rows of the 64 x 64 grid, and the tile of tap (ky, kx) is the box shifted by (ky - 1, kx - 1) with
// the TMA's zero fill as the padding -- the [tokens, 9 x 512] im2col matrix is never written.
template <int BN_, class Epi, int AMODE = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_op16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                  const __grid_constant__ Epi epi) {
  using Cfg = Gemm2Cfg<BN_>;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_smem = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_smem + Cfg::EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);       // leader: arrive.expect_tx (bytes of both CTAs)
      mbar_init(empty_bar(s), 1);      // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);                      // multicast commit
      mbar_init(tempty_bar(a), 2 * GEMM_EPI_WARPS);    // leader's copy collects the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    const bool lead = elect_one();      // converged warp + elected lane, see gemm_op16_kernel
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mt = epi.reverse_m ? num_m - 1 - tile / num_n : tile / num_n;      // optional last-to-first row-tile order
        const int m0 = mt * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
        const int n0 = (tile % num_n) * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (lead) {
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            if (AMODE == 1) {
              const int tap = kb >> 3;
              tma_load_4d_cg2(sa, &tmA, full_bar(stage), (kb & 7) * GEMM_BK, tap % 3 - 1, ((m0 & 4095) >> 6) + tap / 3 - 1, m0 >> 12);
            } else {
              tma_load_2d_cg2(sa, &tmA, full_bar(stage), kb * GEMM_BK, m0);
            }
            tma_load_2d_cg2(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, n0);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const bool lead = elect_one();
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_op16(2 * GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (lead) {
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128(sa, 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_op16_ss_cg2(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2_mc(empty_bar(stage), 3);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (lead) umma_commit_cg2_mc(tfull_bar(acc), 3);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    EpiCtx ctx{epi_smem + static_cast<uint32_t>(warp - 2) * 8192u, 0u, 0, lane, EpiPre{}};
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int mt = epi.reverse_m ? num_m - 1 - tile / num_n : tile / num_n;
      const int m0 = mt * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
      const int n0 = (tile % num_n) * BN;
      ctx.row0 = m0 + q * 32;
      ctx.pre = epi.prefetch(m0 + q * 32 + lane, M, n0 + c_begin, lane);      // per-row / per-column scalars: latency hides under the wait
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(q * 32) << 16);
      epi.run(taddr, m0 + q * 32 + lane, M, n0, N, c_begin, c_end, ctx);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    epi.finish(ctx);
  }
  tc_fence_before();
  cluster_sync_all();      // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
}

template <int BN_ = 256, int AMODE = 0, class Epi>
void launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epi& epi, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN_>;
  auto kern = gemm2_op16_kernel<BN_, Epi, AMODE>;
  ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES);
  const int tiles = ceil_div(M, 2 * GEMM_BM) * (N / Cfg::BN);
  const int pairs = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
  kern<<<2 * pairs, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, epi);
  YSI_CUDA(cudaGetLastError());
}

template <int BN, class Epi>
void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epi& epi,
                 cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  constexpr int SMEM = Cfg::SMEM_BYTES + GEMM_EPI_WARPS * EpiTraits<Epi>::WARP_SMEM;
  static_assert(SMEM <= 232448, "shared memory of the GEMM stages + epilogue staging exceeds one SM");
  static_assert(EpiTraits<Epi>::WARP_SMEM % 1024 == 0, "epilogue staging slabs are 1 KB aligned");
  auto kern = gemm_op16_kernel<BN, Epi>;
  ensure_dyn_smem(reinterpret_cast<const void*>(kern), SMEM);
  const int tiles = ceil_div(M, GEMM_BM) * ceil_div(N, BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, SMEM, stream>>>(tmA, tmB, M, N, K, epi);
  YSI_CUDA(cudaGetLastError());
}

}  // namespace ysi
