// Host side of the tcgen05 GEMM: tensor-map construction and tile-shape dispatch.
#include "gemm.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace ysi {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    YSI_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    YSI_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static CUtensorMap make_tmap_2d(const void* base, int esize, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                                uint32_t box_cols);

CUtensorMap make_tmap_op16_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                              uint32_t box_cols) {
  return make_tmap_2d(base, 2, rows, cols, ld, box_rows, box_cols);
}
CUtensorMap make_tmap_f32_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return make_tmap_2d(base, 4, rows, cols, ld, box_rows, 32);
}

static CUtensorMap make_tmap_2d(const void* base, int esize, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                                uint32_t box_cols) {
  // cached: the context's buffers are fixed, so the same few dozen maps are requested every step
  typedef std::tuple<const void*, int, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key(base, esize, rows, cols, ld, box_rows, box_cols);
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  YSI_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  YSI_CHECK((ld * esize) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  YSI_CHECK(box_cols * esize == 128 || box_cols * esize == 32, "inner box must be 128 bytes (128B swizzle) or 32 bytes (32B swizzle)");
  const CUtensorMapSwizzle swz = box_cols * esize == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  YSI_CHECK(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  CUtensorMap m;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * static_cast<uint64_t>(esize)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(&m, esize == 2 ? OP16_TMAP_TYPE : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YSI_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  std::lock_guard<std::mutex> g(mu);
  cache[key] = m;
  return m;
}

// [images, 64, 64, channels] op16 map (token-major activations of the 64 x 64 grid) as a 4-D tensor map with a
// {64 channels, 64 x, 2 y, 1 image} box = the 128 x 64 A tile of the GEMM; out-of-range x / y read as zero (conv padding).
static CUtensorMap make_tmap_op16_grid4d(const void* base, int n_images, int channels) {
  typedef std::tuple<const void*, int, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key(base, n_images, channels);
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  YSI_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && channels % 64 == 0, "grid map: alignment / channel count");
  CUtensorMap m;
  const cuuint64_t row = static_cast<cuuint64_t>(channels) * 2;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(channels), 64, 64, static_cast<cuuint64_t>(n_images)};
  cuuint64_t gstride[3] = {row, 64 * row, 4096 * row};
  cuuint32_t box[4] = {64, 64, 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode_fn()(&m, OP16_TMAP_TYPE, 4, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YSI_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D) failed with code " + std::to_string(static_cast<int>(r)));
  std::lock_guard<std::mutex> g(mu);
  cache[key] = m;
  return m;
}

void gemm_conv3x3_grid(const op16* in, int n_images, int channels, const op16* W, float* out_f32, int N, cudaStream_t stream) {
  YSI_CHECK(channels == 512 && N == 256, "implicit 3x3 convolution GEMM: 512 input values per position (8 k-blocks per tap), 256 outputs");
  const int M = n_images * 4096, K = 9 * channels;
  const CUtensorMap tmA = make_tmap_op16_grid4d(in, n_images, channels);
  const CUtensorMap tmB = make_tmap_op16_2d(W, N, K, K, 128);
  EpiStaged es;
  es.tm_out = make_tmap_f32_2d(out_f32, M, N, N, 32);
  es.bias = nullptr; es.act = ACT_NONE; es.col_scale = 1.f; es.scale_c0 = es.scale_c1 = 0;
  es.f32_add = 2; es.reverse_m = 0;
  launch_gemm2<256, 1>(tmA, tmB, M, N, K, es, stream);
}

const char* kernel_class_name(int kc) {
  static const char* names[KC_COUNT] = {"preprocess", "layernorm", "gemm_patch", "gemm_qkv", "attn_window", "attn_global",
                                        "gemm_proj", "gemm_fc1", "gemm_fc2", "neck", "dec_token", "dec_gemm", "dec_attn",
                                        "dec_upscale", "post_upsample", "post_hull"};
  return (kc >= 0 && kc < KC_COUNT) ? names[kc] : "?";
}

int Profiler::begin(int kc, double flops, double bytes) {
  while (static_cast<int>(pool.size()) < next + 2) {
    cudaEvent_t e;
    YSI_CUDA(cudaEventCreate(&e));
    pool.push_back(e);
  }
  Rec r{kc, next, next + 1, flops, bytes};
  next += 2;
  YSI_CUDA(cudaEventRecord(pool[r.e0], stream));
  recs.push_back(r);
  return static_cast<int>(recs.size()) - 1;
}
void Profiler::end(int rec) { YSI_CUDA(cudaEventRecord(pool[recs[rec].e1], stream)); }
void Profiler::collect(double* ms, long long* launches, double* flops, double* bytes) {
  for (int i = 0; i < KC_COUNT; ++i) { ms[i] = 0; launches[i] = 0; flops[i] = 0; bytes[i] = 0; }
  for (const Rec& r : recs) {
    float t = 0.f;
    YSI_CUDA(cudaEventElapsedTime(&t, pool[r.e0], pool[r.e1]));
    ms[r.kc] += t; launches[r.kc] += 1; flops[r.kc] += r.flops; bytes[r.kc] += r.bytes;
  }
}
void Profiler::reset() { recs.clear(); next = 0; }
Profiler::~Profiler() { for (auto e : pool) cudaEventDestroy(e); }

int sm_count() {
  static std::mutex mu;
  static std::map<int, int> per_dev;
  int dev = 0;
  YSI_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  auto it = per_dev.find(dev);
  if (it != per_dev.end()) return it->second;
  int n = 0;
  YSI_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  per_dev[dev] = n;
  return n;
}

void ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> granted;     // (device, kernel) -> bytes opted in
  int dev = 0;
  YSI_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  int& have = granted[std::make_pair(dev, func)];
  if (bytes <= have) return;
  YSI_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have = bytes;
}

// tile width the CTA-pair kernel uses for an fp32-output GEMM of this shape (192-column tiles where they need fewer rounds)
static bool pair_uses_bn192(int M, int N) {
  static const int allow192 = [] { const char* e = getenv("YSI_GEMM_BN192"); return e ? atoi(e) : 1; }();
  const int pairs = sm_count() / 2, m_tiles = ceil_div(M, 2 * GEMM_BM);
  return allow192 && N % 192 == 0 && ceil_div(m_tiles * (N / 192), pairs) * 192 < ceil_div(m_tiles * (N / 256), pairs) * 256;
}
int gemm_ln_stat_slots(int M, int N) { return 2 * (pair_uses_bn192(M, N) ? N / 192 : N / 256); }

void gemm_op16(const op16* A, int lda, const op16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
               cudaStream_t stream) {
  YSI_CHECK(M > 0 && N > 0 && K > 0, "empty GEMM");
  YSI_CHECK(N % 32 == 0, "GEMM N must be a multiple of 32");
  YSI_CHECK(K % 8 == 0, "GEMM K must be a multiple of 8");
  EpiGeneric epi{ep};
  epi.reverse_m = ep.reverse_m;
  const CUtensorMap tmA = make_tmap_op16_2d(A, M, K, lda, GEMM_BM);
  static const int use_pair = [] { const char* e = getenv("YSI_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (use_pair && N % 256 == 0 && M >= 2048) {
    // big GEMMs: CTA pairs (cta_group::2), each CTA loads half of the B tile
    const CUtensorMap tmB = make_tmap_op16_2d(W, N, K, ldw, 128);
    // 192-column tiles where they need less time than 256-column ones: rounds over the CTA pairs x columns per tile
    const bool bn192 = pair_uses_bn192(M, N);
    const CUtensorMap tmB192 = bn192 ? make_tmap_op16_2d(W, N, K, ldw, 96) : tmB;
    static const int use_staged = [] { const char* e = getenv("YSI_GEMM_STAGED"); return e ? atoi(e) : 1; }();
    const bool plain = !ep.row_map && !ep.add_src && ep.act != ACT_RELU;
    YSI_CHECK(!ep.ln_stats || (plain && ep.out_op16 && !ep.out_f32 && ep.ln_cs && ep.ln_wb && ep.ln_bw && ep.ln_np > 0 && ep.ln_dim > 0),
              "folded LayerNorm needs the staged op16 epilogue");
    if (ep.stats_out) {
      // residual add with the next LayerNorm's operand copy and statistics (EpiResidLN)
      const bool addend = ep.add_src && !ep.add_group && !ep.accumulate && ep.add_mod % 32 == 0;      // x = acc + bias + add_src[row % add_mod]
      YSI_CHECK(!ep.row_map && ep.act != ACT_RELU && (addend || (!ep.add_src && ep.accumulate)) && ep.out_f32 && !ep.out_op16 &&
                ep.act == ACT_NONE && M % 32 == 0 && ep.bias, "LayerNorm-producing epilogue: residual add (or bias + row-periodic addend) only");
      EpiResidLN er;
      er.x = ep.out_f32; er.ld = ep.ld_out; er.bias = ep.bias;
      if (addend) { er.res = ep.add_src; er.ld_res = ep.ld_add; er.res_mod = ep.add_mod; } er.gamma = ep.x16_gamma; er.x16 = ep.x16_out; er.ld16 = ep.ld_x16;
      er.rowmap = ep.x16_rowmap; er.stats = ep.stats_out; er.np = gemm_ln_stat_slots(M, N); er.reverse_m = ep.reverse_m;
      if (bn192) launch_gemm2<192>(tmA, tmB192, M, N, K, er, stream);
      else launch_gemm2(tmA, tmB, M, N, K, er, stream);
    } else if (use_staged && plain && ep.out_op16 && !ep.out_f32) {
      // op16 activation output: tile staged in shared memory, written with TMA stores
      EpiStaged es;
      es.tm_out = make_tmap_op16_2d(ep.out_op16, M, N, ep.ld_out_op16, 32);
      es.bias = ep.bias; es.act = ep.act; es.col_scale = ep.col_scale; es.scale_c0 = ep.scale_c0; es.scale_c1 = ep.scale_c1;
      es.f32_add = 0; es.reverse_m = ep.reverse_m;
      es.ln_stats = ep.ln_stats; es.ln_rowmap = ep.ln_rowmap; es.ln_np = ep.ln_np; es.ln_eps = ep.ln_eps;
      es.ln_inv_d = ep.ln_dim > 0 ? 1.0f / static_cast<float>(ep.ln_dim) : 0.f; es.ln_cs = ep.ln_cs; es.ln_bw = ep.ln_bw; es.ln_wb = ep.ln_wb;
      es.ln_cols_per_slot = ep.ln_np > 0 ? static_cast<float>(ep.ln_dim) / static_cast<float>(ep.ln_np) : 0.f;
      launch_gemm2(tmA, tmB, M, N, K, es, stream);
    } else if (use_staged && plain && ep.out_f32 && !ep.out_op16 && ep.accumulate && ep.act == ACT_NONE) {
      // residual add: x += tile through cp.reduce.async.bulk (fp32 add in the L2, 128-byte rows)
      EpiStaged es;
      es.tm_out = make_tmap_f32_2d(ep.out_f32, M, N, ep.ld_out, 32);
      es.bias = ep.bias; es.act = ACT_NONE; es.col_scale = 1.f; es.scale_c0 = es.scale_c1 = 0;
      es.f32_add = 1; es.reverse_m = ep.reverse_m;
      if (bn192) launch_gemm2<192>(tmA, tmB192, M, N, K, es, stream);
      else launch_gemm2(tmA, tmB, M, N, K, es, stream);
    } else if (use_staged && plain && ep.out_f32 && !ep.out_op16 && !ep.accumulate && ep.act == ACT_NONE && ep.scale_c1 <= ep.scale_c0) {
      // plain fp32 output (decoder image-side projections): staged 128-byte rows, TMA store
      EpiStaged es;
      es.tm_out = make_tmap_f32_2d(ep.out_f32, M, N, ep.ld_out, 32);
      es.bias = ep.bias; es.act = ACT_NONE; es.col_scale = 1.f; es.scale_c0 = es.scale_c1 = 0;
      es.f32_add = 2; es.reverse_m = ep.reverse_m;
      if (bn192) launch_gemm2<192>(tmA, tmB192, M, N, K, es, stream);
      else launch_gemm2(tmA, tmB, M, N, K, es, stream);
    } else {
      YSI_CHECK(!ep.ln_stats, "folded LayerNorm needs the staged op16 epilogue (YSI_GEMM_STAGED=1)");
      launch_gemm2(tmA, tmB, M, N, K, epi, stream);
    }
    return;
  }
  YSI_CHECK(!ep.ln_stats && !ep.stats_out, "folded LayerNorm is implemented in the CTA-pair kernel only (M >= 2048, N % 256 == 0)");
  // widest tile that does not waste more than a quarter of its columns
  if (ep.narrow_tiles && N % 64 == 0) {
    const CUtensorMap tmB = make_tmap_op16_2d(W, N, K, ldw, 64);
    launch_gemm<64>(tmA, tmB, M, N, K, epi, stream);
  } else if (N % 256 == 0 || N > 512) {
    const CUtensorMap tmB = make_tmap_op16_2d(W, N, K, ldw, 256);
    launch_gemm<256>(tmA, tmB, M, N, K, epi, stream);
  } else if (N % 128 == 0) {
    const CUtensorMap tmB = make_tmap_op16_2d(W, N, K, ldw, 128);
    launch_gemm<128>(tmA, tmB, M, N, K, epi, stream);
  } else {
    const CUtensorMap tmB = make_tmap_op16_2d(W, N, K, ldw, 64);
    launch_gemm<64>(tmA, tmB, M, N, K, epi, stream);
  }
}

}  // namespace ysi
