// Host-side helpers shared by every translation unit of libysi.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

namespace ysi {

// The 16-bit operand type of every tensor-core contraction (tcgen05 kind::f16 takes either encoding at the same
// rate; accumulation is always fp32 in TMEM).  Chosen at compile time: build.py emits one library per encoding.
//   bf16 (default)          : 8 significand bits  -- the north-star's nominal dtype
//   fp16 (-DYSI_OP_FP16=1)  : 11 significand bits -- 8x less operand rounding; activations of this path (LayerNorm
//                             outputs, q/k/v, GELU hidden, softmax P <= 2^8, normalised pixels) sit far inside
//                             fp16's range and the residual stream / statistics stay fp32
#ifdef YSI_OP_FP16
typedef __half op16;
#define YSI_OP_NAME "fp16"
#define OP16_TMAP_TYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
__host__ __device__ inline op16 f2op(float v) { return __float2half_rn(v); }
__host__ __device__ inline float op2f(op16 v) { return __half2float(v); }
#else
typedef __nv_bfloat16 op16;
#define YSI_OP_NAME "bf16"
#define OP16_TMAP_TYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
__host__ __device__ inline op16 f2op(float v) { return __float2bfloat16(v); }
__host__ __device__ inline float op2f(op16 v) { return __bfloat162float(v); }
#endif

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define YSI_CUDA(expr)                                                                               \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw ::ysi::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +    \
                             __FILE__ + ":" + std::to_string(__LINE__));                             \
  } while (0)

#define YSI_CHECK(cond, msg)                                                                         \
  do {                                                                                               \
    if (!(cond))                                                                                     \
      throw ::ysi::CudaError(std::string("check failed: ") + #cond + ": " + (msg) + " at " +         \
                             __FILE__ + ":" + std::to_string(__LINE__));                             \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Row-major op16 matrix [rows, cols] (row pitch ld elements) -> 2-D TMA map with a {box_cols, box_rows}
// box: box_cols = 64 (128-byte rows, 128B swizzle) or 16 (32-byte rows, 32B swizzle). Out-of-bounds reads return zero.
CUtensorMap make_tmap_op16_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                              uint32_t box_cols = 64);

// Row-major fp32 matrix -> 2-D TMA map with a {32 cols, box_rows} box (128 B inner extent), 128-byte swizzle;
// used by the staged GEMM epilogue for cp.reduce.async.bulk (x += tile).
CUtensorMap make_tmap_f32_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// Epilogue description for the tcgen05 GEMM (see gemm.cu).
enum GemmAct { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct GemmEpilogue {
  const float* bias = nullptr;   // [N] added to the accumulator
  int act = ACT_NONE;            // applied after bias
  const float* add_src = nullptr;  // optional fp32 [add_mod, ld_add] added after the activation
  int add_mod = 1;
  int ld_add = 0;
  const int* add_group = nullptr;  // optional: add row = add_group[row / add_mod] * add_mod + row % add_mod
  const int* row_map = nullptr;  // optional destination row per GEMM row (<0: drop the row)
  float* out_f32 = nullptr;      // optional fp32 destination [*, ld_out]
  int accumulate = 0;            // 1: out_f32 += value (load/add/store); 2: same sum through red.global.add (L2 atomics)
  float col_scale = 1.0f;        // columns [scale_c0, scale_c1) are multiplied by col_scale after the bias
  int scale_c0 = 0, scale_c1 = 0;  //   (multiples of 32; used to hand K to the attention kernel in log2 units)
  op16* out_op16 = nullptr;      // optional op16 destination [*, ld_out_op16]
  int ld_out = 0;
  int ld_out_op16 = 0;
  // CTA-pair kernel only: process the row tiles last-to-first. The consumer of a large activation then starts with the
  // rows its producer wrote last -- the ones still in the 126 MB L2 -- instead of the ones evicted first (used for fc2;
  // measured effect at 8 images: within run-to-run noise, LayerNorm's matching row order -2 %).
  int reverse_m = 0;
  int narrow_tiles = 0;          // single-CTA kernel: 64-column tiles whatever N is (few-row GEMMs: more CTAs instead of wider ones)
  // LayerNorm folded into the GEMMs on either side of it (CTA-pair kernel; encoder.cu explains the algebra):
  //  producer (a residual add, accumulate != 0): besides x += ..., write op16(x16_gamma[col] * x_new) to row
  //    x16_rowmap[row] (null: row) of x16_out and the row's partial (mean, sum of squared deviations) to stats_out[row][slot]
  //  consumer (op16 output): A is such an x16; y = rstd * acc - rstd * mean * ln_cs[n] + ln_bw[n] (ln_bw = bias + wb) with mean /
  //    rstd from ln_stats[ln_rowmap[row] or row][0..ln_np); a negative map entry is a zero (pad) row: its output is ln_bw - ln_wb
  op16* x16_out = nullptr;
  int ld_x16 = 0;
  const float* x16_gamma = nullptr;
  const int* x16_rowmap = nullptr;
  float2* stats_out = nullptr;
  const float2* ln_stats = nullptr;
  int ln_np = 0;
  const int* ln_rowmap = nullptr;
  int ln_dim = 0;
  float ln_eps = 1e-6f;
  const float* ln_cs = nullptr;
  const float* ln_bw = nullptr;   // bias + wb (the epilogue then ignores `bias`)
  const float* ln_wb = nullptr;
};

// number of statistics slots per row the producer writes for an output width N (two column halves per N tile)
int gemm_ln_stat_slots(int M, int N);

// C[M,N] = A[M,K] * W[N,K]^T, op16 operands, fp32 accumulation in TMEM. A: row pitch lda, W: row pitch ldw.
void gemm_op16(const op16* A, int lda, const op16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
               cudaStream_t stream);

// 3x3 / pad-1 convolution over the 64 x 64 token grid as an implicit GEMM (no im2col buffer): in = op16 [n_images*4096, channels]
// token-major, W = [N, 9*channels] tap-major ((ky*3+kx)*channels + c), out = fp32 [n_images*4096, N].
void gemm_conv3x3_grid(const op16* in, int n_images, int channels, const op16* W, float* out_f32, int N, cudaStream_t stream);

// SM count of the CURRENT device (cached per device; thread-safe).
int sm_count();
// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device. The attribute is per device and per
// function, so the bookkeeping is too (a process may hold contexts on several GPUs); thread-safe.
void ensure_dyn_smem(const void* func, int bytes);

// Optional per-launch CUDA-event timing by kernel class (bench.py's roofline / breakdown pass).
// Inactive (null) in the product path: zero overhead.
enum KernelClass {
  KC_PREPROCESS = 0, KC_LAYERNORM, KC_GEMM_PATCH, KC_GEMM_QKV, KC_ATTN_WINDOW, KC_ATTN_GLOBAL, KC_GEMM_PROJ,
  KC_GEMM_FC1, KC_GEMM_FC2, KC_NECK, KC_DEC_TOKEN, KC_DEC_GEMM, KC_DEC_ATTN, KC_DEC_UPSCALE, KC_POST_UPSAMPLE,
  KC_POST_HULL, KC_COUNT
};
const char* kernel_class_name(int kc);

struct Profiler {
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> pool;
  struct Rec { int kc; int e0, e1; double flops; double bytes; };
  std::vector<Rec> recs;
  int next = 0;
  bool active = false;
  int begin(int kc, double flops = 0.0, double bytes = 0.0);     // algorithmic FLOPs / HBM bytes of the record
  void end(int rec);
  void collect(double* ms, long long* launches, double* flops, double* bytes);   // arrays of KC_COUNT; call after a stream sync
  void reset();
  ~Profiler();
};
// scope helper: times everything launched between construction and destruction as one record
struct ProfScope {
  Profiler* p; int r;
  ProfScope(Profiler* prof, int kc, double flops = 0.0, double bytes = 0.0) : p(prof && prof->active ? prof : nullptr), r(-1) {
    if (p) r = p->begin(kc, flops, bytes);
  }
  ~ProfScope() { if (p) p->end(r); }
};

}  // namespace ysi
