/* libysi.so -- C ABI of the B200-native SAM stage (image + box prompts -> masks, crops, per-mask metrics).
 *
 * The reference (gavinlouuu-kpt/yolo-sam-inference) is pure Python and has no native layer; this is the
 * boundary a maintainer binds with ctypes.  Every entry point below names the
 * reference code it replaces (paths relative to /root/reference/src/yolo_sam_inference/).
 *
 * Conventions: plain pointers and sizes only; all buffers are caller-allocated HOST memory unless the
 * name says "dev"; return 0 on success, negative on error (text via ysi_last_error). A context is bound
 * to one CUDA device and is not thread-safe; distinct contexts are independent (no global state, no NCCL).
 */
#ifndef YSI_H_
#define YSI_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define YSI_API __attribute__((visibility("default")))
#else
#define YSI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ysi_ctx ysi_ctx;

#define YSI_MAX_GLOBAL_LAYERS 8
#define YSI_PERIM_BINS 10 /* skimage perimeter codes 5,7,13,15,17,21,23,25,27,33 (in this order) */

/* per-mask flags */
#define YSI_FLAG_EMPTY_MASK 1u      /* area == 0: the reference raises IndexError at utils/metrics.py:28 */
#define YSI_FLAG_HULL_DEGENERATE 2u /* no contour / Qhull failure branch, utils/metrics.py:52-59 */
#define YSI_FLAG_CONTOUR_TRUNCATED 4u /* contour longer than the trace budget (never for H,W <= 4096) */

typedef struct {
  int32_t hidden_size;  /* ViT width D: 768 (B) / 1024 (L) / 1280 (H); head_dim = D / num_heads must be 64 or 80 */
  int32_t num_layers;
  int32_t num_heads;
  int32_t mlp_dim;
  int32_t num_global;   /* entries used in global_attn_indexes */
  int32_t global_attn_indexes[YSI_MAX_GLOBAL_LAYERS];
  int32_t max_batch;    /* images per ysi_run_batch call (activation workspace is sized for this) */
  int32_t max_boxes;    /* box prompts decoded per launch; a batch with more is processed in chunks of max_boxes that
                           share the batch's image embeddings (the reference loops over any number, pipeline.py:170) */
  int32_t max_image_h;  /* INITIAL capacity for original images; larger images (up to 4096 x 4096) re-allocate the */
  int32_t max_image_w;  /* image-sized buffers on first use (the reference has no size limit, pipeline.py:206-210) */
} ysi_config;

/* One tensor of SamModel.state_dict() (the object built at pipeline.py:76), fp32, C-contiguous. */
typedef struct {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} ysi_tensor_desc;

/* Raw per-mask morphometrics: every field the 16-key dict of utils/metrics.py:102-119 is derived from.
 * Integers are bit-exact w.r.t. the reference given the same mask; the host shim forms the float64
 * scalars with the reference's own formulas (utils/metrics.py:62-100). Rows are "x", columns are "y"
 * in the reference's dict (utils/metrics.py:97). */
typedef struct {
  int64_t area;                         /* props.area, :62 */
  int64_t sum_r, sum_c;                 /* centroid = sum / area, :87 */
  int32_t min_r, min_c, max_r, max_c;   /* props.bbox (max exclusive), :97 */
  uint32_t perim_hist[YSI_PERIM_BINS];  /* border-pixel code counts of props.perimeter, :65 */
  int64_t hull_area;                    /* convex_props.area, :68 */
  uint32_t hull_perim_hist[YSI_PERIM_BINS]; /* convex_props.perimeter, :69 */
  int64_t disk_n;                       /* pixels of the centre disk, :84-92 */
  uint64_t disk_sum;                    /* sum of (R+G+B) over the disk */
  uint64_t disk_sumsq;                  /* sum of (R+G+B)^2 over the disk */
  uint32_t flags;
  int32_t contour_points;               /* vertices of contours[0] (closing point not repeated), :34 */
  int32_t hull_vertices;
  int32_t reserved;
  uint32_t mask_hist[256];              /* histogram of floor((R+G+B)/3) over the mask (README.md:8) */
} ysi_mask_metrics;

typedef struct {
  float h2d_ms, preprocess_ms, encoder_ms, decoder_ms, postprocess_ms, metrics_ms, d2h_ms, total_ms;
} ysi_timing;

/* pixel formats of host images handed to ysi_submit */
#define YSI_PIX_RGB8 0   /* uint8 [H, W, 3] RGB: what _load_image returns (pipeline.py:206-210) */
#define YSI_PIX_GRAY8 1  /* uint8 [H, W]: the raw samples of an 8-bit single-channel file; the device replicates them to RGB */
#define YSI_PIX_GRAY16 2 /* uint16 [H, W] little endian: raw samples of a 16-bit file; the device takes v >> 8 (what
                            cv2.imread's default flags produce) and replicates to RGB */

/* One batch of same-sized images with their box prompts (ysi_submit). Host pointers; pinned memory makes the copies
 * asynchronous. Buffers must stay valid until the matching ysi_wait_batch. */
typedef struct {
  int32_t n_images;
  int32_t height, width;
  int32_t row_stride;            /* bytes between image rows */
  int32_t pixel_format;          /* YSI_PIX_* */
  const void* const* images;     /* n_images pointers */
  const float* boxes_xyxy;       /* float32 [sum(box_counts), 4], original-image pixels (pipeline.py:84-87) */
  const int32_t* box_counts;     /* boxes per image */
  uint8_t* masks_out;            /* uint8 [nb, H, W] of 0/1, or NULL */
  uint8_t* packed_out;           /* uint8 [nb, ceil(H*W/8)] np.packbits rows (utils/mask_encoding.py:24), or NULL */
  ysi_mask_metrics* metrics_out; /* [nb], or NULL */
} ysi_batch;

/* ---- lifecycle -------------------------------------------------------------------------------- */
/* replaces SamModel.from_pretrained(...).to(device), pipeline.py:69-77 */
YSI_API int ysi_create(int device, const ysi_config* cfg, ysi_ctx** out);
YSI_API int ysi_load_weights(ysi_ctx* ctx, const ysi_tensor_desc* tensors, size_t n);
YSI_API void ysi_destroy(ysi_ctx* ctx);
YSI_API const char* ysi_last_error(const ysi_ctx* ctx); /* ctx may be NULL: error of the failed ysi_create */

/* ---- the hot path ----------------------------------------------------------------------------- */
/* One image: replaces the body of `if len(boxes) > 0:` in process_single_image, pipeline.py:161-175
 * (sam_processor + per-box _process_sam_mask :89-124 + calculate_metrics utils/metrics.py:9-119).
 *   rgb        uint8 [H, W, 3], row pitch row_stride bytes (what _load_image returns, :206-210)
 *   boxes_xyxy float32 [nb, 4] in original-image pixels (what _detect_cells returns, :84-87)
 *   masks_out  uint8 [nb, H, W] of 0/1, or NULL
 *   packed_out uint8 [nb, ceil(H*W/8)] np.packbits order (utils/mask_encoding.py:24), or NULL
 *   metrics_out[nb]
 * nb == 0 returns immediately (pipeline.py:176-179). */
YSI_API int ysi_run(ysi_ctx* ctx, const uint8_t* rgb, int H, int W, int row_stride, const float* boxes_xyxy, int nb,
            uint8_t* masks_out, uint8_t* packed_out, ysi_mask_metrics* metrics_out, ysi_timing* timing);

/* Several same-sized images per launch (the unit of work the folder partition of pipeline.py:537-577
 * hands to one GPU). box_counts[i] boxes belong to image i; boxes / masks / metrics are concatenated. */
YSI_API int ysi_run_batch(ysi_ctx* ctx, int n_images, const uint8_t* const* rgb, int H, int W, int row_stride,
                  const float* boxes_xyxy, const int32_t* box_counts, uint8_t* masks_out, uint8_t* packed_out,
                  ysi_mask_metrics* metrics_out, ysi_timing* timing);

/* Pipelined form of ysi_run_batch: a context has two slots; ysi_submit_batch enqueues a batch into `slot` (0/1) and
 * returns at once, ysi_wait_batch blocks until that slot's results are in the host buffers. Submitting batch i+1 before
 * waiting for batch i overlaps its host->device copy and encoder with the decoder / metrics / device->host copy of
 * batch i (four CUDA streams inside the context). The host buffers (pinned for true overlap) must stay valid until the
 * matching wait; a slot must be waited for before it is reused. */
YSI_API int ysi_submit_batch(ysi_ctx* ctx, int slot, int n_images, const uint8_t* const* rgb, int H, int W, int row_stride,
                     const float* boxes_xyxy, const int32_t* box_counts, uint8_t* masks_out, uint8_t* packed_out,
                     ysi_mask_metrics* metrics_out);
YSI_API int ysi_wait_batch(ysi_ctx* ctx, int slot, ysi_timing* timing);
/* ysi_submit_batch for any pixel format (GPU-side ingest). With YSI_PIX_GRAY8 / GRAY16 the host
 * hands over the raw samples of a baseline (uncompressed) TIFF strip -- 1 or 2 bytes per pixel over PCIe instead of 3 --
 * and the 16 -> 8 bit reduction and grey -> RGB replication of _load_image (pipeline.py:206-210) run on the device. */
YSI_API int ysi_submit(ysi_ctx* ctx, int slot, const ysi_batch* batch);

/* Page-locked host memory for image staging / result buffers (cudaHostAlloc): copies from and to it are truly
 * asynchronous, so a decode thread pool can fill the next batches while the GPU works (the host side of the folder
 * entry point, pipeline.py:212-263). Independent of any context. */
YSI_API int ysi_alloc_pinned(int device, size_t bytes, void** out);   /* portable: usable by every context of the process */
YSI_API void ysi_free_pinned(void* p);

/* ---- measurement support (bench.py) ------------------------------------------------------------- */
/* resident input pool: images uploaded once, then processed straight from HBM (device-resident leg) */
YSI_API int ysi_pool_upload(ysi_ctx* ctx, int pool_size, int idx, const uint8_t* rgb, int H, int W, int row_stride);
/* run images [first_idx, first_idx+n) of the pool through the same two-slot pipeline; sync == 0 only enqueues
 * (results stay on the device) */
YSI_API int ysi_compute_pool(ysi_ctx* ctx, int first_idx, int n, const float* boxes_xyxy, const int32_t* box_counts, int sync,
                     ysi_timing* timing);
/* CUDA events on the context's own stream (torch.cuda.Event would only see torch's stream) */
YSI_API int ysi_timer_record(ysi_ctx* ctx, int slot);
YSI_API int ysi_timer_elapsed_ms(ysi_ctx* ctx, int slot_a, int slot_b, float* ms);
YSI_API int ysi_sync(ysi_ctx* ctx);
/* per-kernel-class event timing: enable, run steps, read (returns the number of classes written) */
YSI_API int ysi_profile(ysi_ctx* ctx, int enable);
YSI_API int ysi_profile_read(ysi_ctx* ctx, int max_classes, const char** names, double* ms, int64_t* records, double* flops,
                     double* bytes /* algorithmic HBM bytes of the HBM-bound classes, may be NULL */);

/* ---- stage-level entry points (parity tests; each is one row of SURVEY.md section 8a) --------------- */
/* a1: sam_processor(image) -> pixel_values fp32 [n,3,1024,1024] (image_processing_sam.py:205-250) */
YSI_API int ysi_preprocess(ysi_ctx* ctx, int n_images, const uint8_t* const* rgb, int H, int W, int row_stride,
                   float* pixel_values_out);
/* a3: SamVisionEncoder.forward (modeling_sam.py:1058-1072). pixel_values fp32 [n,3,1024,1024] ->
 * image_embeddings fp32 [n,256,64,64]; hidden_out (optional) fp32 [num_layers+1, n, 64,64,D]:
 * slot 0 = patch embed + pos, slot i+1 = after layer i. */
YSI_API int ysi_encode(ysi_ctx* ctx, int n_images, const float* pixel_values, float* image_embeddings_out, float* hidden_out);
/* a4+a5: prompt encoder + mask decoder (modeling_sam.py:658-698, 461-543) for one image.
 * boxes_1024 float64 [nb,4] already in the 1024-frame (processing_sam.py:215-234) -> low-res logits
 * fp32 [nb,256,256]; sparse_out (optional) fp32 [nb,2,256]. */
YSI_API int ysi_decode(ysi_ctx* ctx, const float* image_embeddings, const double* boxes_1024, int nb, float* low_res_out,
               float* sparse_out);
/* a6: post_process_masks (image_processing_sam.py:379-430) + `> 0` : low-res logits fp32 [nb,256,256]
 * -> masks uint8 [nb,H,W]; upsampled_out (optional) fp32 [nb,H,W]. */
YSI_API int ysi_postprocess(ysi_ctx* ctx, const float* low_res, int nb, int H, int W, uint8_t* masks_out,
                    float* upsampled_out);
/* a7: calculate_metrics (utils/metrics.py:9-119) on given masks uint8 [nb,H,W] of one image. */
YSI_API int ysi_metrics(ysi_ctx* ctx, const uint8_t* rgb, int H, int W, int row_stride, const uint8_t* masks, int nb,
                ysi_mask_metrics* metrics_out);
/* the GEMM core on its own: C[M,N] fp32 = A[M,K] * W[N,K]^T (+bias[N]) with bf16-rounded operands;
 * act: 0 none, 1 erf-GELU, 2 ReLU. */
YSI_API int ysi_gemm(ysi_ctx* ctx, const float* A, const float* W, const float* bias, int M, int N, int K, int act,
             float* C_out);
/* windowed / global attention of one encoder layer on its own (modeling_sam.py:843-882):
 * qkv fp32 [n_seq, T, 3*heads*head_dim] (T = 196 windowed, 4096 global), rel_pos_h/w fp32 [2S-1, head_dim]
 * -> out fp32 [n_seq, T, heads*head_dim]. head_dim 64 (ViT-B/L) or 80 (ViT-H). */
YSI_API int ysi_attention(ysi_ctx* ctx, const float* qkv, const float* rel_pos_h, const float* rel_pos_w, int n_seq,
                  int heads, int head_dim, int is_global, float* out);

/* the GEMM core through the production tile dispatcher, with the encoder's epilogue kinds:
 * out_kind 0: C = result (fp32); 1: C = bf16-rounded result; 2: C += result (the residual add of
 * modeling_sam.py:969-971). C_inout fp32 [M,N] is read for kind 2. */
YSI_API int ysi_gemm_ex(ysi_ctx* ctx, const float* A, const float* W, const float* bias, int M, int N, int K, int act,
                int out_kind, float* C_inout);
/* measurement support: time `iters` launches of one GEMM shape on device-resident bf16 operands.
 * mode 0 bf16-output epilogue, 1 fp32 red-add epilogue, 2 drain-only (mainloop speed); pair != 0: CTA-pair kernel. */
YSI_API int ysi_gemm_bench(ysi_ctx* ctx, int M, int N, int K, int pair, int mode, int iters, float* ms_per_iter);

/* measurement support: ms per launch of one attention shape on device-resident random operands */
YSI_API int ysi_attention_bench(ysi_ctx* ctx, int n_seq, int heads, int head_dim, int is_global, int iters, float* ms_per_iter);

/* image-wide positional embedding fp32 [256,64,64] (modeling_sam.py:1128-1139), computed at weight load */
YSI_API int ysi_get_image_pe(ysi_ctx* ctx, float* out);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
YSI_API int64_t ysi_launch_count(const ysi_ctx* ctx);
/* last completed timing of the run entry points */
YSI_API int ysi_version(void);
/* 16-bit encoding of the tensor-core operands this library was compiled for: "bf16" (libysi.so) or "fp16"
 * (libysi_fp16.so); accumulation, residual stream, LayerNorm and softmax statistics are fp32 in both. */
YSI_API const char* ysi_operand_dtype(void);

#ifdef __cplusplus
}
#endif
#endif /* YSI_H_ */
