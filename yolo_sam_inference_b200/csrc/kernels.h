// Internal launch API between the translation units of libysi.so (device pointers everywhere).
#pragma once
#include "../../include/ysi.h"
#include "common.h"

#include <cmath>

namespace ysi {

// ------------------------------------------------------------------ post-processing + morphometrics
struct PostGeom {
  int H, W;     // original image size
  int rh, rw;   // reshaped_input_size (resize-longest-edge-to-1024 result)
};
PostGeom make_post_geom(int H, int W);

// per-mask accumulators filled by the fused upsample/threshold/stats kernel
struct MaskStatsDev {
  unsigned long long area, sum_r, sum_c;
  int min_r, min_c, max_r, max_c;      // inclusive; initialised to +inf/-inf
  unsigned int first_cell;             // min over mixed 2x2 cells of r0 * W + c0 (0xFFFFFFFF: none)
  unsigned int perim_hist[YSI_PERIM_BINS];
  unsigned int mask_hist[256];
};

void launch_init_stats(MaskStatsDev* stats, int nmask, cudaStream_t s);
// a6 (+ first half of a7): low-res logits [nmask,256,256] -> mask bytes [nmask,H,W] + stats.
// sum3: per-image R+G+B planes uint16 [n_img,H,W] (may be null: no intensity histogram),
// mask_image[m]: image index of mask m (null => all image 0). up_logits optional fp32 [nmask,H,W].
// gray: floor((R+G+B)/3) uint8 planes [n_img,H,W] (histogram bins; required with sum3 for the 1024x1024 fast path).
// packed [nmask, ceil(H*W/8)] (np.packbits rows, utils/mask_encoding.py:24) is always written: it is what the contour
// kernel reads and the default wire format; `masks` [nmask,H,W] holds the byte masks when want_bytes (and is scratch for
// the generic geometry, which derives the packed rows from it).
void launch_upsample_stats(const float* low, int nmask, PostGeom g, const uint16_t* sum3, const uint8_t* gray,
                           const int* mask_image, uint8_t* masks, uint8_t* packed, bool want_bytes, float* up_logits,
                           MaskStatsDev* stats, cudaStream_t s);
// same statistics from given mask bytes (a7 alone)
void launch_mask_stats(const uint8_t* masks_in, int nmask, int H, int W, const uint16_t* sum3, const int* mask_image,
                       MaskStatsDev* stats, cudaStream_t s);
// second half of a7: contour 0 -> convex hull -> hull raster stats; centre-disk brightness sums; final rows.
// (reads the PACKED mask rows)
void launch_contour_hull_disk(const uint8_t* packed, int nmask, int H, int W, const uint16_t* sum3,
                              const int* mask_image, const MaskStatsDev* stats, ysi_mask_metrics* out,
                              cudaStream_t s);
// np.packbits(mask.reshape(-1)) per mask: [nmask, ceil(H*W/8)]
void launch_packbits(const uint8_t* masks, uint8_t* packed, int nmask, long long npix, cudaStream_t s);
// uint8 RGB [n,H,W,3] (pitch row_stride) -> R+G+B uint16 planes
void launch_sum3(const uint8_t* rgb, int n, int H, int W, int row_stride, uint16_t* sum3, uint8_t* gray, cudaStream_t s);
// f3 of SURVEY section 8f (GPU-side ingest): raw single-channel pixels as they sit in a baseline TIFF strip -- uint8, or
// uint16 reduced to 8 bits exactly like cv2.imread's default flags do (v >> 8, pipeline.py:206-210) -- are expanded on
// the device to the uint8 RGB image the rest of the path reads (grey replicated to three channels) together with the
// R+G+B and floor((R+G+B)/3) planes. src: dense [n,H,W] of bytes_per_px (1 or 2, little endian).
void launch_gray_ingest(const void* src, int bytes_per_px, int n, int H, int W, uint8_t* rgb, uint16_t* sum3, uint8_t* gray,
                        cudaStream_t s);

// ------------------------------------------------------------------ encoder
// fused flash-style attention with decomposed rel-pos bias (attn.cu)
// qkv K columns pre-scaled by hd^-0.5 * log2(e), rel_tab pre-scaled by log2(e) (see attn.cu);
// unwindow (windowed only): rows are written in token order and the 64->70 pad tokens are dropped
constexpr float ATTN_LOG2E = 1.4426950408889634f;
inline float attn_k_scale(int head_dim) { return static_cast<float>(1.0 / std::sqrt(static_cast<double>(head_dim))) * ATTN_LOG2E; }
inline int attn_table_cols(int head_dim) { return head_dim == 64 ? 64 : 128; }    // rel-pos table row pitch (zero padded)
void launch_encoder_attention(const op16* qkv, const op16* rel_tab, op16* out, int n_seq, int T, int heads, int head_dim,
                              bool is_global, bool unwindow, cudaStream_t stream);

constexpr int NECK_C2 = 512;        // neck 3x3 conv input row: 256 channels as [hi | lo]
constexpr int NECK_K2 = 9 * NECK_C2;  // its im2col contraction length
constexpr int PATCH_K = 2 * 768;   // patch-embed contraction length: 3*16*16 pixel hi terms + as many lo terms

struct EncoderLayerW {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const op16 *w_qkv, *w_proj, *w_fc1, *w_fc2;
  const float *b_qkv, *b_proj, *b_fc1, *b_fc2;
  const op16* rel_tab;   // [256,HDP] * log2(e): rows 0..127 rel_pos_h (zero padded), rows 128..255 rel_pos_w
  int is_global;
  // LayerNorm folded into qkv / fc1 (encoder.cu): cs[n] = sum_k gamma[k] W[n,k], wb[n] = sum_k beta[k] W[n,k], bw = bias + wb
  const float *cs_qkv, *wb_qkv, *bw_qkv, *cs_fc1, *wb_fc1, *bw_fc1;
};

struct EncoderW {
  int D, L, heads, mlp;
  int head_dim = 64;
  int residual_mode = 2;   // GemmEpilogue::accumulate for the two residual adds of a layer
  const op16* w_patch;    // [D, PATCH_K] = [W | W] (hi / lo split of the pixels, encoder.cu)
  const float* b_patch;   // [D]
  const float* pos_embed; // [4096, D]
  const EncoderLayerW* layers;   // host array, L entries
  const op16* w_neck1;    // [256, 2D] = [W | W]
  const float *neck_ln1_g, *neck_ln1_b;
  const op16* w_neck2;    // [256, NECK_K2]  (tap-major: [(ky*3+kx)*512 + cin] and the same weight again at +256)
  const float *neck_ln2_g, *neck_ln2_b;
};

struct EncoderWork {      // activation workspace for `cap` images
  int cap;
  op16* a_patch;          // [cap*4096, PATCH_K]: pixel hi terms | lo terms
  float* x;               // [cap*4096, D]   fp32 residual stream
  op16* h;                // [cap*4900, D]
  op16* qkv;              // [cap*4900, 3D]
  op16* attn;             // [cap*4096, D]  attention output in token order
  op16* u;                // [cap*4096, mlp]
  float* n1;              // [cap*4096, 256]
  op16* n1b;              // [cap*4096, NECK_C2]  hi | lo
  op16* a_neck;           // [cap*4096, NECK_K2]
  float* n2;              // [cap*4096, 256]
  const int* win_row_map; // [cap*4900] window row -> token row (or -1)
  // folded LayerNorm: windowed operand copy (pad rows stay zero for ever), token -> window row map, per-row statistics slots
  op16* h_win;            // [cap*4900, D]
  const int* tok_win_map; // [cap*4096] token row -> window row
  float2* ln_stats;       // [cap*4096, LN_STAT_SLOTS]
};
constexpr int LN_STAT_SLOTS = 16;   // >= 2 * (D / 192) for every supported width (ViT-H: 1280 / 256 * 2 = 10, ViT-L: 1024 / 256 * 2 = 8)

// torchvision / Pillow fixed-point antialias resampling tables for one axis (see encoder.cu)
struct ResizeTables {
  int in_size = 0, out_size = 0, ksize = 0, prec = 0;
  std::vector<int> xmin, xsize;
  std::vector<int16_t> weights;      // [out_size, ksize]
};
struct ResizeTablesDev {
  int in_size = 0, out_size = 0, ksize = 0, prec = 0;
  const int* xmin = nullptr; const int* xsize = nullptr; const int16_t* weights = nullptr;
};
ResizeTables build_resize_tables(int in_size, int out_size);
// horizontal pass: src uint8 [n,H,*,3] (row pitch row_stride, image pitch img_stride bytes) -> dst dense [n,H,out,3]
void launch_resize_h(const uint8_t* src, int n, int H, int row_stride, size_t img_stride, const ResizeTablesDev& t, uint8_t* dst,
                     cudaStream_t s);
// vertical pass: src uint8 [n,*,W,3] -> dst dense [n,out,W,3]
void launch_resize_v(const uint8_t* src, int n, int row_stride, size_t img_stride, int W, const ResizeTablesDev& t, uint8_t* dst,
                     cudaStream_t s);
// resized uint8 RGB [n,src_h,src_w,3] (row pitch row_stride, image pitch img_stride bytes) -> normalised, zero-padded
// pixel_values fp32 [n,3,1024,1024] (optional) and/or the patch-embed A matrix op16 [n*4096, 768]
void launch_preprocess(const uint8_t* rgb, int n, int src_h, int src_w, int row_stride, size_t img_stride, const float* mean255,
                       const float* std255, float* pixel_values, op16* a_patch, cudaStream_t s);
void launch_im2col_patch_f32(const float* pixel_values, int n, op16* a_patch, cudaStream_t s);
void launch_build_win_row_map(int* map, int n_images, cudaStream_t s);
void launch_build_tok_win_map(int* map, int n_images, cudaStream_t s);
void launch_layernorm(const float* x, int rows_out, int D, const float* gamma, const float* beta, float eps,
                      op16* out_bf, float* out_f, bool windowed, cudaStream_t s, bool split = false, bool reverse = false);
// runs the whole encoder on work.a_patch (n images); result: image embeddings fp32 token-major [n*4096, 256].
// hidden_dump (optional, device) fp32 [(L+1), n*4096, D]
void encoder_forward(const EncoderW& w, const EncoderWork& work, int n, float* emb_out, float* hidden_dump,
                     cudaStream_t s, int64_t* launches, Profiler* prof = nullptr);

// ------------------------------------------------------------------ prompt encoder + mask decoder
struct DecAttnW {          // SamAttention weights, fp32 (token-side use)
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo;
  const op16 *wq3, *wk3, *wv3, *wo3;      // the four weights as three-term splits [W_hi | W_hi | W_lo] (tensor-core token path)
};
struct DecLayerW {
  DecAttnW self_attn, t2i, i2t;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b, *ln4_g, *ln4_b;
  const float *w_fc1, *b_fc1, *w_fc2, *b_fc2;       // 256 -> 2048 -> 256 (ReLU)
  // the same two weights as three-term op16 splits [W_hi | W_hi | W_lo] ([2048, 768] / [256, 6144]) for the tensor-core path of
  // the token MLP at many boxes: against activations [x_hi | x_lo | x_hi] the products x_hi W_hi + x_lo W_hi + x_hi W_lo keep
  // fp32-level accuracy (the dropped x_lo W_lo term is 2^-22 relative)
  const op16 *w_fc1_s3, *w_fc2_s3;
  const op16* w_kq_img;    // [256,256]: rows 0..127 t2i.k_proj, rows 128..255 i2t.q_proj  (input keys + pos)
  const float* b_kq_img;   // [256]
  const op16* w_v_img;     // [128,256] t2i.v_proj (input keys)
  const op16* w_i2t_out;   // [256,128] i2t.out_proj
};
struct DecoderW {
  const float* gauss;          // [2,128] shared_image_embedding.positional_embedding
  const float* point_embed;    // [4,256]
  const float* no_mask_embed;  // [256]
  const float* iou_token;      // [256]
  const float* mask_tokens;    // [4,256]
  DecLayerW layers[2];
  DecAttnW final_attn;
  const op16 *w_k_final, *w_v_final;     // [128,256]
  const float *lnf_g, *lnf_b;
  const op16* w_ct1;           // [256 = (dy,dx,o64), 512]  [W | W]
  const float *b_ct1, *lnu_g, *lnu_b;    // [64]
  const op16* w_ct2;           // [128 = (dy,dx,o32), 128]  [W | W]
  const float* b_ct2;          // [32]
  const float *hy_w0, *hy_b0, *hy_w1, *hy_b1, *hy_w2, *hy_b2;   // hypernetwork MLP of mask token 0
  const float* image_pe;       // [4096,256] token-major, built at weight load
};
struct DecoderWork {           // workspace for cap_img images and cap_box boxes
  int cap_img, cap_box;
  float* keys0;  op16* keys0_bf;  op16* keyspos0_bf;    // [cap_img*4096, 256]
  op16* kq16;    op16* v16;                              // [cap_box*4096, 256] / [cap_box*4096,128] per-box k|q and v projections
  float* kq0;    float* v0;                              // [cap_img*4096, 256] / [cap_img*4096,128]
  float* keys;   op16* keys_bf;   op16* keyspos_bf;      // [cap_box*4096, 256] per-box keys
  float* kq;                                             // [cap_box*4096, 256] pre-LayerNorm key update (fp32)
  op16* attn_i2t;                                        // [cap_box*4096, 128]
  op16* up1;                                             // [cap_box*16384, 128]  hi | lo
  float *tok0, *queries, *q_t2i, *attn_t2i, *k_tok, *v_tok, *hyper;   // token-side [cap_box, 7, *]
  float* tok_ws;                                         // token-side scratch: cap_box * (7*(6*256 + 2048) + 2*256 + 8*4*126) floats
  op16* tok_a3;                                          // [cap_box*7, 3*2048] three-term split of a token-side GEMM input
  double* boxes1024;  int* box_img;                      // [cap_box,4] / [cap_box]
};
void launch_image_pe(const float* gauss, float* image_pe, cudaStream_t s);
// emb: image embeddings fp32 token-major [n_img*4096,256]; boxes1024 (device, fp64 [nb,4]) and box_img
// (device int [nb]) live in work.  low_res_out: device fp32 [nb,256,256]; sparse_out optional [nb,2,256]
void decoder_forward(const DecoderW& w, const DecoderWork& work, const float* emb, int n_img, int nb,
                     float* low_res_out, float* sparse_out, cudaStream_t s, int64_t* launches, Profiler* prof = nullptr);

}  // namespace ysi
