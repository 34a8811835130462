// Internal launch API between the translation units of libysi.so (device pointers everywhere).
#pragma once
#include "../../include/ysi.h"
#include "common.h"

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
void launch_upsample_stats(const float* low, int nmask, PostGeom g, const uint16_t* sum3, const int* mask_image,
                           uint8_t* masks, float* up_logits, MaskStatsDev* stats, cudaStream_t s);
// same statistics from given mask bytes (a7 alone)
void launch_mask_stats(const uint8_t* masks_in, int nmask, int H, int W, const uint16_t* sum3, const int* mask_image,
                       MaskStatsDev* stats, cudaStream_t s);
// second half of a7: contour 0 -> convex hull -> hull raster stats; centre-disk brightness sums; final rows.
void launch_contour_hull_disk(const uint8_t* masks, int nmask, int H, int W, const uint16_t* sum3,
                              const int* mask_image, const MaskStatsDev* stats, ysi_mask_metrics* out,
                              cudaStream_t s);
// np.packbits(mask.reshape(-1)) per mask: [nmask, ceil(H*W/8)]
void launch_packbits(const uint8_t* masks, uint8_t* packed, int nmask, long long npix, cudaStream_t s);
// uint8 RGB [n,H,W,3] (pitch row_stride) -> R+G+B uint16 planes
void launch_sum3(const uint8_t* rgb, int n, int H, int W, int row_stride, uint16_t* sum3, cudaStream_t s);

}  // namespace ysi
