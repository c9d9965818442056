/*
 * agr_loss.h — C ABI of the photometric loss head that follows the rasterizer (first part).
 *
 * Replaces the trainer's elementwise chain between `avatar_net.render` and `total_loss.backward()`
 * (main_avatar.py:193-222): background fill of the ground truth outside its mask, boundary-mask compositing of
 * rendering and ground truth, the L1 image term and the L1 mask term — about 15 ATen launches and 8 full-size
 * temporaries per view in the reference — by ONE pass that also writes the gradients w.r.t. the rendered colour
 * and opacity maps (the loss is piecewise linear, so its gradient is known as soon as the residual's sign is).
 *
 * Per pixel p of view v (bm = 1 - boundary[p], bg the background colour):
 *   img[c] = rgb[p][c] * bm + (1 - bm) * bg[c]                                   (:200)
 *   gt [c] = (mask[p] ? gt_rgb[p][c] : bg[c]) * bm + (1 - bm) * bg[c]            (:195-196,201)
 *   l1_sum   += sum_c |img[c] - gt[c]|                                           (:206-207, mean over 3*H*W)
 *   mask_sum += |alpha[p] * bm - mask[p] * bm|                                   (:213-219, mean over H*W)
 *   d_rgb[p][c] = w_l1   / (3 * pixels) * sign(img[c] - gt[c]) * bm              (sign(0) = 0, as torch.abs')
 *   d_alpha[p]  = w_mask / pixels       * sign(alpha[p] * bm - mask[p] * bm) * bm
 * with pixels = V*H*W: the mean over a view batch of equally sized views.
 *
 * rgb (pixels,3), alpha (pixels), gt_rgb (pixels,3) fp32; mask, boundary (pixels) uint8 (non-zero = true); bg (3) fp32
 * on the device; sums (2) fp32 on the device, ACCUMULATED (caller zeroes): {l1_sum, mask_sum};
 * d_rgb (pixels,3) / d_alpha (pixels) fp32, fully overwritten, either may be NULL.  alpha / d_alpha may be NULL
 * (no mask term).
 */
#ifndef AGR_LOSS_H_
#define AGR_LOSS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int agr_photometric_loss(const float* rgb, const float* alpha, const float* gt_rgb, const uint8_t* mask,
                         const uint8_t* boundary, const float* bg, int64_t pixels, float w_l1, float w_mask,
                         float* sums, float* d_rgb, float* d_alpha, void* cuda_stream);
/* Same, with the ground-truth colour as the camera stores it: gt_rgb (pixels,3) uint8, value / 255 (the dataset divides by
 * 255 on the host, dataset/dataset_mv_rgb.py; uploading bytes quarters the H2D traffic of a 16-view step). */
int agr_photometric_loss_u8(const float* rgb, const float* alpha, const uint8_t* gt_rgb, const uint8_t* mask,
                            const uint8_t* boundary, const float* bg, int64_t pixels, float w_l1, float w_mask,
                            float* sums, float* d_rgb, float* d_alpha, void* cuda_stream);
#ifdef __cplusplus
}
#endif
#endif
