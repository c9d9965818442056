/*
 * agr_lpips.h — C ABI of the pieces of the perceptual loss (LPIPS on a VGG-16 trunk) that are not convolutions
 * (second part; reference: network/lpips/lpips.py:23-124, network/lpips/__init__.py:40-42,
 * network/lpips/pretrained_networks.py:96-134, called from main_avatar.py:117-124 on the cropped patches).
 *
 * The thirteen 3x3 convolutions + bias + ReLU of the trunk run through agr_conv.h (epilogue activate = 3); what remains:
 *   - the four 2x2 / stride-2 max-poolings between the trunk's slices (torchvision vgg16.features[4,9,16,23]);
 *   - per tapped layer, the LPIPS head: unit-normalise both feature stacks over channels, squared difference, the learned
 *     non-negative 1x1 "lin" weights, spatial mean — one pass forward, one pass backward, nothing materialised.
 * Layout: NHWC; dtype AGR_F32 (0) / AGR_BF16 (1) as in agr_conv.h.
 */
#ifndef AGR_LPIPS_H_
#define AGR_LPIPS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* y[n][oy][ox][c] = max over the 2x2 window of x (floor mode: OH = H/2, OW = W/2), nn.MaxPool2d(2, 2). */
int agr_maxpool2x2_forward(int32_t dtype, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* cuda_stream);
/* dx = dy routed to the FIRST maximum of each window in row-major window order (torch's tie rule), zero elsewhere. */
int agr_maxpool2x2_backward(int32_t dtype, const void* x, const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                            void* cuda_stream);

/* One tapped layer of LPIPS for ONE image pair: f (2, pixels, C) holds the features of image 0 and image 1.
 *   n_i = sqrt(sum_c f_i[c]^2 + 1e-10),  u_i = f_i / (n_i + 1e-10)                 (lpips/__init__.py:40-42)
 *   out[0] += scale * sum_pixels sum_c w[c] * (u_0[c] - u_1[c])^2                  (lpips.py:93-103; scale = 1 / pixels)
 * `out` (1 float on the device) is ACCUMULATED: the five layers of the metric add into the same scalar. */
int agr_lpips_layer_forward(int32_t dtype, const void* f, const float* w, int64_t pixels, int32_t C, float scale, float* out,
                            void* cuda_stream);
/* df (2, pixels, C) = d(out)/d(f) * g[0] for the same inputs; `g` (1 float on the device) is the upstream gradient. */
int agr_lpips_layer_backward(int32_t dtype, const void* f, const float* w, int64_t pixels, int32_t C, float scale, const float* g,
                             void* df, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
