/*
 * agr_avatar.h — C ABI of the map -> per-Gaussian gather of AvatarNet.
 *
 * Replaces, in AvatarNet.get_positions / get_others / get_colors (network/avatar.py:93-124):
 *     front, back = split(map, C, 1); map = cat([front, back], 3)[0].permute(1, 2, 0); vals = map[cano_smpl_mask]
 * i.e. a channel split, a width concat, a permute and a boolean-mask gather (nonzero + index) per call — four
 * full-resolution passes over (2C,1024,1024) maps — by one gather straight out of the NHWC decoder outputs, using the
 * constant pixel list of the canonical mask.
 *   front/back : (V,S,S,C) NHWC, dtype AGR_F32 / AGR_BF16 (agr_styleunet.h)
 *   half (N) int32: 0 -> front, 1 -> back;  pix (N) int32: row*S + col inside that half
 *   out  : (V,N,C) fp32
 * Backward: d_front/d_back (V,S,S,C) dtype are ZERO-FILLED by the call and receive d_out at the N pixels
 * (each pixel belongs to at most one Gaussian, so no atomics).
 */
#ifndef AGR_AVATAR_H_
#define AGR_AVATAR_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int agr_gather_maps_forward(int32_t dtype, const void* front, const void* back, const int32_t* half, const int32_t* pix,
                            float* out, int32_t V, int32_t S, int32_t C, int32_t N, void* cuda_stream);
int agr_gather_maps_backward(int32_t dtype, const float* d_out, const int32_t* half, const int32_t* pix, void* d_front,
                             void* d_back, int32_t V, int32_t S, int32_t C, int32_t N, void* cuda_stream);
#ifdef __cplusplus
}
#endif
#endif
