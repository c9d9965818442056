// Photometric loss head (include/agr_loss.h): one streaming pass over the rendered maps of a view batch.
// HBM-bound: 12+4+12+1+1 B in, 12+4 B out per pixel; two block-reduced sums, one atomic pair per block.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_loss.h"
#include "../../include/agr_rasterizer.h"

namespace agr {

__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float gtf(const float* g, int64_t i) { return g[i]; }
__device__ __forceinline__ float gtf(const uint8_t* g, int64_t i) { return (float)g[i] * (1.f / 255.f); }   // camera images as stored

template <typename GT>
__global__ void __launch_bounds__(256) photometric_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ alpha,
                                                              const GT* __restrict__ gt_rgb, const uint8_t* __restrict__ mask,
                                                              const uint8_t* __restrict__ boundary, const float* __restrict__ bg,
                                                              int64_t pixels, float g_l1, float g_mask, float* __restrict__ sums,
                                                              float* __restrict__ d_rgb, float* __restrict__ d_alpha) {
    __shared__ float s_l1[8], s_mk[8];
    const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
    float l1 = 0.f, mk = 0.f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (int64_t)gridDim.x * blockDim.x) {
        const float bm = boundary[p] ? 0.f : 1.f;
        const bool m = mask[p] != 0;
        const float inv = 1.f - bm;
        const float r[3] = {rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]};
        const float t[3] = {m ? gtf(gt_rgb, 3 * p) : b0, m ? gtf(gt_rgb, 3 * p + 1) : b1, m ? gtf(gt_rgb, 3 * p + 2) : b2};
        const float bgc[3] = {b0, b1, b2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float img = r[c] * bm + inv * bgc[c];
            const float gt = t[c] * bm + inv * bgc[c];
            const float d = img - gt;
            l1 += fabsf(d);
            if (d_rgb) d_rgb[3 * p + c] = g_l1 * sgn(d) * bm;
        }
        if (alpha) {
            const float d = alpha[p] * bm - (m ? 1.f : 0.f) * bm;
            mk += fabsf(d);
            if (d_alpha) d_alpha[p] = g_mask * sgn(d) * bm;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        mk += __shfl_xor_sync(0xffffffffu, mk, o);
    }
    if ((threadIdx.x & 31) == 0) { s_l1[threadIdx.x >> 5] = l1; s_mk[threadIdx.x >> 5] = mk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += s_l1[i]; b += s_mk[i]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

}  // namespace agr

template <typename GT>
static int launch_photometric(const float* rgb, const float* alpha, const GT* gt_rgb, const uint8_t* mask, const uint8_t* boundary,
                              const float* bg, int64_t pixels, float w_l1, float w_mask, float* sums, float* d_rgb, float* d_alpha,
                              void* cuda_stream) {
    if (!rgb || !gt_rgb || !mask || !boundary || !bg || !sums || pixels < 0 || (d_alpha && !alpha)) return AGR_ERR_INVALID_ARGUMENT;
    if (pixels == 0) return AGR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    int64_t blocks = (pixels + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const float g_l1 = w_l1 / (3.f * (float)pixels), g_mask = w_mask / (float)pixels;
    agr::photometric_loss_kernel<GT><<<(unsigned)blocks, 256, 0, s>>>(rgb, alpha, gt_rgb, mask, boundary, bg, pixels, g_l1, g_mask, sums,
                                                                     d_rgb, d_alpha);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

extern "C" int agr_photometric_loss(const float* rgb, const float* alpha, const float* gt_rgb, const uint8_t* mask,
                                    const uint8_t* boundary, const float* bg, int64_t pixels, float w_l1, float w_mask,
                                    float* sums, float* d_rgb, float* d_alpha, void* cuda_stream) {
    return launch_photometric<float>(rgb, alpha, gt_rgb, mask, boundary, bg, pixels, w_l1, w_mask, sums, d_rgb, d_alpha, cuda_stream);
}

extern "C" int agr_photometric_loss_u8(const float* rgb, const float* alpha, const uint8_t* gt_rgb, const uint8_t* mask,
                                       const uint8_t* boundary, const float* bg, int64_t pixels, float w_l1, float w_mask,
                                       float* sums, float* d_rgb, float* d_alpha, void* cuda_stream) {
    return launch_photometric<uint8_t>(rgb, alpha, gt_rgb, mask, boundary, bg, pixels, w_l1, w_mask, sums, d_rgb, d_alpha, cuda_stream);
}
