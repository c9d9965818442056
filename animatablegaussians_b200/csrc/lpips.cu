// Perceptual-loss kernels that are not convolutions (include/agr_lpips.h): 2x2 max-pooling of the VGG-16 trunk and the fused
// LPIPS head of one tapped layer (unit-normalise over channels, squared difference, learned channel weights, spatial mean).
// All HBM-bound streaming passes over NHWC tensors: pooling reads 4 and writes 1 (forward) / reads 5 and writes 4 (backward)
// elements per output; the head reads both feature stacks once (forward) and reads them + writes their gradients once
// (backward) — the reference materialises two normalised stacks, their difference, its square and the 1x1 conv output
// (network/lpips/lpips.py:88-103).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_lpips.h"
#include "../../include/agr_rasterizer.h"
#include "../../include/agr_styleunet.h"

namespace agr {
namespace lp {

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// ---- max pooling: one thread per (output pixel, channel pair); channels fastest -> coalesced ------------------------------
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    const int OH = H / 2, OW = W / 2;
    const int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        const T* p = x + (((int64_t)n * H + 2 * oy) * W + 2 * ox) * C + c;
        const float a = to_f(p[0]), b = to_f(p[C]), d = to_f(p[(int64_t)W * C]), e = to_f(p[(int64_t)W * C + C]);
        y[i] = from_f<T>(fmaxf(fmaxf(a, b), fmaxf(d, e)));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int N, int H,
                                                         int W, int C) {
    const int OH = H / 2, OW = W / 2;
    const int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        const int64_t o = (((int64_t)n * H + 2 * oy) * W + 2 * ox) * C + c;
        const float v[4] = {to_f(x[o]), to_f(x[o + C]), to_f(x[o + (int64_t)W * C]), to_f(x[o + (int64_t)W * C + C])};
        int arg = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if (v[k] > v[arg]) arg = k;        // strict: the first maximum keeps the gradient
        const T g = dy[i], z = from_f<T>(0.f);
        dx[o] = arg == 0 ? g : z;
        dx[o + C] = arg == 1 ? g : z;
        dx[o + (int64_t)W * C] = arg == 2 ? g : z;
        dx[o + (int64_t)W * C + C] = arg == 3 ? g : z;
    }
}

// ---- LPIPS head: one warp per pixel ----------------------------------------------------------------------------------------
constexpr float kEps = 1e-10f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256) lpips_fwd_kernel(const T* __restrict__ f, const float* __restrict__ w, int64_t pixels, int C, float scale,
                                                       float* __restrict__ out) {
    __shared__ float s_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* f0 = f;
    const T* f1 = f + pixels * C;
    float acc = 0.f;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < pixels; p += (int64_t)gridDim.x * 8) {
        const T* a = f0 + p * C;
        const T* b = f1 + p * C;
        float s0 = 0.f, s1 = 0.f;
        for (int c = lane; c < C; c += 32) { const float x0 = to_f(a[c]), x1 = to_f(b[c]); s0 += x0 * x0; s1 += x1 * x1; }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        const float i0 = 1.f / (sqrtf(s0 + kEps) + kEps), i1 = 1.f / (sqrtf(s1 + kEps) + kEps);
        float d = 0.f;
        for (int c = lane; c < C; c += 32) { const float e = to_f(a[c]) * i0 - to_f(b[c]) * i1; d += w[c] * e * e; }
        acc += d;
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        atomicAdd(out, t * scale);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) lpips_bwd_kernel(const T* __restrict__ f, const float* __restrict__ w, int64_t pixels, int C, float scale,
                                                       const float* __restrict__ g, T* __restrict__ df) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* f0 = f;
    const T* f1 = f + pixels * C;
    T* d0 = df;
    T* d1 = df + pixels * C;
    const float gs = 2.f * scale * g[0];
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < pixels; p += (int64_t)gridDim.x * 8) {
        const T* a = f0 + p * C;
        const T* b = f1 + p * C;
        float s0 = 0.f, s1 = 0.f;
        for (int c = lane; c < C; c += 32) { const float x0 = to_f(a[c]), x1 = to_f(b[c]); s0 += x0 * x0; s1 += x1 * x1; }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        const float n0 = sqrtf(s0 + kEps), n1 = sqrtf(s1 + kEps);
        const float i0 = 1.f / (n0 + kEps), i1 = 1.f / (n1 + kEps);
        // e[c] = d out / d u0[c] = -d out / d u1[c];   u = f * i,  d u[c] / d f[k] = i * delta_ck - f[c] f[k] i^2 / n
        float t0 = 0.f, t1 = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float x0 = to_f(a[c]), x1 = to_f(b[c]);
            const float e = gs * w[c] * (x0 * i0 - x1 * i1);
            t0 += e * x0; t1 += e * x1;
        }
        t0 = warp_sum(t0) * i0 * i0 / n0;
        t1 = warp_sum(t1) * i1 * i1 / n1;
        for (int c = lane; c < C; c += 32) {
            const float x0 = to_f(a[c]), x1 = to_f(b[c]);
            const float e = gs * w[c] * (x0 * i0 - x1 * i1);
            d0[p * C + c] = from_f<T>(e * i0 - x0 * t0);
            d1[p * C + c] = from_f<T>(-e * i1 + x1 * t1);
        }
    }
}

static unsigned grid_for(int64_t items, int per_block) {
    int64_t b = (items + per_block - 1) / per_block;
    const int64_t cap = 148 * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace lp
}  // namespace agr

extern "C" {

int agr_maxpool2x2_forward(int32_t dtype, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* cuda_stream) {
    using namespace agr::lp;
    if (!x || !y || N < 1 || H < 2 || W < 2 || C < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * C;
    if (dtype == AGR_BF16) maxpool_fwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
    else if (dtype == AGR_F32) maxpool_fwd_kernel<float><<<grid_for(total, 256), 256, 0, s>>>((const float*)x, (float*)y, N, H, W, C);
    else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_maxpool2x2_backward(int32_t dtype, const void* x, const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, void* cuda_stream) {
    using namespace agr::lp;
    if (!x || !dy || !dx || N < 1 || H < 2 || W < 2 || C < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t esz = dtype == AGR_BF16 ? 2 : 4;
    if ((H & 1) || (W & 1)) {   // floor mode leaves the last row / column outside every window: its gradient is zero
        if (cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * esz, s) != cudaSuccess) return AGR_ERR_CUDA;
    }
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * C;
    if (dtype == AGR_BF16) maxpool_bwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, N, H, W, C);
    else if (dtype == AGR_F32) maxpool_bwd_kernel<float><<<grid_for(total, 256), 256, 0, s>>>((const float*)x, (const float*)dy, (float*)dx, N, H, W, C);
    else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_lpips_layer_forward(int32_t dtype, const void* f, const float* w, int64_t pixels, int32_t C, float scale, float* out, void* cuda_stream) {
    using namespace agr::lp;
    if (!f || !w || !out || pixels < 1 || C < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (dtype == AGR_BF16) lpips_fwd_kernel<__nv_bfloat16><<<grid_for(pixels, 8), 256, 0, s>>>((const __nv_bfloat16*)f, w, pixels, C, scale, out);
    else if (dtype == AGR_F32) lpips_fwd_kernel<float><<<grid_for(pixels, 8), 256, 0, s>>>((const float*)f, w, pixels, C, scale, out);
    else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_lpips_layer_backward(int32_t dtype, const void* f, const float* w, int64_t pixels, int32_t C, float scale, const float* g, void* df,
                             void* cuda_stream) {
    using namespace agr::lp;
    if (!f || !w || !g || !df || pixels < 1 || C < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (dtype == AGR_BF16) lpips_bwd_kernel<__nv_bfloat16><<<grid_for(pixels, 8), 256, 0, s>>>((const __nv_bfloat16*)f, w, pixels, C, scale, g, (__nv_bfloat16*)df);
    else if (dtype == AGR_F32) lpips_bwd_kernel<float><<<grid_for(pixels, 8), 256, 0, s>>>((const float*)f, w, pixels, C, scale, g, (float*)df);
    else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // extern "C"
