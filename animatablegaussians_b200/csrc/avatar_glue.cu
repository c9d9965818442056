// Map -> per-Gaussian gather of AvatarNet (include/agr_avatar.h); network/avatar.py:93-124.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_avatar.h"
#include "../../include/agr_rasterizer.h"
#include "../../include/agr_styleunet.h"

namespace agr {
__device__ __forceinline__ float ld_f(const float* p) { return *p; }
__device__ __forceinline__ float ld_f(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) gather_maps_kernel(const T* __restrict__ front, const T* __restrict__ back,
                                                         const int32_t* __restrict__ half, const int32_t* __restrict__ pix,
                                                         float* __restrict__ out, int V, int S, int C, int N) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (v, n, c), c fastest
    const int64_t total = (int64_t)V * N * C;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const int64_t r = idx / C;
    const int n = (int)(r % N), v = (int)(r / N);
    const T* src = half[n] ? back : front;
    out[idx] = ld_f(src + ((int64_t)v * S * S + pix[n]) * C + c);
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_maps_kernel(const float* __restrict__ d_out, const int32_t* __restrict__ half,
                                                          const int32_t* __restrict__ pix, T* __restrict__ d_front, T* __restrict__ d_back,
                                                          int V, int S, int C, int N) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)V * N * C;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const int64_t r = idx / C;
    const int n = (int)(r % N), v = (int)(r / N);
    T* dst = half[n] ? d_back : d_front;
    st_f(dst + ((int64_t)v * S * S + pix[n]) * C + c, d_out[idx]);
}
}  // namespace agr

extern "C" {
int agr_gather_maps_forward(int32_t dtype, const void* front, const void* back, const int32_t* half, const int32_t* pix, float* out,
                            int32_t V, int32_t S, int32_t C, int32_t N, void* cuda_stream) {
    if (!front || !back || !half || !pix || !out || V < 1 || S < 1 || C < 1 || N < 0) return AGR_ERR_INVALID_ARGUMENT;
    if (N == 0) return AGR_OK;
    const int64_t total = (int64_t)V * N * C;
    const unsigned g = (unsigned)((total + 255) / 256);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (dtype == AGR_BF16) agr::gather_maps_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)front, (const __nv_bfloat16*)back, half, pix, out, V, S, C, N);
    else agr::gather_maps_kernel<float><<<g, 256, 0, s>>>((const float*)front, (const float*)back, half, pix, out, V, S, C, N);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_gather_maps_backward(int32_t dtype, const float* d_out, const int32_t* half, const int32_t* pix, void* d_front, void* d_back,
                             int32_t V, int32_t S, int32_t C, int32_t N, void* cuda_stream) {
    if (!d_out || !half || !pix || !d_front || !d_back || V < 1 || S < 1 || C < 1 || N < 0) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t bytes = (size_t)V * S * S * C * (dtype == AGR_BF16 ? 2 : 4);
    if (cudaMemsetAsync(d_front, 0, bytes, s) != cudaSuccess || cudaMemsetAsync(d_back, 0, bytes, s) != cudaSuccess) return AGR_ERR_CUDA;
    if (N == 0) return AGR_OK;
    const int64_t total = (int64_t)V * N * C;
    const unsigned g = (unsigned)((total + 255) / 256);
    if (dtype == AGR_BF16) agr::scatter_maps_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(d_out, half, pix, (__nv_bfloat16*)d_front, (__nv_bfloat16*)d_back, V, S, C, N);
    else agr::scatter_maps_kernel<float><<<g, 256, 0, s>>>(d_out, half, pix, (float*)d_front, (float*)d_back, V, S, C, N);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}
}
