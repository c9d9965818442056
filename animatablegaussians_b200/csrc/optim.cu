// Fused Adam over one flat fp32 bucket (include/agr_optim.h). Pure HBM streaming: 16 B read + 12 B (+4 B
// zero-fill of the gradient) written per parameter, 128-bit accesses, grid sized to the SM count.
#include "../../include/agr_optim.h"
#include "../../include/agr_rasterizer.h"
#include <cuda_runtime.h>

namespace agr {
__global__ void adam_tick_kernel(int32_t* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_kernel(int64_t n4, int64_t n, float4* __restrict__ p, float4* __restrict__ g,
                                                  float4* __restrict__ m, float4* __restrict__ v, float lr_c, float b1,
                                                  float b2, float eps, float inv_sqrt_bc2, float gs, int zero,
                                                  const int32_t* __restrict__ dev_step, float lr) {
    if (dev_step != nullptr) {  // graph mode: bias corrections from the device-resident step counter
        const double t = (double)*dev_step;
        lr_c = (float)((double)lr / (1.0 - pow((double)b1, t)));
        inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = p[i], G = g[i], M = m[i], V = v[i];
        float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * gs;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vv[k] = b2 * vv[k] + (1.f - b2) * gk * gk;
            pp[k] -= lr_c * mm[k] / (sqrtf(vv[k]) * inv_sqrt_bc2 + eps);
        }
        p[i] = P; m[i] = M; v[i] = V;
        if (zero) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n % 4 elements) handled by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = n4 * 4 + threadIdx.x;
        float* P = reinterpret_cast<float*>(p); float* G = reinterpret_cast<float*>(g);
        float* M = reinterpret_cast<float*>(m); float* V = reinterpret_cast<float*>(v);
        const float gk = G[i] * gs;
        M[i] = b1 * M[i] + (1.f - b1) * gk;
        V[i] = b2 * V[i] + (1.f - b2) * gk * gk;
        P[i] -= lr_c * M[i] / (sqrtf(V[i]) * inv_sqrt_bc2 + eps);
        if (zero) G[i] = 0.f;
    }
}

// ---- segmented form: per-parameter step counters and activity (torch.optim.Adam semantics for parameters without a
// gradient), learning rate / gradient scale read from device memory so a captured CUDA graph follows the schedule.
__global__ void adam_segment_tick_kernel(int S, const int32_t* __restrict__ active, int32_t* __restrict__ step, float2* __restrict__ corr,
                                         const float* __restrict__ hyper, float b1, float b2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S || !active[s]) return;
    const int32_t t = ++step[s];
    corr[s] = make_float2((float)((double)hyper[0] / (1.0 - pow((double)b1, (double)t))), (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t))));
}

// one chunk = 64 float4 = 256 parameters of ONE segment (segments are padded to chunk multiples by the host)
__global__ void __launch_bounds__(256) adam_segment_kernel(int64_t n4, float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                          float4* __restrict__ v, const int32_t* __restrict__ chunk_seg,
                                                          const int32_t* __restrict__ active, const float2* __restrict__ corr,
                                                          const float* __restrict__ hyper, float b1, float b2, float eps, int zero) {
    const float gs = hyper[1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int seg = chunk_seg[i >> 6];
        if (!active[seg]) continue;          // no gradient this step: parameter, moments and step stay untouched
        const float2 c = corr[seg];
        float4 P = p[i], G = g[i], M = m[i], V = v[i];
        float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * gs;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vv[k] = b2 * vv[k] + (1.f - b2) * gk * gk;
            pp[k] -= c.x * mm[k] / (sqrtf(vv[k]) * c.y + eps);
        }
        p[i] = P; m[i] = M; v[i] = V;
        if (zero) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
}  // namespace agr

extern "C" int agr_adam_step(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                             float beta2, float eps, int32_t step, float grad_scale, int32_t zero_grad, void* cuda_stream) {
    if (n < 0 || step < 1) return AGR_ERR_INVALID_ARGUMENT;
    if (n == 0) return AGR_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return AGR_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return AGR_ERR_INVALID_ARGUMENT;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float lr_c = (float)(lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
    agr::adam_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        n4, n, reinterpret_cast<float4*>(param), reinterpret_cast<float4*>(grad), reinterpret_cast<float4*>(exp_avg),
        reinterpret_cast<float4*>(exp_avg_sq), lr_c, beta1, beta2, eps, inv_sqrt_bc2, grad_scale, zero_grad, nullptr, lr);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

extern "C" int agr_adam_step_graph(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                                   float beta2, float eps, int32_t* device_step, float grad_scale, int32_t zero_grad,
                                   void* cuda_stream) {
    if (n < 0 || !device_step) return AGR_ERR_INVALID_ARGUMENT;
    if (n == 0) return AGR_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return AGR_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return AGR_ERR_INVALID_ARGUMENT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    agr::adam_tick_kernel<<<1, 1, 0, s>>>(device_step);
    agr::adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(
        n4, n, reinterpret_cast<float4*>(param), reinterpret_cast<float4*>(grad), reinterpret_cast<float4*>(exp_avg),
        reinterpret_cast<float4*>(exp_avg_sq), 0.f, beta1, beta2, eps, 0.f, grad_scale, zero_grad, device_step, lr);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

extern "C" int agr_adam_step_segments(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, int32_t num_segments,
                                      const int32_t* chunk_segment, const int32_t* segment_active, int32_t* segment_step,
                                      float* segment_corr, const float* hyper, float beta1, float beta2, float eps, int32_t zero_grad,
                                      void* cuda_stream) {
    if (n < 0 || (n & 255) || num_segments < 1) return AGR_ERR_INVALID_ARGUMENT;
    if (n == 0) return AGR_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !chunk_segment || !segment_active || !segment_step || !segment_corr || !hyper)
        return AGR_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return AGR_ERR_INVALID_ARGUMENT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    agr::adam_segment_tick_kernel<<<(num_segments + 255) / 256, 256, 0, s>>>(num_segments, segment_active, segment_step,
                                                                              reinterpret_cast<float2*>(segment_corr), hyper, beta1, beta2);
    agr::adam_segment_kernel<<<(unsigned)blocks, 256, 0, s>>>(n4, reinterpret_cast<float4*>(param), reinterpret_cast<float4*>(grad),
                                                            reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
                                                            chunk_segment, segment_active, reinterpret_cast<const float2*>(segment_corr), hyper,
                                                            beta1, beta2, eps, zero_grad);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}
