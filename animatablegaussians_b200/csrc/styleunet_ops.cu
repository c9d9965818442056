// StyleUNet glue operators on NHWC activations (fp32 / bf16), include/agr_styleunet.h.
//
// All of these are HBM-streaming kernels: each output element is produced from a handful of neighbouring
// inputs, so the design rule is "touch HBM once, 128-bit accesses along the channel axis".
//   upfirdn2d  : reference upfirdn2d_kernel.cu:107-207 stages an input tile + taps in shared memory for a
//                (major,H,W,minor=1) layout; in NHWC the channel axis is contiguous, a thread owns one output
//                pixel x 8 (bf16) / 4 (fp32) channels and the <=16 taps hit L1/L2 (neighbouring threads share them).
//   haar       : the four 2x2 sub-band filters of dual_styleunet.py:374-425 applied in ONE pass (reference: 4
//                upfirdn2d launches + cat / 4 launches + 3 adds).
//   bias_act   : fused_bias_act_kernel.cu:18-65 (act=3) fused with NoiseInjection (dual_styleunet.py:303-313); the
//                backward also produces the bias / noise-weight reductions the reference computes with a
//                separate .sum().
//   modweight  : dual_styleunet.py:256-265 in one pass per layer, writing the KRSC layout the implicit GEMM wants.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_rasterizer.h"
#include "../../include/agr_styleunet.h"

namespace agr {

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; using V = float4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int N = 8; using V = uint4; };

// Vectors are loaded BY VALUE through ld_vec (one LDG.128): unpacking through a reference to global memory lets the
// compiler split the access into four 32-bit loads, which quarters the sector efficiency of every warp-level load.
template <typename V> __device__ __forceinline__ V ld_vec(const void* p) { return __ldg(reinterpret_cast<const V*>(p)); }
__device__ __forceinline__ void unpack(float4 v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
__device__ __forceinline__ void unpack(uint4 v, float* f) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void pack(const float* f, float4& v) { v = make_float4(f[0], f[1], f[2], f[3]); }
__device__ __forceinline__ void pack(const float* f, uint4& v) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// ------------------------------------------------------------------------------------------------ upfirdn2d
struct FirParams {
    float taps[64];
    int kh, kw, up, down, pad_x0, pad_y0;
};

// UP / DOWN / K are compile-time (0 = generic run-time values): the zero-inserted taps of an upsampling filter are
// skipped at compile time (4 of 16 loads survive for up=2), all index divisions strength-reduce, and the tap loops unroll.
template <typename T, bool VECTOR, int UP_, int DOWN_, int K_>
__global__ void __launch_bounds__(256) upfirdn_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                                                     int outH, int outW, FirParams fp) {
    constexpr int VN = VECTOR ? VecOf<T>::N : 1;
    const int up = UP_ ? UP_ : fp.up, down = DOWN_ ? DOWN_ : fp.down;
    const int kh = K_ ? K_ : fp.kh, kw = K_ ? K_ : fp.kw;
    const int cv = C / VN;
    const int64_t total = (int64_t)N * outH * outW * cv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cv);
    int64_t p = idx / cv;
    const int ox = (int)(p % outW); p /= outW;
    const int oy = (int)(p % outH);
    const int n = (int)(p / outH);
    float acc[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) acc[i] = 0.f;
    const int by = oy * down - fp.pad_y0, bx = ox * down - fp.pad_x0;
    // first tap whose upsampled coordinate is a multiple of `up`:  (b + k) % up == 0
    const int ky0 = (up == 1) ? 0 : ((up - (by % up + up) % up) % up);
    const int kx0 = (up == 1) ? 0 : ((up - (bx % up + up) % up) % up);
    const T* xn = x + (int64_t)n * H * W * C + (int64_t)c * VN;
#pragma unroll
    for (int kyi = 0; kyi < (K_ ? (K_ + (UP_ ? UP_ : 1) - 1) / (UP_ ? UP_ : 1) : 64); ++kyi) {
        const int ky = ky0 + kyi * up;
        if (ky >= kh) break;
        const int uy = by + ky;
        const int iy = (up == 1) ? uy : uy / up;   // uy % up == 0 by construction (uy may be negative: C division is fine, multiple of up)
        if (uy < 0 || iy >= H) continue;
#pragma unroll
        for (int kxi = 0; kxi < (K_ ? (K_ + (UP_ ? UP_ : 1) - 1) / (UP_ ? UP_ : 1) : 64); ++kxi) {
            const int kx = kx0 + kxi * up;
            if (kx >= kw) break;
            const int ux = bx + kx;
            const int ix = (up == 1) ? ux : ux / up;
            if (ux < 0 || ix >= W) continue;
            const float t = fp.taps[ky * kw + kx];
            const T* src = xn + ((int64_t)iy * W + ix) * C;
            if (VECTOR) {
                float f[VN];
                unpack(ld_vec<typename VecOf<T>::V>(src), f);
#pragma unroll
                for (int i = 0; i < VN; ++i) acc[i] += t * f[i];
            } else {
                acc[0] += t * to_f(src[0]);
            }
        }
    }
    T* dst = y + (((int64_t)n * outH + oy) * outW + ox) * C + (int64_t)c * VN;
    if (VECTOR) {
        typename VecOf<T>::V v;
        pack(acc, v);
        *reinterpret_cast<typename VecOf<T>::V*>(dst) = v;
    } else {
        dst[0] = from_f<T>(acc[0]);
    }
}

// Blur (up = down = 1, K x K taps) on a channel-vectorisable tensor.  The one-pixel-per-thread kernel above issues K*K
// vector loads + index arithmetic per output and is issue-bound (r01 ncu: ~1.1 TB/s at 16x512x512x64).  Here a
// thread owns ROWS vertically adjacent outputs of one channel vector and walks the ROWS+K-1 input rows once: each
// loaded vector feeds up to K outputs ((ROWS+K-1)*K/ROWS = 7 loads per output for K = ROWS = 4 instead of 16).
// Every output still accumulates its taps in (ky, kx) ascending order, so results are bit-identical to upfirdn_kernel.
template <typename T, int K, int ROWS>
__global__ void __launch_bounds__(256) blur_strip_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                                                        int outH, int outW, FirParams fp) {
    constexpr int VN = VecOf<T>::N;
    using V = typename VecOf<T>::V;
    const int cv = C / VN;
    const int strips = (outH + ROWS - 1) / ROWS;
    const int64_t total = (int64_t)N * strips * outW * cv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cv);
    int64_t p = idx / cv;
    const int ox = (int)(p % outW); p /= outW;
    const int st = (int)(p % strips);
    const int n = (int)(p / strips);
    const int oy0 = st * ROWS;
    float acc[ROWS][VN];
#pragma unroll
    for (int o = 0; o < ROWS; ++o)
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[o][i] = 0.f;
    const int by = oy0 - fp.pad_y0, bx = ox - fp.pad_x0;
    const T* xn = x + (int64_t)n * H * W * C + (int64_t)c * VN;
#pragma unroll
    for (int r = 0; r < ROWS + K - 1; ++r) {
        const int iy = by + r;
        if (iy < 0 || iy >= H) continue;
        const T* xr = xn + (int64_t)iy * W * C;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int ix = bx + kx;
            if (ix < 0 || ix >= W) continue;
            float f[VN];
            unpack(ld_vec<V>(xr + (int64_t)ix * C), f);
#pragma unroll
            for (int o = 0; o < ROWS; ++o) {
                const int ky = r - o;   // compile-time after unrolling
                if (ky < 0 || ky >= K) continue;
                const float t = fp.taps[ky * K + kx];
#pragma unroll
                for (int i = 0; i < VN; ++i) acc[o][i] += t * f[i];
            }
        }
    }
    T* dst = y + (((int64_t)n * outH + oy0) * outW + ox) * C + (int64_t)c * VN;
#pragma unroll
    for (int o = 0; o < ROWS; ++o) {
        if (oy0 + o >= outH) break;
        V v;
        pack(acc[o], v);
        *reinterpret_cast<V*>(dst + (int64_t)o * outW * C) = v;
    }
}

// ------------------------------------------------------------------------------------------------ haar
// analysis: x (H,W,C) -> y (H/2,W/2,4C);  synthesis: x (H,W,4C) -> y (2H,2W,C).  Orthonormal: each is the
// other's adjoint, so backward(dwt) = synthesis and backward(iwt) = analysis.
template <typename T, bool VECTOR, bool ANALYSIS>
__global__ void __launch_bounds__(256) haar_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    constexpr int VN = VECTOR ? VecOf<T>::N : 1;
    // C = channels of the SPATIALLY LARGER tensor (the image side); Hs,Ws = size of the sub-band side
    const int Hs = ANALYSIS ? H / 2 : H, Ws = ANALYSIS ? W / 2 : W;
    const int Ci = ANALYSIS ? C : C / 4;
    const int cv = Ci / VN;
    const int64_t total = (int64_t)N * Hs * Ws * cv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cv) * VN;
    int64_t p = idx / cv;
    const int j = (int)(p % Ws); p /= Ws;
    const int i = (int)(p % Hs);
    const int n = (int)(p / Hs);
    const int Hi = Hs * 2, Wi = Ws * 2;
    float a[4][VN];  // image pixels (0,0) (0,1) (1,0) (1,1) or sub-bands ll lh hl hh
    using V = typename VecOf<T>::V;
    auto ld = [&](const T* ptr, float* f) {
        if (VECTOR) unpack(ld_vec<V>(ptr), f); else f[0] = to_f(ptr[0]);
    };
    auto st = [&](T* ptr, const float* f) {
        if (VECTOR) { V v; pack(f, v); *reinterpret_cast<V*>(ptr) = v; } else ptr[0] = from_f<T>(f[0]);
    };
    if (ANALYSIS) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ld(x + (((int64_t)n * Hi + 2 * i + (q >> 1)) * Wi + 2 * j + (q & 1)) * Ci + c, a[q]);
        float o[4][VN];
#pragma unroll
        for (int v = 0; v < VN; ++v) {
            const float x00 = a[0][v], x01 = a[1][v], x10 = a[2][v], x11 = a[3][v];
            o[0][v] = 0.5f * (x00 + x01 + x10 + x11);
            o[1][v] = 0.5f * (x00 + x01 - x10 - x11);
            o[2][v] = 0.5f * (x00 - x01 + x10 - x11);
            o[3][v] = 0.5f * (x00 - x01 - x10 + x11);
        }
        T* dst = y + (((int64_t)n * Hs + i) * Ws + j) * (4 * Ci);
#pragma unroll
        for (int b = 0; b < 4; ++b) st(dst + b * Ci + c, o[b]);
    } else {
        const T* src = x + (((int64_t)n * Hs + i) * Ws + j) * (4 * Ci);
#pragma unroll
        for (int b = 0; b < 4; ++b) ld(src + b * Ci + c, a[b]);
        float o[4][VN];
#pragma unroll
        for (int v = 0; v < VN; ++v) {
            const float ll = a[0][v], lh = a[1][v], hl = a[2][v], hh = a[3][v];
            o[0][v] = 0.5f * (ll + lh + hl + hh);
            o[1][v] = 0.5f * (ll + lh - hl - hh);
            o[2][v] = 0.5f * (ll - lh + hl - hh);
            o[3][v] = 0.5f * (ll - lh - hl + hh);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            st(y + (((int64_t)n * Hi + 2 * i + (q >> 1)) * Wi + 2 * j + (q & 1)) * Ci + c, o[q]);
    }
}

// ------------------------------------------------------------------------------------------------ wavelet upsample
// dwt(upsample_2x(iwt(skip))) of the ToRGB skip path as ONE pass.  All three stages are linear and shift-invariant up
// to output parity, so the chain collapses (on the host, include/agr_styleunet.h) into
//   y[2m+pi, 2n+pj, bo] = sum_{bi,di,dj} K[pi][pj][bo][bi][di][dj] * x[m+di+pi-1, n+dj+pj-1, bi]        (zero outside)
// per colour channel.  A thread owns one skip pixel x one colour channel: 36 loads -> 2x2 output pixels x 4 bands.
// The unfused chain writes and re-reads a (4h,4w,3) image with 2-byte accesses (r01 ncu: 660 us per call at V=16).
struct WaveletTaps { float k[256]; };

template <typename T>
__global__ void __launch_bounds__(256) wavelet_up_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int h, int w, int Ci,
                                                            WaveletTaps kp) {
    const int64_t total = (int64_t)N * h * w * Ci;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % Ci);
    int64_t p = idx / Ci;
    const int n = (int)(p % w); p /= w;
    const int m = (int)(p % h);
    const int b = (int)(p / h);
    const int C4 = 4 * Ci;
    float v[3][3][4];
#pragma unroll
    for (int dm = 0; dm < 3; ++dm)
#pragma unroll
        for (int dn = 0; dn < 3; ++dn) {
            const int yy = m + dm - 1, xx = n + dn - 1;
            const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
            const T* src = x + (((int64_t)b * h + yy) * w + xx) * C4 + c;
#pragma unroll
            for (int bi = 0; bi < 4; ++bi) v[dm][dn][bi] = in ? to_f(src[bi * Ci]) : 0.f;
        }
#pragma unroll
    for (int pi = 0; pi < 2; ++pi)
#pragma unroll
        for (int pj = 0; pj < 2; ++pj) {
            T* dst = y + (((int64_t)b * 2 * h + 2 * m + pi) * (2 * w) + 2 * n + pj) * C4 + c;
#pragma unroll
            for (int bo = 0; bo < 4; ++bo) {
                float acc = 0.f;
#pragma unroll
                for (int bi = 0; bi < 4; ++bi)
#pragma unroll
                    for (int di = 0; di < 2; ++di)
#pragma unroll
                        for (int dj = 0; dj < 2; ++dj)
                            acc += kp.k[((((pi * 2 + pj) * 4 + bo) * 4 + bi) * 2 + di) * 2 + dj] * v[di + pi][dj + pj][bi];
                dst[bo * Ci] = from_f<T>(acc);
            }
        }
}

// adjoint: dx[m, n, bi] = sum_{bo,a,b} KT[bi][bo][a][b] * g[2m-1+a, 2n-1+b, bo]   (gather form, no atomics)
template <typename T>
__global__ void __launch_bounds__(256) wavelet_up_bwd_kernel(const T* __restrict__ g, T* __restrict__ dx, int N, int h, int w, int Ci,
                                                            WaveletTaps kp) {
    const int64_t total = (int64_t)N * h * w * Ci;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % Ci);
    int64_t p = idx / Ci;
    const int n = (int)(p % w); p /= w;
    const int m = (int)(p % h);
    const int b = (int)(p / h);
    const int C4 = 4 * Ci, H2 = 2 * h, W2 = 2 * w;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = 2 * m - 1 + a;
        if (i < 0 || i >= H2) continue;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            const int j = 2 * n - 1 + bb;
            if (j < 0 || j >= W2) continue;
            const T* src = g + (((int64_t)b * H2 + i) * W2 + j) * C4 + c;
#pragma unroll
            for (int bo = 0; bo < 4; ++bo) {
                const float gv = to_f(src[bo * Ci]);
#pragma unroll
                for (int bi = 0; bi < 4; ++bi) acc[bi] += kp.k[((bi * 4 + bo) * 4 + a) * 4 + bb] * gv;
            }
        }
    }
    T* dst = dx + (((int64_t)b * h + m) * w + n) * C4 + c;
#pragma unroll
    for (int bi = 0; bi < 4; ++bi) dst[bi * Ci] = from_f<T>(acc[bi]);
}

// ------------------------------------------------------------------------------------------------ bias_act
constexpr float kSqrt2 = 1.4142135623730951f;

template <typename T, bool VECTOR>
__global__ void __launch_bounds__(256) bias_act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t pixels, int C,
                                                          const float* __restrict__ bias, const float* __restrict__ noise,
                                                          const float* __restrict__ noise_w, int64_t nper, int activate) {
    constexpr int VN = VECTOR ? VecOf<T>::N : 1;
    const int cv = C / VN;
    const int64_t total = pixels * cv;
    const float nw = (noise && noise_w) ? noise_w[0] : 0.f;
    const float gain = activate == 1 ? kSqrt2 : 1.f;   // 1: FusedLeakyReLU (x sqrt2), 2: plain LeakyReLU(0.2), 3: ReLU
    const float neg_slope = activate == 3 ? 0.f : 0.2f;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VN;
        const int64_t p = idx / cv;
        const float add = noise ? nw * noise[p % nper] : 0.f;
        float f[VN];
        if (VECTOR) unpack(ld_vec<typename VecOf<T>::V>(x + p * C + c), f); else f[0] = to_f(x[p * C + c]);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            float v = f[i] + add + (bias ? bias[c + i] : 0.f);
            if (activate) v = (v > 0.f ? v : v * neg_slope) * gain;
            f[i] = v;
        }
        if (VECTOR) { typename VecOf<T>::V v; pack(f, v); *reinterpret_cast<typename VecOf<T>::V*>(y + p * C + c) = v; }
        else y[p * C + c] = from_f<T>(f[0]);
    }
}

// Each block owns a slab of pixels; thread = (channel vector, pixel lane). Per-channel partial sums are reduced
// in shared memory and added with one atomic per channel per block.
template <typename T, bool VECTOR>
__global__ void __launch_bounds__(256) bias_act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx,
                                                          int64_t pixels, int C, const float* __restrict__ noise, int64_t nper,
                                                          float* __restrict__ d_bias, float* __restrict__ d_noise_w,
                                                          int activate, int pixels_per_block) {
    constexpr int VN = VECTOR ? VecOf<T>::N : 1;
    extern __shared__ float s_red[];  // C floats (bias) + 1 (noise)
    const int cv = C / VN;
    for (int i = threadIdx.x; i < C + 1; i += blockDim.x) s_red[i] = 0.f;
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * pixels_per_block;
    const int64_t p1 = min(pixels, p0 + pixels_per_block);
    const int lanes = blockDim.x / cv > 0 ? blockDim.x / cv : 1;   // pixel lanes per block
    const int my_c = (threadIdx.x % cv) * VN;
    const int my_lane = threadIdx.x / cv;
    float bsum[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) bsum[i] = 0.f;
    float nsum = 0.f;
    const float gain = activate == 1 ? kSqrt2 : 1.f;
    const float neg_slope = activate == 3 ? 0.f : 0.2f;
    if (my_lane < lanes) {
        for (int64_t p = p0 + my_lane; p < p1; p += lanes) {
            float g[VN], o[VN];
            if (VECTOR) {
                unpack(ld_vec<typename VecOf<T>::V>(dy + p * C + my_c), g);
                if (activate) unpack(ld_vec<typename VecOf<T>::V>(y + p * C + my_c), o);
            } else {
                g[0] = to_f(dy[p * C + my_c]);
                if (activate) o[0] = to_f(y[p * C + my_c]);
            }
            float psum = 0.f;
#pragma unroll
            for (int i = 0; i < VN; ++i) {
                if (activate) g[i] *= (o[i] > 0.f ? gain : neg_slope * gain);
                bsum[i] += g[i];
                psum += g[i];
            }
            if (noise) nsum += psum * noise[p % nper];
            if (dx) {   // dx == NULL: identity activation, the caller keeps using dy (only the reductions are wanted)
                if (VECTOR) { typename VecOf<T>::V v; pack(g, v); *reinterpret_cast<typename VecOf<T>::V*>(dx + p * C + my_c) = v; }
                else dx[p * C + my_c] = from_f<T>(g[0]);
            }
        }
        if (d_bias) {
#pragma unroll
            for (int i = 0; i < VN; ++i) atomicAdd(&s_red[my_c + i], bsum[i]);
        }
        if (d_noise_w && noise) atomicAdd(&s_red[C], nsum);
    }
    __syncthreads();
    if (d_bias) for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&d_bias[i], s_red[i]);
    if (d_noise_w && noise && threadIdx.x == 0) atomicAdd(d_noise_w, s_red[C]);
}

// ---- the reference's `fused.fused_bias_act` op, element for element (fused_bias_act_kernel.cu:18-65): any contiguous
// layout, bias indexed (i / step_b) % size_b, act in {1 linear, 3 lrelu}, grad in {0, 1, 2}.  The StyleUNet of this package
// uses the NHWC kernels above; this entry point exists so that the reference's own fused_act.py runs on this library.
template <typename T>
__global__ void __launch_bounds__(256) fused_bias_act_ref_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ b,
                                                                const T* __restrict__ ref, int act, int grad, float alpha, float scale,
                                                                int64_t size_x, int64_t step_b, int64_t size_b) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size_x; i += (int64_t)gridDim.x * blockDim.x) {
        float v = to_f(x[i]);
        if (b) v += to_f(b[(i / step_b) % size_b]);
        const float r = ref ? to_f(ref[i]) : 0.f;
        float y;
        switch (act * 10 + grad) {
            default:
            case 10: case 11: y = v; break;
            case 12: case 32: y = 0.f; break;
            case 30: y = v > 0.f ? v : v * alpha; break;
            case 31: y = r > 0.f ? v : v * alpha; break;
        }
        out[i] = from_f<T>(y * scale);
    }
}

// ------------------------------------------------------------------------------------------------ modweight
// One block per output channel. w: (Cout, Cin, k, k) fp32.  Row bodies are shared by the per-layer kernels and the
// grouped kernels (one launch for many layers).
//
// SMEM = true (rows of <= 12288 floats, 16-byte aligned): the (Cin,k,k) master-weight row is staged with coalesced
// float4 loads issued back to back (one DRAM round trip instead of n/256 dependent ones), the strided [ci*kk+t]
// accesses then hit shared memory (stride kk is odd or 1: conflict-free), and the backward builds the d_w row in
// place and writes it back coalesced.  SMEM = false reads / writes global memory directly (any row).  Both variants
// perform the same per-element arithmetic in the same order.
__device__ __forceinline__ void stage_row(const float* __restrict__ src, float* dst, int n4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int base = 0; base < n4; base += 4 * 256) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int j = base + k * 256 + threadIdx.x; if (j < n4) v[k] = __ldg(s4 + j); }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int j = base + k * 256 + threadIdx.x; if (j < n4) d4[j] = v[k]; }
    }
}

__device__ __forceinline__ float block_sum_256(float v, float* s_part, float* s_out) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) *s_out = t;
    }
    __syncthreads();
    return *s_out;
}

template <typename T, bool SMEM>
__device__ __forceinline__ void modweight_fwd_row(const float* __restrict__ w, const float* __restrict__ s, float scale, int Cout,
                                                  int Cin, int kk, int demodulate, int transpose_io, T* __restrict__ w_out,
                                                  float* __restrict__ demod_out, int co, float* sw, float* s_part, float* s_val) {
    const int n = Cin * kk;
    const float* wr = w + (size_t)co * n;
    if (SMEM) { stage_row(wr, sw, n >> 2); __syncthreads(); }
    const float* row = SMEM ? sw : wr;
    float d = 1.f;
    if (demodulate) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float u = scale * row[i] * s[i / kk];
            acc += u * u;
        }
        const float tot = block_sum_256(acc, s_part, s_val);
        d = rsqrtf(tot + 1e-8f);
        if (demod_out && threadIdx.x == 0) demod_out[co] = d;
    }
    // output index: KRSC [co][t][ci]  or (transpose_io) [ci][t][co]
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int t = i / Cin, ci = i - t * Cin;  // iterate with ci fastest for coalesced KRSC writes
        const float u = scale * row[ci * kk + t] * s[ci] * d;
        const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
        w_out[o] = from_f<T>(u);
    }
}

template <typename T, bool SMEM>
__device__ __forceinline__ void modweight_bwd_row(const float* __restrict__ w, const float* __restrict__ s, float scale, int Cout,
                                                  int Cin, int kk, int demodulate, int transpose_io, const T* __restrict__ d_wout,
                                                  const float* __restrict__ demod, float* __restrict__ d_w, float* __restrict__ d_s,
                                                  int co, float* sw, float* s_part, float* s_val) {
    const int n = Cin * kk;
    const float* wr = w + (size_t)co * n;
    if (SMEM) { stage_row(wr, sw, n >> 2); __syncthreads(); }
    const float* row = SMEM ? sw : wr;
    const float d = demodulate ? demod[co] : 1.f;
    float dot = 0.f;  // sum_i dW'_i * u_i
    if (demodulate) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int t = i / Cin, ci = i - t * Cin;
            const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
            dot += to_f(d_wout[o]) * (scale * row[ci * kk + t] * s[ci]);
        }
        dot = block_sum_256(dot, s_part, s_val);   // its barriers also order the reads of sw before the overwrite below
    }
    const float d3dot = demodulate ? d * d * d * dot : 0.f;
    float* orow = SMEM ? sw : d_w + (size_t)co * n;
    // thread owns input channels ci = threadIdx.x, +blockDim.x, ... so d_s needs one atomic per (co, ci)
    for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
        float ds = 0.f;
        const float sc = s[ci];
        for (int t = 0; t < kk; ++t) {
            const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
            const float wv = row[ci * kk + t];
            const float u = scale * wv * sc;
            const float du = d * to_f(d_wout[o]) - u * d3dot;   // dL/du
            orow[ci * kk + t] = du * scale * sc;                // (SMEM: element owned by this thread only)
            ds += du * scale * wv;
        }
        if (d_s) atomicAdd(&d_s[ci], ds);
    }
    if (SMEM) {
        __syncthreads();
        float4* out4 = reinterpret_cast<float4*>(d_w + (size_t)co * n);
        const float4* in4 = reinterpret_cast<const float4*>(sw);
        for (int j = threadIdx.x; j < (n >> 2); j += blockDim.x) out4[j] = in4[j];
    }
}

template <typename T, bool SMEM>
__global__ void __launch_bounds__(256) modweight_fwd_kernel(const float* __restrict__ w, const float* __restrict__ s, float scale,
                                                           int Cout, int Cin, int kk, int demodulate, int transpose_io,
                                                           T* __restrict__ w_out, float* __restrict__ demod_out) {
    extern __shared__ float4 s_row4[];
    __shared__ float s_part[32];
    __shared__ float s_val;
    modweight_fwd_row<T, SMEM>(w, s, scale, Cout, Cin, kk, demodulate, transpose_io, w_out, demod_out, blockIdx.x,
                               reinterpret_cast<float*>(s_row4), s_part, &s_val);
}

template <typename T, bool SMEM>
__global__ void __launch_bounds__(256) modweight_bwd_kernel(const float* __restrict__ w, const float* __restrict__ s, float scale,
                                                           int Cout, int Cin, int kk, int demodulate, int transpose_io,
                                                           const T* __restrict__ d_wout, const float* __restrict__ demod,
                                                           float* __restrict__ d_w, float* __restrict__ d_s) {
    extern __shared__ float4 s_row4[];
    __shared__ float s_part[32];
    __shared__ float s_val;
    modweight_bwd_row<T, SMEM>(w, s, scale, Cout, Cin, kk, demodulate, transpose_io, d_wout, demod, d_w, d_s, blockIdx.x,
                               reinterpret_cast<float*>(s_row4), s_part, &s_val);
}

// Grouped form: every conv layer of a U-Net in a few launches (the per-layer kernels are latency-bound: 58 layers x 3
// nets x (fwd + bwd) = 348 launches of 3-15 us moving ~4 GB in total).  Layer descriptors travel as kernel parameters;
// a block finds its layer by scanning the block-offset table.
constexpr int kModGroup = 40;
struct ModEntry {
    const float* w; const float* s; const float* demod_in; float* demod_out;
    const void* d_wout; void* w_out; float* d_w; float* d_s;
    float scale; int Cout, Cin, kk, flags, block_begin;   // flags: 1 demodulate, 2 transpose_io, 4 row fits shared memory
};
struct ModGroup { ModEntry e[kModGroup]; int count; };

template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(256) modweight_group_kernel(const __grid_constant__ ModGroup g) {
    extern __shared__ float4 s_row4[];
    __shared__ float s_part[32];
    __shared__ float s_val;
    int e = 0;
    while (e + 1 < g.count && (int)blockIdx.x >= g.e[e + 1].block_begin) ++e;
    const ModEntry& L = g.e[e];
    const int co = blockIdx.x - L.block_begin;
    float* sw = reinterpret_cast<float*>(s_row4);
    const int demod = L.flags & 1, tr = (L.flags >> 1) & 1;
    if (!BACKWARD) {
        if (L.flags & 4) modweight_fwd_row<T, true>(L.w, L.s, L.scale, L.Cout, L.Cin, L.kk, demod, tr, static_cast<T*>(L.w_out), L.demod_out, co, sw, s_part, &s_val);
        else modweight_fwd_row<T, false>(L.w, L.s, L.scale, L.Cout, L.Cin, L.kk, demod, tr, static_cast<T*>(L.w_out), L.demod_out, co, sw, s_part, &s_val);
    } else {
        if (L.flags & 4) modweight_bwd_row<T, true>(L.w, L.s, L.scale, L.Cout, L.Cin, L.kk, demod, tr, static_cast<const T*>(L.d_wout), L.demod_in, L.d_w, L.d_s, co, sw, s_part, &s_val);
        else modweight_bwd_row<T, false>(L.w, L.s, L.scale, L.Cout, L.Cin, L.kk, demod, tr, static_cast<const T*>(L.d_wout), L.demod_in, L.d_w, L.d_s, co, sw, s_part, &s_val);
    }
}

// ------------------------------------------------------------------------------------------------ EqualLinear (style vector)
// forward: one warp per output row (`blk` = block index within the layer).
__device__ __forceinline__ void equal_linear_fwd_body(const float* __restrict__ w, const float* __restrict__ bias,
                                                      const float* __restrict__ x, float scale, float lr_mul, int out_dim,
                                                      int in_dim, float* __restrict__ y, int blk) {
    const int row = (blk * (int)blockDim.x + (int)threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= out_dim) return;
    const float* wr = w + (size_t)row * in_dim;
    float acc = 0.f;
#pragma unroll 4
    for (int i = lane; i < in_dim; i += 32) acc += wr[i] * x[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = (bias ? lr_mul * bias[row] : 0.f) + scale * acc;
}

// backward: a block owns ROWS output rows and walks the columns; d_w rows are written coalesced, the column sums of
// the block's rows go to d_x with one atomic per (block, column).
constexpr int kLinRows = 4;
__device__ __forceinline__ void equal_linear_bwd_body(const float* __restrict__ w, const float* __restrict__ x,
                                                      const float* __restrict__ dy, float scale, float lr_mul, int out_dim,
                                                      int in_dim, float* __restrict__ d_w, float* __restrict__ d_bias,
                                                      float* __restrict__ d_x, int blk) {
    const int j0 = blk * kLinRows;
    float g[kLinRows];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) g[r] = (j0 + r < out_dim) ? dy[j0 + r] : 0.f;
    if (d_bias && threadIdx.x < kLinRows && j0 + threadIdx.x < out_dim) d_bias[j0 + threadIdx.x] = lr_mul * dy[j0 + threadIdx.x];
    for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
        const float xi = x[i];
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) {
            if (j0 + r < out_dim) {
                const size_t o = (size_t)(j0 + r) * in_dim + i;
                acc += w[o] * g[r];
                d_w[o] = scale * g[r] * xi;
            }
        }
        if (d_x) atomicAdd(&d_x[i], scale * acc);
    }
}

__global__ void __launch_bounds__(256) equal_linear_fwd_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                              const float* __restrict__ x, float scale, float lr_mul, int out_dim,
                                                              int in_dim, float* __restrict__ y) {
    equal_linear_fwd_body(w, bias, x, scale, lr_mul, out_dim, in_dim, y, blockIdx.x);
}

__global__ void __launch_bounds__(256) equal_linear_bwd_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                              const float* __restrict__ dy, float scale, float lr_mul, int out_dim,
                                                              int in_dim, float* __restrict__ d_w, float* __restrict__ d_bias,
                                                              float* __restrict__ d_x) {
    equal_linear_bwd_body(w, x, dy, scale, lr_mul, out_dim, in_dim, d_w, d_bias, d_x, blockIdx.x);
}

// Grouped form (all modulation layers of a U-Net in one launch), descriptors as kernel parameters like ModGroup.
constexpr int kLinGroup = 40;
struct LinEntry {
    const float* w; const float* bias; const float* x; const float* dy;
    float* y; float* d_w; float* d_bias; float* d_x;
    float scale, lr_mul; int out_dim, in_dim, block_begin, pad_;
};
struct LinGroup { LinEntry e[kLinGroup]; int count; };

template <bool BACKWARD>
__global__ void __launch_bounds__(256) equal_linear_group_kernel(const __grid_constant__ LinGroup g) {
    int e = 0;
    while (e + 1 < g.count && (int)blockIdx.x >= g.e[e + 1].block_begin) ++e;
    const LinEntry& L = g.e[e];
    const int blk = blockIdx.x - L.block_begin;
    if (BACKWARD) equal_linear_bwd_body(L.w, L.x, L.dy, L.scale, L.lr_mul, L.out_dim, L.in_dim, L.d_w, L.d_bias, L.d_x, blk);
    else equal_linear_fwd_body(L.w, L.bias, L.x, L.scale, L.lr_mul, L.out_dim, L.in_dim, L.y, blk);
}

template <typename T>
__global__ void __launch_bounds__(256) sum_batch_kernel(const T* __restrict__ x, T* __restrict__ y, int V, int64_t nvec) {
    constexpr int VN = VecOf<T>::N;
    using Vt = typename VecOf<T>::V;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float acc[VN];
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[k] = 0.f;
        for (int v = 0; v < V; ++v) {
            float f[VN];
            unpack(ld_vec<Vt>(reinterpret_cast<const Vt*>(x) + (int64_t)v * nvec + i), f);
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[k] += f[k];
        }
        Vt o; pack(acc, o);
        reinterpret_cast<Vt*>(y)[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------ bilinear 2x (+ add)
// PyTorch's align_corners=False source index: max(0.5 * (dst + 0.5) - 0.5, 0); x1 = min(x0 + 1, in - 1).
__device__ __forceinline__ void bil_src(int dst, int in, int& i0, int& i1, float& lam) {
    const float src = fmaxf(0.5f * ((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = (int)src;
    i1 = min(i0 + 1, in - 1);
    lam = src - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256) bilinear2x_add_kernel(const float* __restrict__ vf, const T* __restrict__ base, T* __restrict__ y,
                                                            int V, int Vb, int h, int w, int C) {
    constexpr int VN = 4;   // 4 channels per thread (float4 of vf)
    const int H = 2 * h, W = 2 * w, cv = C / VN;
    const int64_t total = (int64_t)V * H * W * cv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cv) * VN;
    int64_t p = idx / cv;
    const int ox = (int)(p % W); p /= W;
    const int oy = (int)(p % H);
    const int v = (int)(p / H);
    int y0, y1, x0, x1; float ly, lx;
    bil_src(oy, h, y0, y1, ly);
    bil_src(ox, w, x0, x1, lx);
    const float* f = vf + (int64_t)v * h * w * C + c;
    const float4 a = *reinterpret_cast<const float4*>(f + ((int64_t)y0 * w + x0) * C);
    const float4 b = *reinterpret_cast<const float4*>(f + ((int64_t)y0 * w + x1) * C);
    const float4 cc = *reinterpret_cast<const float4*>(f + ((int64_t)y1 * w + x0) * C);
    const float4 d = *reinterpret_cast<const float4*>(f + ((int64_t)y1 * w + x1) * C);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    float r[4] = {w00 * a.x + w01 * b.x + w10 * cc.x + w11 * d.x, w00 * a.y + w01 * b.y + w10 * cc.y + w11 * d.y,
                  w00 * a.z + w01 * b.z + w10 * cc.z + w11 * d.z, w00 * a.w + w01 * b.w + w10 * cc.w + w11 * d.w};
    const int64_t pix = ((int64_t)oy * W + ox) * C + c;
    const T* bp = base + (Vb == 1 ? 0 : (int64_t)v * H * W * C) + pix;
    T* yp = y + (int64_t)v * H * W * C + pix;
#pragma unroll
    for (int i = 0; i < 4; ++i) yp[i] = from_f<T>(to_f(from_f<T>(r[i])) + to_f(bp[i]));   // resize result rounds to T first, like .to(dtype)
}

template <typename T>
__global__ void __launch_bounds__(256) bilinear2x_bwd_kernel(const T* __restrict__ g, float* __restrict__ d_vf, int V, int h, int w, int C) {
    constexpr int VN = 4;
    const int H = 2 * h, W = 2 * w, cv = C / VN;
    const int64_t total = (int64_t)V * h * w * cv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cv) * VN;
    int64_t p = idx / cv;
    const int sx = (int)(p % w); p /= w;
    const int sy = (int)(p % h);
    const int v = (int)(p / h);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const T* gv = g + (int64_t)v * H * W * C + c;
    for (int oy = max(2 * sy - 1, 0); oy <= min(2 * sy + 2, H - 1); ++oy) {
        int y0, y1; float ly;
        bil_src(oy, h, y0, y1, ly);
        const float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
        if (wy == 0.f) continue;
        for (int ox = max(2 * sx - 1, 0); ox <= min(2 * sx + 2, W - 1); ++ox) {
            int x0, x1; float lx;
            bil_src(ox, w, x0, x1, lx);
            const float wx = (x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f);
            if (wx == 0.f) continue;
            const T* gp = gv + ((int64_t)oy * W + ox) * C;
            const float ww = wy * wx;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += ww * to_f(gp[i]);
        }
    }
    *reinterpret_cast<float4*>(d_vf + (((int64_t)v * h + sy) * w + sx) * C + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

inline int grid_for(int64_t total, int block = 256) { return (int)((total + block - 1) / block); }

// modweight_*_smem_kernel preconditions: float4-able rows that fit the default 48 KB of dynamic shared memory
inline bool row_fits_smem(int64_t n, const void* a, const void* b) {
    return n % 4 == 0 && n * 4 <= 48 * 1024 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0;
}

}  // namespace agr

using namespace agr;

extern "C" {

int agr_upfirdn2d(int32_t dtype, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t out_h,
                  int32_t out_w, const float* taps, int32_t kh, int32_t kw, int32_t up, int32_t down, int32_t pad_x0,
                  int32_t pad_y0, void* cuda_stream) {
    if (!x || !y || !taps || kh < 1 || kw < 1 || kh * kw > 64 || up < 1 || down < 1 || N < 1 || C < 1) return AGR_ERR_INVALID_ARGUMENT;
    if (out_h < 1 || out_w < 1) return AGR_OK;
    FirParams fp;
    for (int i = 0; i < kh * kw; ++i) fp.taps[i] = taps[i];
    fp.kh = kh; fp.kw = kw; fp.up = up; fp.down = down; fp.pad_x0 = pad_x0; fp.pad_y0 = pad_y0;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
#define AGR_FIR_LAUNCH(T, VEC, VN, U, D, K)                                                                             \
    upfirdn_kernel<T, VEC, U, D, K><<<grid_for((int64_t)N * out_h * out_w * (C / VN)), 256, 0, s>>>(                    \
        (const T*)x, (T*)y, N, H, W, C, out_h, out_w, fp)
#define AGR_FIR_SHAPE(T, VEC, VN)                                                                                       \
    do {                                                                                                                \
        if (kh == 4 && kw == 4 && up == 1 && down == 1) {                                                               \
            /* strips once the strip grid still fills the machine; the per-pixel kernel keeps tiny maps parallel */    \
            const int64_t strip_threads = (int64_t)N * ((out_h + 3) / 4) * out_w * (C / VN);                            \
            if (VEC && strip_threads >= 148 * 1024)                                                                     \
                blur_strip_kernel<T, 4, 4><<<grid_for(strip_threads), 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C, out_h, out_w, fp); \
            else AGR_FIR_LAUNCH(T, VEC, VN, 1, 1, 4);                                                                   \
        }                                                                                                               \
        else if (kh == 4 && kw == 4 && up == 2 && down == 1) AGR_FIR_LAUNCH(T, VEC, VN, 2, 1, 4);                       \
        else if (kh == 4 && kw == 4 && up == 1 && down == 2) AGR_FIR_LAUNCH(T, VEC, VN, 1, 2, 4);                       \
        else AGR_FIR_LAUNCH(T, VEC, VN, 0, 0, 0);                                                                       \
    } while (0)
    if (dtype == AGR_BF16) {
        if (C % 8 == 0) AGR_FIR_SHAPE(__nv_bfloat16, true, 8); else AGR_FIR_SHAPE(__nv_bfloat16, false, 1);
    } else {
        if (C % 4 == 0) AGR_FIR_SHAPE(float, true, 4); else AGR_FIR_SHAPE(float, false, 1);
    }
#undef AGR_FIR_SHAPE
#undef AGR_FIR_LAUNCH
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_haar(int32_t dtype, int32_t mode, const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* cuda_stream) {
    if (!x || !y || mode < 0 || mode > 3 || N < 1) return AGR_ERR_INVALID_ARGUMENT;
    const bool analysis = (mode == 0 || mode == 3);
    if (analysis && ((H & 1) || (W & 1))) return AGR_ERR_INVALID_ARGUMENT;
    if (!analysis && (C % 4)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int Ci = analysis ? C : C / 4;
    const int Hs = analysis ? H / 2 : H, Ws = analysis ? W / 2 : W;
#define AGR_HAAR(T, VN)                                                                                                 \
    do {                                                                                                                \
        if (Ci % VN == 0) {                                                                                             \
            const int g = grid_for((int64_t)N * Hs * Ws * (Ci / VN));                                                   \
            if (analysis) haar_kernel<T, true, true><<<g, 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C);                 \
            else haar_kernel<T, true, false><<<g, 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C);                         \
        } else {                                                                                                        \
            const int g = grid_for((int64_t)N * Hs * Ws * Ci);                                                          \
            if (analysis) haar_kernel<T, false, true><<<g, 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C);                \
            else haar_kernel<T, false, false><<<g, 256, 0, s>>>((const T*)x, (T*)y, N, H, W, C);                        \
        }                                                                                                               \
    } while (0)
    if (dtype == AGR_BF16) AGR_HAAR(__nv_bfloat16, 8); else AGR_HAAR(float, 4);
#undef AGR_HAAR
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_wavelet_upsample(int32_t dtype, int32_t adjoint, const void* x, void* y, int32_t N, int32_t h, int32_t w, int32_t Ci,
                         const float* taps256, void* cuda_stream) {
    if (!x || !y || !taps256 || N < 1 || h < 1 || w < 1 || Ci < 1) return AGR_ERR_INVALID_ARGUMENT;
    WaveletTaps kp;
    for (int i = 0; i < 256; ++i) kp.k[i] = taps256[i];
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int g = grid_for((int64_t)N * h * w * Ci);
    if (dtype == AGR_BF16) {
        using T = __nv_bfloat16;
        if (adjoint) wavelet_up_bwd_kernel<T><<<g, 256, 0, s>>>((const T*)x, (T*)y, N, h, w, Ci, kp);
        else wavelet_up_fwd_kernel<T><<<g, 256, 0, s>>>((const T*)x, (T*)y, N, h, w, Ci, kp);
    } else {
        if (adjoint) wavelet_up_bwd_kernel<float><<<g, 256, 0, s>>>((const float*)x, (float*)y, N, h, w, Ci, kp);
        else wavelet_up_fwd_kernel<float><<<g, 256, 0, s>>>((const float*)x, (float*)y, N, h, w, Ci, kp);
    }
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_fused_bias_act(int32_t dtype, const void* x, const void* bias, const void* refer, void* out, int64_t size_x, int64_t step_b,
                       int64_t size_b, int32_t act, int32_t grad, float alpha, float scale, void* cuda_stream) {
    if (!x || !out || size_x < 0 || (bias && (step_b < 1 || size_b < 1)) || (act != 1 && act != 3) || grad < 0 || grad > 2) return AGR_ERR_INVALID_ARGUMENT;
    if (size_x == 0) return AGR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    int64_t g = (size_x + 255) / 256; if (g > 148 * 16) g = 148 * 16;
    if (dtype == AGR_BF16) {
        using T = __nv_bfloat16;
        fused_bias_act_ref_kernel<T><<<(unsigned)g, 256, 0, s>>>((T*)out, (const T*)x, (const T*)bias, (const T*)refer, act, grad, alpha, scale, size_x, step_b, size_b);
    } else if (dtype == AGR_F32) {
        fused_bias_act_ref_kernel<float><<<(unsigned)g, 256, 0, s>>>((float*)out, (const float*)x, (const float*)bias, (const float*)refer, act, grad, alpha, scale, size_x, step_b, size_b);
    } else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_bias_act_forward(int32_t dtype, const void* x, void* y, int64_t pixels, int32_t C, const float* bias, const float* noise,
                         const float* noise_w, int64_t noise_period, int32_t activate, void* cuda_stream) {
    if (!x || !y || pixels < 0 || C < 1 || (noise && noise_period < 1)) return AGR_ERR_INVALID_ARGUMENT;
    if (!noise) noise_period = 1;
    if (pixels == 0) return AGR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int VN = dtype == AGR_BF16 ? 8 : 4;
    const bool vec = (C % VN) == 0;
    const int64_t total = pixels * (vec ? C / VN : C);
    int g = grid_for(total); if (g > 148 * 32) g = 148 * 32;
    if (dtype == AGR_BF16) {
        using T = __nv_bfloat16;
        if (vec) bias_act_fwd_kernel<T, true><<<g, 256, 0, s>>>((const T*)x, (T*)y, pixels, C, bias, noise, noise_w, noise_period, activate);
        else bias_act_fwd_kernel<T, false><<<g, 256, 0, s>>>((const T*)x, (T*)y, pixels, C, bias, noise, noise_w, noise_period, activate);
    } else {
        if (vec) bias_act_fwd_kernel<float, true><<<g, 256, 0, s>>>((const float*)x, (float*)y, pixels, C, bias, noise, noise_w, noise_period, activate);
        else bias_act_fwd_kernel<float, false><<<g, 256, 0, s>>>((const float*)x, (float*)y, pixels, C, bias, noise, noise_w, noise_period, activate);
    }
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_bias_act_backward(int32_t dtype, const void* dy, const void* y, void* dx, int64_t pixels, int32_t C, const float* noise,
                          int64_t noise_period, float* d_bias, float* d_noise_w, int32_t activate, void* cuda_stream) {
    if (!dy || (!dx && activate) || (activate && !y) || pixels < 0 || C < 1 || (noise && noise_period < 1)) return AGR_ERR_INVALID_ARGUMENT;
    if (!noise) noise_period = 1;
    if (pixels == 0) return AGR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int VN = dtype == AGR_BF16 ? 8 : 4;
    const bool vec = (C % VN) == 0;
    const int cv = vec ? C / VN : C;
    if (cv > 256) return AGR_ERR_INVALID_ARGUMENT;  // up to 2048 bf16 / 1024 fp32 channels
    const int lanes = 256 / cv;
    // serial pixels per lane: 16 on large maps (amortises the C atomics a block ends with), fewer when that would
    // leave most SMs idle -- the small maps of the coarse levels were latency-bound at 16 dependent iterations
    int64_t iters = pixels / ((int64_t)lanes * 148 * 2);
    iters = iters < 1 ? 1 : (iters > 16 ? 16 : iters);
    int ppb = lanes * (int)iters;  // pixels per block
    int64_t blocks = (pixels + ppb - 1) / ppb;
    if (blocks > 148 * 16) { blocks = 148 * 16; ppb = (int)((pixels + blocks - 1) / blocks); }
    const size_t smem = (size_t)(C + 1) * sizeof(float);
    if (dtype == AGR_BF16) {
        using T = __nv_bfloat16;
        if (vec) bias_act_bwd_kernel<T, true><<<(unsigned)blocks, 256, smem, s>>>((const T*)dy, (const T*)y, (T*)dx, pixels, C, noise, noise_period, d_bias, d_noise_w, activate, ppb);
        else bias_act_bwd_kernel<T, false><<<(unsigned)blocks, 256, smem, s>>>((const T*)dy, (const T*)y, (T*)dx, pixels, C, noise, noise_period, d_bias, d_noise_w, activate, ppb);
    } else {
        if (vec) bias_act_bwd_kernel<float, true><<<(unsigned)blocks, 256, smem, s>>>((const float*)dy, (const float*)y, (float*)dx, pixels, C, noise, noise_period, d_bias, d_noise_w, activate, ppb);
        else bias_act_bwd_kernel<float, false><<<(unsigned)blocks, 256, smem, s>>>((const float*)dy, (const float*)y, (float*)dx, pixels, C, noise, noise_period, d_bias, d_noise_w, activate, ppb);
    }
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_bilinear2x_add_forward(int32_t dtype, const float* vf, const void* base, void* y, int32_t V, int32_t Vb, int32_t h, int32_t w,
                               int32_t C, void* cuda_stream) {
    if (!vf || !base || !y || V < 1 || (Vb != 1 && Vb != V) || h < 1 || w < 1 || C < 4 || (C % 4)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int g = grid_for((int64_t)V * 4 * h * w * (C / 4));
    if (dtype == AGR_BF16) bilinear2x_add_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(vf, (const __nv_bfloat16*)base, (__nv_bfloat16*)y, V, Vb, h, w, C);
    else bilinear2x_add_kernel<float><<<g, 256, 0, s>>>(vf, (const float*)base, (float*)y, V, Vb, h, w, C);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_bilinear2x_backward(int32_t dtype, const void* g, float* d_vf, int32_t V, int32_t h, int32_t w, int32_t C, void* cuda_stream) {
    if (!g || !d_vf || V < 1 || h < 1 || w < 1 || C < 4 || (C % 4)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int gr = grid_for((int64_t)V * h * w * (C / 4));
    if (dtype == AGR_BF16) bilinear2x_bwd_kernel<__nv_bfloat16><<<gr, 256, 0, s>>>((const __nv_bfloat16*)g, d_vf, V, h, w, C);
    else bilinear2x_bwd_kernel<float><<<gr, 256, 0, s>>>((const float*)g, d_vf, V, h, w, C);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_sum_batch(int32_t dtype, const void* x, void* y, int32_t V, int64_t n, void* cuda_stream) {
    if (!x || !y || V < 1 || n < 0) return AGR_ERR_INVALID_ARGUMENT;
    const int VN = dtype == AGR_BF16 ? 8 : 4;
    if (n % VN) return AGR_ERR_INVALID_ARGUMENT;
    if (n == 0) return AGR_OK;
    const int64_t nvec = n / VN;
    int g = grid_for(nvec); if (g > 148 * 16) g = 148 * 16;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (dtype == AGR_BF16) sum_batch_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, V, nvec);
    else sum_batch_kernel<float><<<g, 256, 0, s>>>((const float*)x, (float*)y, V, nvec);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_equal_linear_forward(const float* w, const float* bias, const float* x, float scale, float lr_mul, int32_t out_dim,
                             int32_t in_dim, float* y, void* cuda_stream) {
    if (!w || !x || !y || out_dim < 1 || in_dim < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    equal_linear_fwd_kernel<<<grid_for((int64_t)out_dim * 32), 256, 0, st>>>(w, bias, x, scale, lr_mul, out_dim, in_dim, y);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_equal_linear_backward(const float* w, const float* x, const float* dy, float scale, float lr_mul, int32_t out_dim,
                              int32_t in_dim, float* d_w, float* d_bias, float* d_x, void* cuda_stream) {
    if (!w || !x || !dy || !d_w || out_dim < 1 || in_dim < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    equal_linear_bwd_kernel<<<(out_dim + kLinRows - 1) / kLinRows, 256, 0, st>>>(w, x, dy, scale, lr_mul, out_dim, in_dim, d_w, d_bias, d_x);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

static int equal_linear_group(const AgrEqualLinearItem* items, int32_t count, bool backward, cudaStream_t st) {
    if (!items || count < 0) return AGR_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < count; ++i) {
        const AgrEqualLinearItem& it = items[i];
        if (!it.w || !it.x || it.out_dim < 1 || it.in_dim < 1) return AGR_ERR_INVALID_ARGUMENT;
        if (!backward && !it.y) return AGR_ERR_INVALID_ARGUMENT;
        if (backward && (!it.dy || !it.d_w)) return AGR_ERR_INVALID_ARGUMENT;
    }
    for (int base = 0; base < count; base += kLinGroup) {
        LinGroup g;
        g.count = count - base < kLinGroup ? count - base : kLinGroup;
        int blocks = 0;
        for (int i = 0; i < g.count; ++i) {
            const AgrEqualLinearItem& it = items[base + i];
            LinEntry& e = g.e[i];
            e.w = it.w; e.bias = it.bias; e.x = it.x; e.dy = it.dy; e.y = it.y; e.d_w = it.d_w; e.d_bias = it.d_bias; e.d_x = it.d_x;
            e.scale = it.scale; e.lr_mul = it.lr_mul; e.out_dim = it.out_dim; e.in_dim = it.in_dim; e.block_begin = blocks; e.pad_ = 0;
            blocks += backward ? (it.out_dim + kLinRows - 1) / kLinRows : (it.out_dim + 7) / 8;
        }
        if (blocks == 0) continue;
        if (backward) equal_linear_group_kernel<true><<<blocks, 256, 0, st>>>(g);
        else equal_linear_group_kernel<false><<<blocks, 256, 0, st>>>(g);
        if (cudaGetLastError() != cudaSuccess) return AGR_ERR_CUDA;
    }
    return AGR_OK;
}

int agr_equal_linear_group_forward(const AgrEqualLinearItem* items, int32_t count, void* cuda_stream) {
    return equal_linear_group(items, count, false, static_cast<cudaStream_t>(cuda_stream));
}

int agr_equal_linear_group_backward(const AgrEqualLinearItem* items, int32_t count, void* cuda_stream) {
    return equal_linear_group(items, count, true, static_cast<cudaStream_t>(cuda_stream));
}

int agr_modweight_forward(int32_t dtype, const float* w, const float* s, float scale, int32_t Cout, int32_t Cin, int32_t k,
                          int32_t demodulate, int32_t transpose_io, void* w_out, float* demod_out, void* cuda_stream) {
    if (!w || !s || !w_out || Cout < 1 || Cin < 1 || k < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t n = (int64_t)Cin * k * k;
    const bool fits = row_fits_smem(n, w, w);
    const size_t smem = fits ? (size_t)n * sizeof(float) : 0;
#define AGR_MW(T, SM) modweight_fwd_kernel<T, SM><<<Cout, 256, smem, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (T*)w_out, demod_out)
    if (dtype == AGR_BF16) { if (fits) AGR_MW(__nv_bfloat16, true); else AGR_MW(__nv_bfloat16, false); }
    else { if (fits) AGR_MW(float, true); else AGR_MW(float, false); }
#undef AGR_MW
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_modweight_backward(int32_t dtype, const float* w, const float* s, float scale, int32_t Cout, int32_t Cin, int32_t k,
                           int32_t demodulate, int32_t transpose_io, const void* d_wout, const float* demod, float* d_w,
                           float* d_s, void* cuda_stream) {
    if (!w || !s || !d_wout || !d_w || (demodulate && !demod)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t n = (int64_t)Cin * k * k;
    const bool fits = row_fits_smem(n, w, d_w);
    const size_t smem = fits ? (size_t)n * sizeof(float) : 0;
#define AGR_MW(T, SM) modweight_bwd_kernel<T, SM><<<Cout, 256, smem, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (const T*)d_wout, demod, d_w, d_s)
    if (dtype == AGR_BF16) { if (fits) AGR_MW(__nv_bfloat16, true); else AGR_MW(__nv_bfloat16, false); }
    else { if (fits) AGR_MW(float, true); else AGR_MW(float, false); }
#undef AGR_MW
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

static int modweight_group(int32_t dtype, const AgrModWeightItem* items, int32_t count, bool backward, cudaStream_t st) {
    if (!items || count < 0) return AGR_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < count; ++i) {
        const AgrModWeightItem& it = items[i];
        if (!it.w || !it.s || it.Cout < 1 || it.Cin < 1 || it.k < 1) return AGR_ERR_INVALID_ARGUMENT;
        if (!backward && !it.w_out) return AGR_ERR_INVALID_ARGUMENT;
        if (backward && (!it.d_wout || !it.d_w || (it.demodulate && !it.demod))) return AGR_ERR_INVALID_ARGUMENT;
    }
    for (int base = 0; base < count; base += kModGroup) {
        ModGroup g;
        g.count = count - base < kModGroup ? count - base : kModGroup;
        int blocks = 0;
        size_t smem = 0;
        for (int i = 0; i < g.count; ++i) {
            const AgrModWeightItem& it = items[base + i];
            ModEntry& e = g.e[i];
            const int64_t n = (int64_t)it.Cin * it.k * it.k;
            const bool fits = row_fits_smem(n, it.w, backward ? (const void*)it.d_w : (const void*)it.w);
            e.w = it.w; e.s = it.s; e.demod_in = it.demod; e.demod_out = it.demod; e.d_wout = it.d_wout; e.w_out = it.w_out;
            e.d_w = it.d_w; e.d_s = it.d_s; e.scale = it.scale; e.Cout = it.Cout; e.Cin = it.Cin; e.kk = it.k * it.k;
            e.flags = (it.demodulate ? 1 : 0) | (it.transpose_io ? 2 : 0) | (fits ? 4 : 0);
            e.block_begin = blocks;
            blocks += it.Cout;
            if (fits && (size_t)n * sizeof(float) > smem) smem = (size_t)n * sizeof(float);
        }
        if (blocks == 0) continue;
        if (dtype == AGR_BF16) {
            if (backward) modweight_group_kernel<__nv_bfloat16, true><<<blocks, 256, smem, st>>>(g);
            else modweight_group_kernel<__nv_bfloat16, false><<<blocks, 256, smem, st>>>(g);
        } else {
            if (backward) modweight_group_kernel<float, true><<<blocks, 256, smem, st>>>(g);
            else modweight_group_kernel<float, false><<<blocks, 256, smem, st>>>(g);
        }
        if (cudaGetLastError() != cudaSuccess) return AGR_ERR_CUDA;
    }
    return AGR_OK;
}

int agr_modweight_group_forward(int32_t dtype, const AgrModWeightItem* items, int32_t count, void* cuda_stream) {
    return modweight_group(dtype, items, count, false, static_cast<cudaStream_t>(cuda_stream));
}

int agr_modweight_group_backward(int32_t dtype, const AgrModWeightItem* items, int32_t count, void* cuda_stream) {
    return modweight_group(dtype, items, count, true, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
