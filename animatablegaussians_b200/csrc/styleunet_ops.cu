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
// vector loads + index arithmetic per output and is issue-bound (r01 ncu: ~1.1 TB/s at 16x1024x1024x16).  Here a
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
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VN;
        const int64_t p = idx / cv;
        const float add = noise ? nw * noise[p % nper] : 0.f;
        float f[VN];
        if (VECTOR) unpack(ld_vec<typename VecOf<T>::V>(x + p * C + c), f); else f[0] = to_f(x[p * C + c]);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            float v = f[i] + add + (bias ? bias[c + i] : 0.f);
            if (activate) v = (v > 0.f ? v : v * 0.2f) * kSqrt2;
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
                if (activate) g[i] *= (o[i] > 0.f ? kSqrt2 : 0.2f * kSqrt2);
                bsum[i] += g[i];
                psum += g[i];
            }
            if (noise) nsum += psum * noise[p % nper];
            if (VECTOR) { typename VecOf<T>::V v; pack(g, v); *reinterpret_cast<typename VecOf<T>::V*>(dx + p * C + my_c) = v; }
            else dx[p * C + my_c] = from_f<T>(g[0]);
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

// ------------------------------------------------------------------------------------------------ modweight
// One block per output channel. w: (Cout, Cin, k, k) fp32.
template <typename T>
__global__ void __launch_bounds__(256) modweight_fwd_kernel(const float* __restrict__ w, const float* __restrict__ s, float scale,
                                                           int Cout, int Cin, int kk, int demodulate, int transpose_io,
                                                           T* __restrict__ w_out, float* __restrict__ demod_out) {
    __shared__ float s_part[32];
    __shared__ float s_demod;
    const int co = blockIdx.x;
    const int n = Cin * kk;
    const float* wr = w + (size_t)co * n;
    float d = 1.f;
    if (demodulate) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float u = scale * wr[i] * s[i / kk];
            acc += u * u;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) { s_demod = rsqrtf(v + 1e-8f); if (demod_out) demod_out[co] = s_demod; }
        }
        __syncthreads();
        d = s_demod;
    }
    // output index: KRSC [co][t][ci]  or (transpose_io) [ci][t][co]
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int t = i / Cin, ci = i - t * Cin;  // iterate with ci fastest for coalesced KRSC writes
        const float u = scale * wr[ci * kk + t] * s[ci] * d;
        const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
        w_out[o] = from_f<T>(u);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) modweight_bwd_kernel(const float* __restrict__ w, const float* __restrict__ s, float scale,
                                                           int Cout, int Cin, int kk, int demodulate, int transpose_io,
                                                           const T* __restrict__ d_wout, const float* __restrict__ demod,
                                                           float* __restrict__ d_w, float* __restrict__ d_s) {
    __shared__ float s_part[32];
    __shared__ float s_dot;
    const int co = blockIdx.x;
    const int n = Cin * kk;
    const float* wr = w + (size_t)co * n;
    const float d = demodulate ? demod[co] : 1.f;
    float dot = 0.f;  // sum_i dW'_i * u_i
    if (demodulate) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int t = i / Cin, ci = i - t * Cin;
            const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
            dot += to_f(d_wout[o]) * (scale * wr[ci * kk + t] * s[ci]);
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) s_dot = v;
        }
        __syncthreads();
        dot = s_dot;
    }
    const float d3dot = demodulate ? d * d * d * dot : 0.f;
    // thread owns input channels ci = threadIdx.x, +blockDim.x, ... so d_s needs one atomic per (co, ci)
    for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
        float ds = 0.f;
        for (int t = 0; t < kk; ++t) {
            const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
            const float wv = wr[ci * kk + t];
            const float u = scale * wv * s[ci];
            const float du = d * to_f(d_wout[o]) - u * d3dot;   // dL/du
            d_w[(size_t)co * n + ci * kk + t] = du * scale * s[ci];
            ds += du * scale * wv;
        }
        atomicAdd(&d_s[ci], ds);
    }
}

// Shared-memory variants (rows of <= 12288 floats, 16-byte aligned): the (Cin,k,k) master-weight row is staged with
// coalesced float4 loads issued back to back (one DRAM round trip instead of n/256 dependent ones), the strided
// [ci*kk+t] accesses then hit shared memory (stride kk is odd or 1: conflict-free), and the backward builds the d_w
// row in place and writes it back coalesced.  Same per-element arithmetic and reduction order as the kernels above.
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

template <typename T>
__global__ void __launch_bounds__(256) modweight_fwd_smem_kernel(const float* __restrict__ w, const float* __restrict__ s, float scale,
                                                                int Cout, int Cin, int kk, int demodulate, int transpose_io,
                                                                T* __restrict__ w_out, float* __restrict__ demod_out) {
    extern __shared__ float4 s_row4[];
    float* sw = reinterpret_cast<float*>(s_row4);
    __shared__ float s_part[32];
    __shared__ float s_demod;
    const int co = blockIdx.x;
    const int n = Cin * kk;
    stage_row(w + (size_t)co * n, sw, n >> 2);
    __syncthreads();
    float d = 1.f;
    if (demodulate) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float u = scale * sw[i] * s[i / kk];
            acc += u * u;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) { s_demod = rsqrtf(v + 1e-8f); if (demod_out) demod_out[co] = s_demod; }
        }
        __syncthreads();
        d = s_demod;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int t = i / Cin, ci = i - t * Cin;
        const float u = scale * sw[ci * kk + t] * s[ci] * d;
        const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
        w_out[o] = from_f<T>(u);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) modweight_bwd_smem_kernel(const float* __restrict__ w, const float* __restrict__ s, float scale,
                                                                int Cout, int Cin, int kk, int demodulate, int transpose_io,
                                                                const T* __restrict__ d_wout, const float* __restrict__ demod,
                                                                float* __restrict__ d_w, float* __restrict__ d_s) {
    extern __shared__ float4 s_row4[];
    float* sw = reinterpret_cast<float*>(s_row4);
    __shared__ float s_part[32];
    __shared__ float s_dot;
    const int co = blockIdx.x;
    const int n = Cin * kk;
    stage_row(w + (size_t)co * n, sw, n >> 2);
    __syncthreads();
    const float d = demodulate ? demod[co] : 1.f;
    float dot = 0.f;
    if (demodulate) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int t = i / Cin, ci = i - t * Cin;
            const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
            dot += to_f(d_wout[o]) * (scale * sw[ci * kk + t] * s[ci]);
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) s_dot = v;
        }
        __syncthreads();   // also orders the dot-phase reads of sw before the in-place overwrite below
        dot = s_dot;
    }
    const float d3dot = demodulate ? d * d * d * dot : 0.f;
    for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
        float ds = 0.f;
        const float sc = s[ci];
        for (int t = 0; t < kk; ++t) {
            const size_t o = transpose_io ? ((size_t)ci * kk + t) * Cout + co : ((size_t)co * kk + t) * Cin + ci;
            const float wv = sw[ci * kk + t];
            const float u = scale * wv * sc;
            const float du = d * to_f(d_wout[o]) - u * d3dot;
            sw[ci * kk + t] = du * scale * sc;   // element owned by this thread only
            ds += du * scale * wv;
        }
        atomicAdd(&d_s[ci], ds);
    }
    __syncthreads();
    float4* out4 = reinterpret_cast<float4*>(d_w + (size_t)co * n);
    for (int j = threadIdx.x; j < (n >> 2); j += blockDim.x) out4[j] = s_row4[j];
}

// ------------------------------------------------------------------------------------------------ EqualLinear (style vector)
// forward: one warp per output row.
__global__ void __launch_bounds__(256) equal_linear_fwd_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                              const float* __restrict__ x, float scale, float lr_mul, int out_dim,
                                                              int in_dim, float* __restrict__ y) {
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
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
template <int ROWS>
__global__ void __launch_bounds__(256) equal_linear_bwd_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                              const float* __restrict__ dy, float scale, float lr_mul, int out_dim,
                                                              int in_dim, float* __restrict__ d_w, float* __restrict__ d_bias,
                                                              float* __restrict__ d_x) {
    const int j0 = blockIdx.x * ROWS;
    float g[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) g[r] = (j0 + r < out_dim) ? dy[j0 + r] : 0.f;
    if (d_bias && threadIdx.x < ROWS && j0 + threadIdx.x < out_dim) d_bias[j0 + threadIdx.x] = lr_mul * dy[j0 + threadIdx.x];
    for (int i = threadIdx.x; i < in_dim; i += blockDim.x) {
        const float xi = x[i];
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (j0 + r < out_dim) {
                const size_t o = (size_t)(j0 + r) * in_dim + i;
                acc += w[o] * g[r];
                d_w[o] = scale * g[r] * xi;
            }
        }
        if (d_x) atomicAdd(&d_x[i], scale * acc);
    }
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
    if (!dy || !dx || (activate && !y) || pixels < 0 || C < 1 || (noise && noise_period < 1)) return AGR_ERR_INVALID_ARGUMENT;
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
    constexpr int ROWS = 4;
    equal_linear_bwd_kernel<ROWS><<<(out_dim + ROWS - 1) / ROWS, 256, 0, st>>>(w, x, dy, scale, lr_mul, out_dim, in_dim, d_w, d_bias, d_x);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_modweight_forward(int32_t dtype, const float* w, const float* s, float scale, int32_t Cout, int32_t Cin, int32_t k,
                          int32_t demodulate, int32_t transpose_io, void* w_out, float* demod_out, void* cuda_stream) {
    if (!w || !s || !w_out || Cout < 1 || Cin < 1 || k < 1) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t n = (int64_t)Cin * k * k;
    if (row_fits_smem(n, w, w)) {
        const size_t smem = (size_t)n * sizeof(float);
        if (dtype == AGR_BF16) modweight_fwd_smem_kernel<__nv_bfloat16><<<Cout, 256, smem, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (__nv_bfloat16*)w_out, demod_out);
        else modweight_fwd_smem_kernel<float><<<Cout, 256, smem, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (float*)w_out, demod_out);
    } else if (dtype == AGR_BF16) modweight_fwd_kernel<__nv_bfloat16><<<Cout, 256, 0, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (__nv_bfloat16*)w_out, demod_out);
    else modweight_fwd_kernel<float><<<Cout, 256, 0, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (float*)w_out, demod_out);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_modweight_backward(int32_t dtype, const float* w, const float* s, float scale, int32_t Cout, int32_t Cin, int32_t k,
                           int32_t demodulate, int32_t transpose_io, const void* d_wout, const float* demod, float* d_w,
                           float* d_s, void* cuda_stream) {
    if (!w || !s || !d_wout || !d_w || !d_s || (demodulate && !demod)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t n = (int64_t)Cin * k * k;
    if (row_fits_smem(n, w, d_w)) {
        const size_t smem = (size_t)n * sizeof(float);
        if (dtype == AGR_BF16) modweight_bwd_smem_kernel<__nv_bfloat16><<<Cout, 256, smem, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (const __nv_bfloat16*)d_wout, demod, d_w, d_s);
        else modweight_bwd_smem_kernel<float><<<Cout, 256, smem, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (const float*)d_wout, demod, d_w, d_s);
    } else if (dtype == AGR_BF16) modweight_bwd_kernel<__nv_bfloat16><<<Cout, 256, 0, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (const __nv_bfloat16*)d_wout, demod, d_w, d_s);
    else modweight_bwd_kernel<float><<<Cout, 256, 0, st>>>(w, s, scale, Cout, Cin, k * k, demodulate, transpose_io, (const float*)d_wout, demod, d_w, d_s);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // extern "C"
