// CUDA-core implicit-GEMM convolutions (path 2 of include/agr_conv.h): fp32 accumulate, fp32 or bf16 tensors, any
// channel count / kernel size <= 4 / stride <= 2, convolution or transposed convolution.
//
// Used where the tcgen05 tiles do not apply: the fp32 parity mode of the StyleUNet, and the narrow layers of the bf16
// product path — 3-channel pose-map inputs (conv_in, FromRGB: dual_styleunet.py:686-701), the 1-channel view map
// (viewdir_net[0], network/avatar.py:46-50) and the 12 / 32-channel ToRGB outputs (dual_styleunet.py:607-633).  Those are
// HBM-bound layers (a few FLOP per byte), so a register-tiled SGEMM-style kernel is the right tool; the reference runs
// them through cuDNN like every other convolution (conv2d_gradfix.py:34,66).
//
//   conv_direct_kernel        y[P][co] = epilogue( sum_{t,ci} x[src(P,t)][ci] * w[co][t][ci] )      tile 64 pixels x BN channels
//   conv_direct_wgrad_kernel  dw[co][t][ci] += sum_q dy[a(q,t)][co] * x[b(q,t)][ci]                 tile BM co x 64 (t,ci), split over pixels
// K is walked in chunks of 16: within one tap when Cin % 16 == 0 (vector loads along the channel axis), over the flattened
// (tap, channel) index otherwise (scalar loads; 3- and 1-channel inputs).
#include "conv_common.cuh"

namespace agr {
namespace direct {

struct DirectParams {
    int N, H, W, Cin, OH, OW, Cout;
    int k, stride, pad, transposed;
    const void* x; const void* w; void* y;
    const float* bias; const float* noise; const float* noise_w;
    int activate;
    int w_cin_total, w_cin_offset;
    // weight gradient only
    const void* dy; float* dw; int ci_total, ci_offset, slices;
};

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void ld4(const float* p, float* f) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float* f) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// source pixel of output pixel (oy, ox) for tap (ky, kx); false = reads zero
__device__ __forceinline__ bool src_pixel(const DirectParams& p, int oy, int ox, int ky, int kx, int* iy, int* ix) {
    if (!p.transposed) {
        *iy = oy * p.stride + ky - p.pad; *ix = ox * p.stride + kx - p.pad;
    } else {
        const int ty = oy + p.pad - ky, tx = ox + p.pad - kx;
        if (ty < 0 || tx < 0 || (ty % p.stride) || (tx % p.stride)) return false;
        *iy = ty / p.stride; *ix = tx / p.stride;
    }
    return *iy >= 0 && *iy < p.H && *ix >= 0 && *ix < p.W;
}

constexpr int KC = 16, PT = 64;   // K chunk, pixel tile

template <typename TI, typename TO, int BN, bool VEC>
__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectParams p) {
    __shared__ float As[KC][PT + 4];
    __shared__ float Bs[KC][BN + 4];
    constexpr int CN = BN / 16;                       // output channels per thread
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long total = (long)p.N * p.OH * p.OW;
    const long pix0 = (long)blockIdx.x * PT;
    const int co0 = blockIdx.y * BN;
    const int taps = p.k * p.k;
    const TI* x = static_cast<const TI*>(p.x);
    const TI* w = static_cast<const TI*>(p.w);
    const int wrow = taps * p.w_cin_total;            // elements per output channel of the weight

    float acc[4][CN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;

    if (VEC) {
        // thread -> (pixel tid/4, 4 contiguous channels) of the A tile; (co = tid/4 (+64), 4 channels) of the B tile
        const int apx = tid >> 2, ac = (tid & 3) * 4;
        const long P = pix0 + apx;
        const bool pv = P < total;
        int n = 0, oy = 0, ox = 0;
        if (pv) { n = (int)(P / ((long)p.OH * p.OW)); const int r = (int)(P - (long)n * p.OH * p.OW); oy = r / p.OW; ox = r - oy * p.OW; }
        for (int t = 0; t < taps; ++t) {
            int iy, ix;
            const bool v = pv && src_pixel(p, oy, ox, t / p.k, t % p.k, &iy, &ix);
            const TI* xs = v ? x + (((long)n * p.H + iy) * p.W + ix) * p.Cin + ac : nullptr;
            for (int c0 = 0; c0 < p.Cin; c0 += KC) {
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                if (v) ld4(xs + c0, a);
#pragma unroll
                for (int i = 0; i < 4; ++i) As[ac + i][apx] = a[i];
                for (int e = tid; e < BN * 4; e += 256) {
                    const int co = e >> 2, bc = (e & 3) * 4;
                    float b[4] = {0.f, 0.f, 0.f, 0.f};
                    if (co0 + co < p.Cout) ld4(w + (long)(co0 + co) * wrow + (long)t * p.w_cin_total + p.w_cin_offset + c0 + bc, b);
#pragma unroll
                    for (int i = 0; i < 4; ++i) Bs[bc + i][co] = b[i];
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                    float bv[CN];
#pragma unroll
                    for (int j = 0; j < CN; ++j) bv[j] = Bs[kk][tx * CN + j];
                    const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(aa[i], bv[j], acc[i][j]);
                }
                __syncthreads();
            }
        }
    } else {
        // flattened K = taps * Cin; thread -> (pixel tid % 64, k = tid / 64 + 4 i)
        const int apx = tid & 63, ak = tid >> 6;
        const long P = pix0 + apx;
        const bool pv = P < total;
        int n = 0, oy = 0, ox = 0;
        if (pv) { n = (int)(P / ((long)p.OH * p.OW)); const int r = (int)(P - (long)n * p.OH * p.OW); oy = r / p.OW; ox = r - oy * p.OW; }
        const int K = taps * p.Cin;
        for (int k0 = 0; k0 < K; k0 += KC) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = k0 + ak + 4 * i;
                float a = 0.f;
                if (pv && kk < K) {
                    const int t = kk / p.Cin, ci = kk - t * p.Cin;
                    int iy, ix;
                    if (src_pixel(p, oy, ox, t / p.k, t % p.k, &iy, &ix)) a = ldf(x + (((long)n * p.H + iy) * p.W + ix) * p.Cin + ci);
                }
                As[ak + 4 * i][apx] = a;
            }
            for (int e = tid; e < BN * KC; e += 256) {
                const int co = e / KC, kk = k0 + (e % KC);
                float b = 0.f;
                if (co0 + co < p.Cout && kk < K) {
                    const int t = kk / p.Cin, ci = kk - t * p.Cin;
                    b = ldf(w + (long)(co0 + co) * wrow + (long)t * p.w_cin_total + p.w_cin_offset + ci);
                }
                Bs[e % KC][co] = b;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                float bv[CN];
#pragma unroll
                for (int j = 0; j < CN; ++j) bv[j] = Bs[kk][tx * CN + j];
                const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(aa[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // epilogue
    TO* y = static_cast<TO*>(p.y);
    const float nw = (p.noise && p.noise_w) ? p.noise_w[0] : 0.f;
    const float gain = p.activate == 1 ? 1.4142135623730951f : 1.f, neg_slope = p.activate == 3 ? 0.f : 0.2f;
    const long plane = (long)p.OH * p.OW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long P = pix0 + ty * 4 + i;
        if (P >= total) continue;
        const float add = p.noise ? nw * p.noise[P % plane] : 0.f;
#pragma unroll
        for (int j = 0; j < CN; ++j) {
            const int co = co0 + tx * CN + j;
            if (co >= p.Cout) continue;
            float v = acc[i][j] + add + (p.bias ? p.bias[co] : 0.f);
            if (p.activate) v = (v > 0.f ? v : neg_slope * v) * gain;
            stf(y + P * p.Cout + co, v);
        }
    }
}

// dw[co][t][ci] += sum_q dy[a(q,t)][co] * x[b(q,t)][ci]; q runs over the plain operand's pixels:
//   convolution:             a = q (output pixel),  b = q*stride + k - pad
//   transposed convolution:  b = q (input pixel),   a = q*stride + k - pad    (needs the 64-wide (t,ci) tile inside one tap)
template <typename T, int BMc, int NT>
__global__ void __launch_bounds__(256) conv_direct_wgrad_kernel(const DirectParams p) {
    __shared__ float As[KC][BMc + 4];     // [pixel][co]
    __shared__ float Bs[KC][NT + 4];      // [pixel][(t,ci)]
    constexpr int TXN = NT / 4, TYM = 256 / TXN, CM = BMc / TYM;   // thread grid: TXN x 4 columns, TYM x CM rows
    static_assert(CM >= 1, "tile too small for 256 threads");
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const int taps = p.k * p.k, NN = taps * p.Cin;
    const int n0 = blockIdx.x * NT, co0 = blockIdx.y * BMc;
    const int GH = p.transposed ? p.H : p.OH, GW = p.transposed ? p.W : p.OW;
    const long total = (long)p.N * GH * GW;
    const long q_begin = (blockIdx.z * total) / p.slices, q_end = ((blockIdx.z + 1) * total) / p.slices;
    const T* x = static_cast<const T*>(p.x);
    const T* dy = static_cast<const T*>(p.dy);

    // this thread's B column (fixed): n = n0 + tid % NT -> (tap, ci); A column: co = co0 + tid % BMc
    const int bn = n0 + (tid % NT);
    const bool bvalid = bn < NN;
    const int bt = bvalid ? bn / p.Cin : 0, bci = bvalid ? bn - bt * p.Cin : 0;
    const int at = n0 / p.Cin;                      // tap of the whole tile (transposed mode: tile inside one tap)
    const int aco = co0 + (tid % BMc);
    const int bk = tid / NT;                        // B pixel row: bk + (256 / NT) i
    const int ak = tid / BMc;                       // A pixel row: ak + (256 / BMc) i

    float acc[CM][4];
#pragma unroll
    for (int i = 0; i < CM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long q0 = q_begin; q0 < q_end; q0 += KC) {
#pragma unroll
        for (int i = 0; i < KC * NT / 256; ++i) {
            const int kk = bk + (256 / NT) * i;
            const long q = q0 + kk;
            float b = 0.f;
            if (bvalid && q < q_end) {
                const int n = (int)(q / ((long)GH * GW)); const int r = (int)(q - (long)n * GH * GW); const int gy = r / GW, gx = r - gy * GW;
                if (p.transposed) b = ldf(x + ((long)q) * p.Cin + bci);
                else {
                    const int iy = gy * p.stride + bt / p.k - p.pad, ix = gx * p.stride + bt % p.k - p.pad;
                    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) b = ldf(x + (((long)n * p.H + iy) * p.W + ix) * p.Cin + bci);
                }
            }
            Bs[kk][tid % NT] = b;
        }
#pragma unroll
        for (int i = 0; i < KC * BMc / 256; ++i) {
            const int kk = ak + (256 / BMc) * i;
            const long q = q0 + kk;
            float a = 0.f;
            if (aco < p.Cout && q < q_end) {
                if (!p.transposed) a = ldf(dy + q * p.Cout + aco);
                else {
                    const int n = (int)(q / ((long)GH * GW)); const int r = (int)(q - (long)n * GH * GW); const int gy = r / GW, gx = r - gy * GW;
                    const int oy = gy * p.stride + at / p.k - p.pad, ox = gx * p.stride + at % p.k - p.pad;
                    if (oy >= 0 && oy < p.OH && ox >= 0 && ox < p.OW) a = ldf(dy + (((long)n * p.OH + oy) * p.OW + ox) * p.Cout + aco);
                }
            }
            As[kk][tid % BMc] = a;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < CM; ++i) {
                const float a = As[kk][ty * CM + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, bb[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < CM; ++i) {
        const int co = co0 + ty * CM + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= NN) continue;
            const int t = n / p.Cin, ci = n - t * p.Cin;
            atomicAdd(p.dw + ((long)co * taps + t) * p.ci_total + p.ci_offset + ci, acc[i][j]);
        }
    }
}

// ---- 1x1 convolutions with 12 output channels (ToRGB of the position / colour nets, dual_styleunet.py:607-633) -----------
// At the top decoder level these stream 16 views x 512^2 x 64 channels for 12 outputs: pure HBM streaming (6 FLOP / byte).
// A quarter warp owns one pixel: each lane loads 8 consecutive channels (16 B, fully coalesced across the warp), the
// weights sit in shared memory as fp32, and the 12 partial sums are combined across the 8 lanes with a transposed
// (halving) butterfly — 11 shuffles instead of 36.
constexpr int PW = 12;

template <typename T> __device__ __forceinline__ void ld8(const T* p, float* f);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float* f) { ld4(p, f); ld4(p + 4, f + 4); }
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void st8(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* f) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
}
// the 12 values of one pixel, 4-byte aligned groups: dy[p][0..12)
__device__ __forceinline__ void ld12(const float* p, float* f) { ld4(p, f); ld4(p + 4, f + 4); ld4(p + 8, f + 8); }
__device__ __forceinline__ void ld12(const __nv_bfloat16* p, float* f) { ld4(p, f); ld4(p + 4, f + 4); ld4(p + 8, f + 8); }

struct PwParams {
    const void* x; const void* w; void* y; const void* dy; float* dw;
    long pixels, plane;
    int Cin, w_cin_total, w_cin_offset, ci_total, ci_offset;
    const float* bias; const float* noise; const float* noise_w;
    int activate;
};

template <typename T>
__device__ __forceinline__ void pw_stage_weights(const PwParams& p, float* ws, bool transposed_operand) {
    // ws[co][ci] fp32.  forward / wgrad operand: w[co][w_cin_total] rows; dgrad operand w_t[ci][co]
    const T* w = static_cast<const T*>(p.w);
    for (int e = threadIdx.x; e < PW * p.Cin; e += blockDim.x) {
        const int co = e / p.Cin, ci = e - co * p.Cin;
        ws[e] = transposed_operand ? ldf(w + (long)ci * PW + co) : ldf(w + (long)co * p.w_cin_total + p.w_cin_offset + ci);
    }
    __syncthreads();
}

// REGW: Cin == 64 (the 512^2 level, 16 views: the one that matters) keeps this lane's 12 x 8 weights in registers — the
// shared-memory form spends 24 LDS.128 (4 cycles each on the SM's one shared-memory pipe) per 4 pixels and caps at ~1.5 TB/s.
template <typename T, bool REGW>
__global__ void __launch_bounds__(256) pw12_fwd_kernel(const PwParams p) {
    extern __shared__ float ws[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, c = lane & 7;
    float wr[REGW ? PW : 1][8];
    if (REGW) {
        const T* w = static_cast<const T*>(p.w);
#pragma unroll
        for (int co = 0; co < (REGW ? PW : 1); ++co) ld8<T>(w + (long)co * p.w_cin_total + p.w_cin_offset + c * 8, wr[co]);
    } else {
        pw_stage_weights<T>(p, ws, false);
    }
    const T* x = static_cast<const T*>(p.x);
    T* y = static_cast<T*>(p.y);
    const float nw = (p.noise && p.noise_w) ? p.noise_w[0] : 0.f;
    const float gain = p.activate == 1 ? 1.4142135623730951f : 1.f, neg_slope = p.activate == 3 ? 0.f : 0.2f;
    // the output channel this lane ends up holding after the halving butterfly (+ co_b = its pair's third value)
    const int co_a = 6 * ((c >> 2) & 1) + 3 * ((c >> 1) & 1) + (c & 1), co_b = 6 * ((c >> 2) & 1) + 3 * ((c >> 1) & 1) + 2;
    for (long p4 = ((long)blockIdx.x * 8 + warp) * 4; p4 < p.pixels; p4 += (long)gridDim.x * 32) {
        const long px = p4 + g;
        const bool valid = px < p.pixels;
        float acc[PW];
#pragma unroll
        for (int i = 0; i < PW; ++i) acc[i] = 0.f;
        if (REGW) {
            float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) ld8<T>(x + px * 64 + c * 8, xv);
#pragma unroll
            for (int co = 0; co < PW; ++co)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[co] = fmaf(xv[j], wr[REGW ? co : 0][j], acc[co]);
        } else {
            for (int c0 = 0; c0 < p.Cin; c0 += 64) {
                float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) ld8<T>(x + px * p.Cin + c0 + c * 8, xv);
#pragma unroll
                for (int co = 0; co < PW; ++co) {
                    const float4 w0 = *reinterpret_cast<const float4*>(ws + co * p.Cin + c0 + c * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(ws + co * p.Cin + c0 + c * 8 + 4);
                    acc[co] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y + xv[6] * w1.z + xv[7] * w1.w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float keep = (c & 4) ? acc[i + 6] : acc[i], send = (c & 4) ? acc[i] : acc[i + 6];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float keep = (c & 2) ? acc[i + 3] : acc[i], send = (c & 2) ? acc[i] : acc[i + 3];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        const float keep = (c & 1) ? acc[1] : acc[0], send = (c & 1) ? acc[0] : acc[1];
        float ra = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        float rb = acc[2] + __shfl_xor_sync(0xffffffffu, acc[2], 1);
        if (valid) {
            const float add = p.noise ? nw * p.noise[px % p.plane] : 0.f;
            ra += add + (p.bias ? p.bias[co_a] : 0.f);
            rb += add + (p.bias ? p.bias[co_b] : 0.f);
            if (p.activate) { ra = (ra > 0.f ? ra : neg_slope * ra) * gain; rb = (rb > 0.f ? rb : neg_slope * rb) * gain; }
            stf(y + px * PW + co_a, ra);
            if (!(c & 1)) stf(y + px * PW + co_b, rb);
        }
    }
}

// dx[p][ci] = sum_co dy[p][co] * w_t[ci][co]
template <typename T, bool REGW>
__global__ void __launch_bounds__(256) pw12_dgrad_kernel(const PwParams p) {
    extern __shared__ float ws[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, c = lane & 7;
    float wr[REGW ? PW : 1][8];
    if (REGW) {   // w_t[ci][co]: this lane's 8 input channels x 12
        const T* w = static_cast<const T*>(p.w);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int co = 0; co < (REGW ? PW : 1); ++co) wr[co][j] = ldf(w + (long)(c * 8 + j) * PW + co);
    } else {
        pw_stage_weights<T>(p, ws, true);
    }
    const T* dy = static_cast<const T*>(p.dy);
    T* dx = static_cast<T*>(p.y);
    for (long p4 = ((long)blockIdx.x * 8 + warp) * 4; p4 < p.pixels; p4 += (long)gridDim.x * 32) {
        const long px = p4 + g;
        if (px >= p.pixels) continue;
        float d[PW];
        ld12(dy + px * PW, d);
        if (REGW) {
            float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int co = 0; co < PW; ++co)
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(d[co], wr[REGW ? co : 0][j], o[j]);
            st8(dx + px * 64 + c * 8, o);
        } else {
            for (int c0 = 0; c0 < p.Cin; c0 += 64) {
                float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int co = 0; co < PW; ++co) {
                    const float4 w0 = *reinterpret_cast<const float4*>(ws + co * p.Cin + c0 + c * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(ws + co * p.Cin + c0 + c * 8 + 4);
                    o[0] += d[co] * w0.x; o[1] += d[co] * w0.y; o[2] += d[co] * w0.z; o[3] += d[co] * w0.w;
                    o[4] += d[co] * w1.x; o[5] += d[co] * w1.y; o[6] += d[co] * w1.z; o[7] += d[co] * w1.w;
                }
                st8(dx + px * p.Cin + c0 + c * 8, o);
            }
        }
    }
}

// dw[co][ci_offset + ci] += sum_p dy[p][co] * x[p][ci]; blockIdx.y = 64-channel slab of ci
template <typename T>
__global__ void __launch_bounds__(256) pw12_wgrad_kernel(const PwParams p) {
    __shared__ float red[PW * 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, c = lane & 7;
    const int c0 = blockIdx.y * 64;
    const T* x = static_cast<const T*>(p.x);
    const T* dy = static_cast<const T*>(p.dy);
    float acc[PW][8];
#pragma unroll
    for (int i = 0; i < PW; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int e = threadIdx.x; e < PW * 64; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
    for (long p4 = ((long)blockIdx.x * 8 + warp) * 4; p4 < p.pixels; p4 += (long)gridDim.x * 32) {
        const long px = p4 + g;
        if (px >= p.pixels) continue;
        float d[PW], xv[8];
        ld12(dy + px * PW, d);
        ld8<T>(x + px * p.Cin + c0 + c * 8, xv);
#pragma unroll
        for (int co = 0; co < PW; ++co)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[co][j] = fmaf(d[co], xv[j], acc[co][j]);
    }
#pragma unroll
    for (int co = 0; co < PW; ++co)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[co][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) atomicAdd(&red[co * 64 + c * 8 + j], v);
        }
    __syncthreads();
    for (int e = threadIdx.x; e < PW * 64; e += blockDim.x)
        atomicAdd(p.dw + (long)(e / 64) * p.ci_total + p.ci_offset + c0 + (e % 64), red[e]);
}

// ---- bf16 forms on the warp-level tensor-core instruction (mma.sync m16n8k16) ---------------------------------------------
// 12 outputs per pixel cannot fill a tcgen05 tile and their 24-byte rows cannot be TMA boxes, but the op is a (pixels x Cin)
// x (Cin x 12) product all the same: one HMMA per 16 pixels x 16 channels x 8 outputs keeps the kernel at the HBM roofline
// (the fp32-FMA form above needs 96 FMA per 16 bytes and is issue-bound).  The K (channel) and N (channel) orders of the
// fragments are permuted so that every lane's global accesses are whole 16-byte vectors:
//   forward: lane (gid, tig) loads channels [16 tig, 16 tig + 16) of rows gid and gid + 8; k-step s takes 4s .. 4s+3 of them;
//   dgrad:   output column n of n-tile j is channel (n/2)*16 + 2j + (n&1): a lane's 8 C fragments are 16 consecutive channels.
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) pw12_fwd_mma_kernel(const PwParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x);
    const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(p.w);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y);
    const float nw = (p.noise && p.noise_w) ? p.noise_w[0] : 0.f;
    const float gain = p.activate == 1 ? 1.4142135623730951f : 1.f, neg_slope = p.activate == 3 ? 0.f : 0.2f;
    const int slabs = p.Cin / 64;
    uint4 wq[2][2];   // [n-tile][half]: weights of output gid (+8) for this lane's 16 channels (slab 0 kept in registers)
    auto load_w = [&](int slab) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int co = j * 8 + gid;
            const uint4* src = reinterpret_cast<const uint4*>(w + (long)co * p.w_cin_total + p.w_cin_offset + slab * 64 + tig * 16);
            wq[j][0] = co < PW ? __ldg(src) : make_uint4(0, 0, 0, 0);
            wq[j][1] = co < PW ? __ldg(src + 1) : make_uint4(0, 0, 0, 0);
        }
    };
    load_w(0);
    const float b0 = p.bias ? p.bias[tig * 2] : 0.f, b1 = p.bias ? p.bias[tig * 2 + 1] : 0.f;
    const float b8 = (p.bias && tig < 2) ? p.bias[8 + tig * 2] : 0.f, b9 = (p.bias && tig < 2) ? p.bias[9 + tig * 2] : 0.f;
    for (long p16 = ((long)blockIdx.x * 8 + warp) * 16; p16 < p.pixels; p16 += (long)gridDim.x * 128) {
        const long r0 = p16 + gid, r1 = r0 + 8;
        const bool v0 = r0 < p.pixels, v1 = r1 < p.pixels;
        float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int slab = 0; slab < slabs; ++slab) {
            if (slabs > 1) load_w(slab);
            const uint4* s0 = reinterpret_cast<const uint4*>(x + r0 * p.Cin + slab * 64 + tig * 16);
            const uint4* s1 = reinterpret_cast<const uint4*>(x + r1 * p.Cin + slab * 64 + tig * 16);
            const uint4 z = make_uint4(0, 0, 0, 0);
            const uint4 q00 = v0 ? __ldg(s0) : z, q01 = v0 ? __ldg(s0 + 1) : z, q10 = v1 ? __ldg(s1) : z, q11 = v1 ? __ldg(s1 + 1) : z;
            const uint32_t ra[8] = {q00.x, q00.y, q00.z, q00.w, q01.x, q01.y, q01.z, q01.w};   // row gid: channel pairs 0..7
            const uint32_t rb[8] = {q10.x, q10.y, q10.z, q10.w, q11.x, q11.y, q11.z, q11.w};   // row gid + 8
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t wj[8] = {wq[j][0].x, wq[j][0].y, wq[j][0].z, wq[j][0].w, wq[j][1].x, wq[j][1].y, wq[j][1].z, wq[j][1].w};
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const uint32_t a[4] = {ra[2 * s4], rb[2 * s4], ra[2 * s4 + 1], rb[2 * s4 + 1]};
                    mma_bf16(c[j], a, wj[2 * s4], wj[2 * s4 + 1]);
                }
            }
        }
        auto fin = [&](float v, float b, long px) {
            v += b + (p.noise ? nw * p.noise[px % p.plane] : 0.f);
            if (p.activate) v = (v > 0.f ? v : neg_slope * v) * gain;
            return v;
        };
        if (v0) {
            *reinterpret_cast<__nv_bfloat162*>(y + r0 * PW + tig * 2) = __floats2bfloat162_rn(fin(c[0][0], b0, r0), fin(c[0][1], b1, r0));
            if (tig < 2) *reinterpret_cast<__nv_bfloat162*>(y + r0 * PW + 8 + tig * 2) = __floats2bfloat162_rn(fin(c[1][0], b8, r0), fin(c[1][1], b9, r0));
        }
        if (v1) {
            *reinterpret_cast<__nv_bfloat162*>(y + r1 * PW + tig * 2) = __floats2bfloat162_rn(fin(c[0][2], b0, r1), fin(c[0][3], b1, r1));
            if (tig < 2) *reinterpret_cast<__nv_bfloat162*>(y + r1 * PW + 8 + tig * 2) = __floats2bfloat162_rn(fin(c[1][2], b8, r1), fin(c[1][3], b9, r1));
        }
    }
}

__global__ void __launch_bounds__(256) pw12_dgrad_mma_kernel(const PwParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(p.dy);
    const __nv_bfloat16* wt = static_cast<const __nv_bfloat16*>(p.w);   // w_t[ci][12]
    __nv_bfloat16* dx = static_cast<__nv_bfloat16*>(p.y);
    const int slabs = p.Cin / 64;
    uint32_t wb[8][2];   // [n-tile j][k half]: column n = gid of n-tile j is channel (gid/2)*16 + 2j + (gid&1)
    auto load_w = [&](int slab) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ci = slab * 64 + (gid >> 1) * 16 + 2 * j + (gid & 1);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(wt + (long)ci * PW);
            wb[j][0] = __ldg(src + tig);
            wb[j][1] = tig < 2 ? __ldg(src + 4 + tig) : 0u;
        }
    };
    load_w(0);
    for (long p16 = ((long)blockIdx.x * 8 + warp) * 16; p16 < p.pixels; p16 += (long)gridDim.x * 128) {
        const long r0 = p16 + gid, r1 = r0 + 8;
        const bool v0 = r0 < p.pixels, v1 = r1 < p.pixels;
        const uint32_t* d0 = reinterpret_cast<const uint32_t*>(dy + r0 * PW);
        const uint32_t* d1 = reinterpret_cast<const uint32_t*>(dy + r1 * PW);
        const uint32_t a[4] = {v0 ? __ldg(d0 + tig) : 0u, v1 ? __ldg(d1 + tig) : 0u, (v0 && tig < 2) ? __ldg(d0 + 4 + tig) : 0u,
                               (v1 && tig < 2) ? __ldg(d1 + 4 + tig) : 0u};
        for (int slab = 0; slab < slabs; ++slab) {
            if (slabs > 1) load_w(slab);
            float c[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f; mma_bf16(c[j], a, wb[j][0], wb[j][1]); }
            uint4 o0[2], o1[2];
            __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(o0);
            __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(o1);
#pragma unroll
            for (int j = 0; j < 8; ++j) { h0[j] = __floats2bfloat162_rn(c[j][0], c[j][1]); h1[j] = __floats2bfloat162_rn(c[j][2], c[j][3]); }
            if (v0) { uint4* d = reinterpret_cast<uint4*>(dx + r0 * p.Cin + slab * 64 + tig * 16); d[0] = o0[0]; d[1] = o0[1]; }
            if (v1) { uint4* d = reinterpret_cast<uint4*>(dx + r1 * p.Cin + slab * 64 + tig * 16); d[0] = o1[0]; d[1] = o1[1]; }
        }
    }
}

static bool pw12_ok(const AgrConvGeom& g) {
    return g.ksize == 1 && g.stride == 1 && g.pad == 0 && g.Cout == PW && g.Cin % 64 == 0 && g.Cin <= 1024;
}

template <typename T>
static int launch_pw12(int what, PwParams& p, cudaStream_t s) {
    const long groups = (p.pixels + 31) / 32;
    unsigned blocks = (unsigned)(groups < 148 * 8 ? (groups < 1 ? 1 : groups) : 148 * 8);
    const int smem = PW * p.Cin * (int)sizeof(float);
    if (what == 2) {
        pw12_wgrad_kernel<T><<<dim3(blocks > 148 * 2 ? 148 * 2 : blocks, p.Cin / 64), 256, 0, s>>>(p);
    } else if (sizeof(T) == 2 && (what == 1 || (p.w_cin_total % 8 == 0 && p.w_cin_offset % 8 == 0))) {
        const long g16 = (p.pixels + 127) / 128;
        const unsigned b2 = (unsigned)(g16 < 148 * 8 ? (g16 < 1 ? 1 : g16) : 148 * 8);
        if (what == 0) pw12_fwd_mma_kernel<<<b2, 256, 0, s>>>(p); else pw12_dgrad_mma_kernel<<<b2, 256, 0, s>>>(p);
    } else if (p.Cin == 64) {
        if (what == 0) pw12_fwd_kernel<T, true><<<blocks, 256, 0, s>>>(p); else pw12_dgrad_kernel<T, true><<<blocks, 256, 0, s>>>(p);
    } else {
        auto kern = what == 0 ? pw12_fwd_kernel<T, false> : pw12_dgrad_kernel<T, false>;
        if (smem > 48 * 1024) {
            static bool attr[2] = {false, false};
            if (!attr[what]) { if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess) return AGR_ERR_CUDA; attr[what] = true; }
        }
        kern<<<blocks, 256, smem, s>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

template <typename T>
__global__ void __launch_bounds__(256) weight_transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int Cout, int Cin, int taps) {
    __shared__ T tile[32][33];
    const int t = blockIdx.z;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        if (co < Cout && ci < Cin) tile[r][tx] = in[((size_t)co * taps + t) * Cin + ci];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cin && co < Cout) out[((size_t)ci * taps + t) * Cout + co] = tile[tx][r];
    }
}

static DirectParams make_params(const AgrConvGeom& g) {
    DirectParams p{};
    p.N = g.N; p.H = g.H; p.W = g.W; p.Cin = g.Cin; p.OH = g.OH; p.OW = g.OW; p.Cout = g.Cout;
    p.k = g.ksize; p.stride = g.stride; p.pad = g.pad; p.transposed = g.transposed;
    return p;
}

template <typename TI, typename TO>
static int launch_fwd_t(const DirectParams& p, cudaStream_t s) {
    const long total = (long)p.N * p.OH * p.OW;
    const bool vec = (p.Cin % KC == 0) && (p.w_cin_total % 4 == 0) && (p.w_cin_offset % 4 == 0);
    const bool narrow = p.Cout <= 32;
    const int BN = narrow ? 16 : 64;
    dim3 grid((unsigned)((total + PT - 1) / PT), (unsigned)((p.Cout + BN - 1) / BN));
    if (narrow) {
        if (vec) conv_direct_kernel<TI, TO, 16, true><<<grid, 256, 0, s>>>(p); else conv_direct_kernel<TI, TO, 16, false><<<grid, 256, 0, s>>>(p);
    } else {
        if (vec) conv_direct_kernel<TI, TO, 64, true><<<grid, 256, 0, s>>>(p); else conv_direct_kernel<TI, TO, 64, false><<<grid, 256, 0, s>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int launch_forward(int dtype, const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, cudaStream_t s) {
    if (!tc::geom_ok(g) || ep.residual || ep.out_fp32) return AGR_ERR_INVALID_ARGUMENT;
    DirectParams p = make_params(g);
    p.x = x; p.w = w; p.y = y; p.bias = ep.bias; p.noise = ep.noise; p.noise_w = ep.noise_w; p.activate = ep.activate;
    p.w_cin_total = ep.w_cin_total > 0 ? ep.w_cin_total : g.Cin; p.w_cin_offset = ep.w_cin_offset;
    if (p.w_cin_offset < 0 || p.w_cin_offset + g.Cin > p.w_cin_total) return AGR_ERR_INVALID_ARGUMENT;
    const bool dgrad_of_pw12 = g.transposed && g.ksize == 1 && g.stride == 1 && g.pad == 0 && g.Cin == PW && g.Cout % 64 == 0 && g.Cout <= 1024 &&
                               !ep.bias && !ep.noise && !ep.activate && p.w_cin_total == g.Cin;
    if ((pw12_ok(g) && !g.transposed && p.w_cin_total % 4 == 0 && p.w_cin_offset % 4 == 0) || dgrad_of_pw12) {
        PwParams q{};
        q.w = w; q.y = y; q.pixels = (long)g.N * g.OH * g.OW; q.plane = (long)g.OH * g.OW;
        q.bias = ep.bias; q.noise = ep.noise; q.noise_w = ep.noise_w; q.activate = ep.activate;
        if (dgrad_of_pw12) { q.dy = x; q.Cin = g.Cout; }                    // adjoint geometry: x is dy (12 ch), y is dx (Cin ch)
        else { q.x = x; q.Cin = g.Cin; q.w_cin_total = p.w_cin_total; q.w_cin_offset = p.w_cin_offset; }
        if (dtype == AGR_BF16) return launch_pw12<__nv_bfloat16>(dgrad_of_pw12 ? 1 : 0, q, s);
        if (dtype == AGR_F32) return launch_pw12<float>(dgrad_of_pw12 ? 1 : 0, q, s);
        return AGR_ERR_INVALID_ARGUMENT;
    }
    if (dtype == AGR_BF16) return launch_fwd_t<__nv_bfloat16, __nv_bfloat16>(p, s);
    if (dtype == AGR_F32) return launch_fwd_t<float, float>(p, s);
    return AGR_ERR_INVALID_ARGUMENT;
}

int launch_wgrad(int dtype, const AgrConvGeom& g, const void* x, const void* dy, float* dw, int ci_total, int ci_offset, cudaStream_t s) {
    if (!tc::geom_ok(g)) return AGR_ERR_INVALID_ARGUMENT;
    if (g.transposed && (g.Cin % 64)) return AGR_ERR_INVALID_ARGUMENT;   // the (t,ci) tile must sit inside one tap
    if (pw12_ok(g) && !g.transposed) {
        PwParams q{};
        q.x = x; q.dy = dy; q.dw = dw; q.pixels = (long)g.N * g.OH * g.OW; q.Cin = g.Cin; q.ci_total = ci_total; q.ci_offset = ci_offset;
        if (dtype == AGR_BF16) return launch_pw12<__nv_bfloat16>(2, q, s);
        if (dtype == AGR_F32) return launch_pw12<float>(2, q, s);
        return AGR_ERR_INVALID_ARGUMENT;
    }
    DirectParams p = make_params(g);
    p.x = x; p.dy = dy; p.dw = dw; p.ci_total = ci_total; p.ci_offset = ci_offset;
    const int NN = g.ksize * g.ksize * g.Cin;
    const bool narrow = g.Cout <= 32;          // 16-row tile
    const bool thin = !narrow && NN <= 32;     // 16-column tile: 1- and 3-channel inputs with few taps
    const int BMc = narrow ? 16 : 64, NT = thin ? 16 : 64;
    const long total = (long)g.N * (g.transposed ? (long)g.H * g.W : (long)g.OH * g.OW);
    const long tiles = (long)((NN + NT - 1) / NT) * ((g.Cout + BMc - 1) / BMc);
    long slices = (4 * 148 + tiles - 1) / tiles;
    if (slices > total / 64) slices = total / 64;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    p.slices = (int)slices;
    dim3 grid((unsigned)((NN + NT - 1) / NT), (unsigned)((g.Cout + BMc - 1) / BMc), (unsigned)slices);
    if (dtype == AGR_BF16) {
        using T = __nv_bfloat16;
        if (narrow) conv_direct_wgrad_kernel<T, 16, 64><<<grid, 256, 0, s>>>(p);
        else if (thin) conv_direct_wgrad_kernel<T, 64, 16><<<grid, 256, 0, s>>>(p);
        else conv_direct_wgrad_kernel<T, 64, 64><<<grid, 256, 0, s>>>(p);
    } else if (dtype == AGR_F32) {
        if (narrow) conv_direct_wgrad_kernel<float, 16, 64><<<grid, 256, 0, s>>>(p);
        else if (thin) conv_direct_wgrad_kernel<float, 64, 16><<<grid, 256, 0, s>>>(p);
        else conv_direct_wgrad_kernel<float, 64, 64><<<grid, 256, 0, s>>>(p);
    } else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int launch_transpose(int dtype, const void* in, void* out, int Cout, int Cin, int taps, cudaStream_t s) {
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, taps);
    if (dtype == AGR_BF16) weight_transpose_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), Cout, Cin, taps);
    else if (dtype == AGR_F32) weight_transpose_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(in), static_cast<float*>(out), Cout, Cin, taps);
    else return AGR_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // namespace direct
}  // namespace agr

// ---- C entry points (include/agr_conv.h): shape-based choice between the two paths -------------------------------------
namespace agr {
namespace tc {
bool forward_supported(const AgrConvGeom& g);
int launch_forward(const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, bool w_mn, cudaStream_t s);
bool wgrad_supported(const AgrConvGeom& g);
int launch_wgrad(const AgrConvGeom& g, const void* x, const void* dy, float* dw, int ci_total, int ci_offset, cudaStream_t s);
}  // namespace tc
}  // namespace agr

extern "C" {

int agr_conv2d_path(int32_t dtype, const AgrConvGeom* g, int32_t what) {
    using namespace agr;
    if (!g || !tc::geom_ok(*g) || (dtype != AGR_F32 && dtype != AGR_BF16)) return 0;
    if (dtype == AGR_BF16) {
        if (what == 0 && tc::forward_supported(*g)) return 1;
        if (what == 1 && tc::forward_supported(tc::adjoint(*g))) return 1;
        if (what == 2 && tc::wgrad_supported(*g)) return 1;
    }
    if (what == 2 && g->transposed && (g->Cin % 64)) return 0;
    return 2;
}

int agr_conv2d_forward(int32_t dtype, const AgrConvGeom* g, const void* x, const void* w_krsc, void* y, const AgrConvEpilogue* ep,
                       void* cuda_stream) {
    using namespace agr;
    if (!g || !x || !w_krsc || !y) return AGR_ERR_INVALID_ARGUMENT;
    AgrConvEpilogue e{};
    if (ep) e = *ep;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int path = agr_conv2d_path(dtype, g, 0);
    if (path == 1) return tc::launch_forward(*g, x, w_krsc, y, e, false, s);
    if (path == 2) return direct::launch_forward(dtype, *g, x, w_krsc, y, e, s);
    return AGR_ERR_INVALID_ARGUMENT;
}

int agr_conv2d_dgrad_krsc(int32_t dtype, const AgrConvGeom* g, const void* dy, const void* w_krsc, int32_t w_cin_total, int32_t w_cin_offset,
                          void* dx, void* cuda_stream) {
    using namespace agr;
    if (!g || !dy || !w_krsc || !dx || !tc::geom_ok(*g)) return AGR_ERR_INVALID_ARGUMENT;
    if (agr_conv2d_path(dtype, g, 1) != 1) return AGR_ERR_INVALID_ARGUMENT;   // path 2 contracts with the transposed operand
    const AgrConvGeom a = tc::adjoint(*g);
    AgrConvEpilogue e{};
    e.w_cin_total = w_cin_total; e.w_cin_offset = w_cin_offset;
    return tc::launch_forward(a, dy, w_krsc, dx, e, true, static_cast<cudaStream_t>(cuda_stream));
}

int agr_conv2d_dgrad(int32_t dtype, const AgrConvGeom* g, const void* dy, const void* w_t, void* dx, void* cuda_stream) {
    using namespace agr;
    if (!g || !dy || !w_t || !dx || !tc::geom_ok(*g)) return AGR_ERR_INVALID_ARGUMENT;
    const AgrConvGeom a = tc::adjoint(*g);
    return agr_conv2d_forward(dtype, &a, dy, w_t, dx, nullptr, cuda_stream);
}

int agr_conv2d_wgrad(int32_t dtype, const AgrConvGeom* g, const void* x, const void* dy, float* dw, int32_t ci_total, int32_t ci_offset,
                     int32_t zero_first, void* cuda_stream) {
    using namespace agr;
    if (!g || !x || !dy || !dw || !tc::geom_ok(*g)) return AGR_ERR_INVALID_ARGUMENT;
    if (ci_total <= 0) ci_total = g->Cin;
    if (ci_offset < 0 || ci_offset + g->Cin > ci_total) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (zero_first && cudaMemsetAsync(dw, 0, (size_t)g->Cout * g->ksize * g->ksize * ci_total * sizeof(float), s) != cudaSuccess) return AGR_ERR_CUDA;
    const int path = agr_conv2d_path(dtype, g, 2);
    if (path == 1) return tc::launch_wgrad(*g, x, dy, dw, ci_total, ci_offset, s);
    if (path == 2) return direct::launch_wgrad(dtype, *g, x, dy, dw, ci_total, ci_offset, s);
    return AGR_ERR_INVALID_ARGUMENT;
}

int agr_weight_transpose(int32_t dtype, const void* w_krsc, void* w_out, int32_t Cout, int32_t Cin, int32_t ksize, void* cuda_stream) {
    if (!w_krsc || !w_out || Cout < 1 || Cin < 1 || ksize < 1) return AGR_ERR_INVALID_ARGUMENT;
    return agr::direct::launch_transpose(dtype, w_krsc, w_out, Cout, Cin, ksize * ksize, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
