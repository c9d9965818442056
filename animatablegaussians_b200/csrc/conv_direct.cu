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
    const float gain = p.activate == 1 ? 1.4142135623730951f : 1.f;
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
            if (p.activate) v = (v > 0.f ? v : 0.2f * v) * gain;
            stf(y + P * p.Cout + co, v);
        }
    }
}

// dw[co][t][ci] += sum_q dy[a(q,t)][co] * x[b(q,t)][ci]; q runs over the plain operand's pixels:
//   convolution:             a = q (output pixel),  b = q*stride + k - pad
//   transposed convolution:  b = q (input pixel),   a = q*stride + k - pad    (needs the 64-wide (t,ci) tile inside one tap)
template <typename T, int BMc>
__global__ void __launch_bounds__(256) conv_direct_wgrad_kernel(const DirectParams p) {
    __shared__ float As[KC][BMc + 4];     // [pixel][co]
    __shared__ float Bs[KC][64 + 4];      // [pixel][(t,ci)]
    constexpr int CM = BMc / 16;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int taps = p.k * p.k, NN = taps * p.Cin;
    const int n0 = blockIdx.x * 64, co0 = blockIdx.y * BMc;
    const int GH = p.transposed ? p.H : p.OH, GW = p.transposed ? p.W : p.OW;
    const long total = (long)p.N * GH * GW;
    const long q_begin = (blockIdx.z * total) / p.slices, q_end = ((blockIdx.z + 1) * total) / p.slices;
    const T* x = static_cast<const T*>(p.x);
    const T* dy = static_cast<const T*>(p.dy);

    // this thread's B column (fixed): n = n0 + tid % 64 -> (tap, ci); A column: co = co0 + tid % BMc
    const int bn = n0 + (tid & 63);
    const bool bvalid = bn < NN;
    const int bt = bvalid ? bn / p.Cin : 0, bci = bvalid ? bn - bt * p.Cin : 0;
    const int at = n0 / p.Cin;                      // tap of the whole tile (transposed mode: tile inside one tap)
    const int aco = co0 + (tid % BMc);
    const int bk = tid >> 6;                        // B pixel row: bk + 4 i
    const int ak = tid / BMc;                       // A pixel row: ak + (256 / BMc) i

    float acc[CM][4];
#pragma unroll
    for (int i = 0; i < CM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long q0 = q_begin; q0 < q_end; q0 += KC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = bk + 4 * i;
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
            Bs[kk][tid & 63] = b;
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
    if (dtype == AGR_BF16) return launch_fwd_t<__nv_bfloat16, __nv_bfloat16>(p, s);
    if (dtype == AGR_F32) return launch_fwd_t<float, float>(p, s);
    return AGR_ERR_INVALID_ARGUMENT;
}

int launch_wgrad(int dtype, const AgrConvGeom& g, const void* x, const void* dy, float* dw, int ci_total, int ci_offset, cudaStream_t s) {
    if (!tc::geom_ok(g)) return AGR_ERR_INVALID_ARGUMENT;
    if (g.transposed && (g.Cin % 64)) return AGR_ERR_INVALID_ARGUMENT;   // the (t,ci) tile must sit inside one tap
    DirectParams p = make_params(g);
    p.x = x; p.dy = dy; p.dw = dw; p.ci_total = ci_total; p.ci_offset = ci_offset;
    const int NN = g.ksize * g.ksize * g.Cin;
    const bool narrow = g.Cout <= 32;
    const int BMc = narrow ? 16 : 64;
    const long total = (long)g.N * (g.transposed ? (long)g.H * g.W : (long)g.OH * g.OW);
    const long tiles = (long)((NN + 63) / 64) * ((g.Cout + BMc - 1) / BMc);
    long slices = (4 * 148 + tiles - 1) / tiles;
    if (slices > total / 64) slices = total / 64;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    p.slices = (int)slices;
    dim3 grid((unsigned)((NN + 63) / 64), (unsigned)((g.Cout + BMc - 1) / BMc), (unsigned)slices);
    if (dtype == AGR_BF16) {
        if (narrow) conv_direct_wgrad_kernel<__nv_bfloat16, 16><<<grid, 256, 0, s>>>(p); else conv_direct_wgrad_kernel<__nv_bfloat16, 64><<<grid, 256, 0, s>>>(p);
    } else if (dtype == AGR_F32) {
        if (narrow) conv_direct_wgrad_kernel<float, 16><<<grid, 256, 0, s>>>(p); else conv_direct_wgrad_kernel<float, 64><<<grid, 256, 0, s>>>(p);
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
int launch_forward(const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, cudaStream_t s);
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
    if (path == 1) return tc::launch_forward(*g, x, w_krsc, y, e, s);
    if (path == 2) return direct::launch_forward(dtype, *g, x, w_krsc, y, e, s);
    return AGR_ERR_INVALID_ARGUMENT;
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
