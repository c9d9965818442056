// Implicit-GEMM 3x3 / 1x1 convolution (stride 1, "same" padding) on the Blackwell tensor cores:
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), operands staged in shared memory by TMA,
// warp-specialised (1 TMA producer warp, 1 MMA issuer warp, 4 epilogue warps), mbarrier pipelines.
//
// Replaces the reference's dense contractions, which are cuDNN calls (conv2d_gradfix.py:34,66 via
// dual_styleunet.py:114,275-296) followed by separate noise / bias / activation passes
// (dual_styleunet.py:598-604): here noise injection + bias + leaky-ReLU(0.2)*sqrt(2) run in the epilogue,
// straight out of TMEM.
//
// GEMM view (batch 1, NHWC bf16):  Y[p][co] = sum_{tap} sum_{ci} X[p + off(tap)][ci] * Wt[co][tap][ci]
//   M = 128 output pixels = one 16 (w) x 8 (h) patch,   N = BN output channels,   K = taps * Cin in 64-blocks.
// "im2col" never exists in memory: for each tap the A tile is ONE 3-D TMA box {64 ch, 16 w, 8 h} of the
// activation, fetched at the tap's (dx, dy) shift; TMA zero-fills the out-of-image part (= the conv padding) and
// writes the 128B-swizzled K-major layout tcgen05.mma reads directly.  B tiles are {64 ci, 1 tap, BN co} boxes of
// the KRSC weight.  fp32 accumulators (128 lanes x BN columns) live in TMEM.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_rasterizer.h"
#include "../../include/agr_styleunet.h"

namespace agr {
namespace tc {

constexpr int TILE_W = 16, TILE_H = 8, BM = TILE_W * TILE_H;  // 128 pixels
constexpr int BK = 64;                                       // channels per k-block (128 B of bf16)
// pipeline depth is a template parameter: 4 stages when the grid is a single wave (1 CTA/SM anyway), 3 stages (96 KB at
// BN=128) when there are more tiles than SMs so that two CTAs per SM overlap one's epilogue with the other's mainloop
constexpr int A_BYTES = BM * BK * 2;                         // 16 KB
constexpr int NUM_THREADS = 192;                             // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TCW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TCD_%=;\n\t"
        "bra TCW_%=;\n\t"
        "TCD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start >> 4, [16,30) LBO >> 4 (=1, unused for swizzled K-major), [32,46) SBO >> 4 (= 1024 B: 8 rows x 128 B),
//   [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10, K-major both,
// N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

struct ConvParams {
    int N, H, W, Cin, Cout, taps, ksize;  // taps = ksize*ksize; N images share the weights
    const float* bias;     // (Cout) or null
    const float* noise;    // (H*W) or null
    const float* noise_w;  // (1) or null
    const float* residual;  // (H, W, Cout) fp32 added before bias/activation, shared by the N images; or null
    float* y_f32;           // when set, the result is stored in fp32 here instead of bf16 in y (partial sums)
    int w_cin_offset;      // first input channel of the weight slice this call contracts with (split-K over a concat)
    int splits;            // > 1: blockIdx.z owns a contiguous slice of the (tap, channel-block) loop and ADDS its fp32
                           // partial tile into y_f32 (zeroed by the caller); bias / noise / activation run in conv_finish_kernel
    int activate;
    __nv_bfloat16* y;      // (H, W, Cout)
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, (STAGES * (A_BYTES + BN * BK * 2) + 1024) * 2 <= 227 * 1024 ? 2 : 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, ConvParams p) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = p.W / TILE_W;
    const int tiles_img = tiles_w * (p.H / TILE_H);
    const int img = blockIdx.x / tiles_img;
    const int tile_m = blockIdx.x - img * tiles_img;
    const int h0 = (tile_m / tiles_w) * TILE_H, w0 = (tile_m % tiles_w) * TILE_W;
    const int n0 = blockIdx.y * BN;
    const int kchunks = p.Cin / BK;
    const int num_kb_all = p.taps * kchunks;
    const int kb0 = p.splits > 1 ? (int)(((long)blockIdx.z * num_kb_all) / p.splits) : 0;
    const int kb1 = p.splits > 1 ? (int)(((long)(blockIdx.z + 1) * num_kb_all) / p.splits) : num_kb_all;
    const int num_kb = kb1 - kb0;   // >= 1: the host keeps splits <= num_kb_all
    const int pad = p.ksize / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1) {  // TMEM allocation: BN fp32 columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ================= TMA producer (one elected lane) =================
        if (lane == 0) {
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int kb = kb0 + it;
                const int tap = kb / kchunks, ck = kb - tap * kchunks;
                const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
                unsigned char* a_dst = smem + s * STAGE_BYTES;
                unsigned char* b_dst = a_dst + A_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_4d(a_dst, &map_x, &full_bar[s], ck * BK, w0 + dx, h0 + dy, img);   // OOB -> zeros = padding
                tma_load_3d(b_dst, &map_w, &full_bar[s], p.w_cin_offset + ck * BK, tap, n0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected lane) =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(BM, BN);
            for (int it = 0; it < num_kb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BYTES;
                const uint64_t adesc = umma_desc(a_addr), bdesc = umma_desc(b_addr);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16 bf16 = 32 B -> +2 in the (>>4) start-address field
                    umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[s]);      // frees the smem stage once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);         // accumulator complete
        }
    } else {
        // ================= epilogue: TMEM -> registers -> (+noise, +bias, lrelu) -> global =================
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;             // pixel within the tile
        const int h = h0 + row / TILE_W, w = w0 + row % TILE_W;
        const size_t pix = (size_t)h * p.W + w;   // noise is per pixel, shared by the N images
        const float add = (p.noise && p.noise_w) ? p.noise_w[0] * p.noise[pix] : 0.f;
        __nv_bfloat16* out = p.y + ((size_t)img * p.H * p.W + pix) * p.Cout + n0;
        const float* res = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;
        float* out32 = p.y_f32 ? p.y_f32 + ((size_t)img * p.H * p.W + pix) * p.Cout + n0 : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint4 packed[4];
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(packed);
            if (res) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 rr = reinterpret_cast<const float4*>(res + c0)[i];
                    r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) + rr.x);
                    r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + rr.y);
                    r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + rr.z);
                    r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + rr.w);
                }
            }
            if (p.splits > 1) {   // split-K partial tile: accumulate into the zeroed fp32 workspace
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(out32 + c0 + i, __uint_as_float(r[i]));
                continue;
            }
            if (out32) {   // fp32 partial result (no epilogue math): consumed as `residual` by the second half
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    reinterpret_cast<float4*>(out32 + c0)[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                                           __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                continue;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float v0 = __uint_as_float(r[2 * i]) + add, v1 = __uint_as_float(r[2 * i + 1]) + add;
                if (p.bias) { v0 += p.bias[n0 + c0 + 2 * i]; v1 += p.bias[n0 + c0 + 2 * i + 1]; }
                if (p.activate) {
                    v0 = (v0 > 0.f ? v0 : 0.2f * v0) * 1.4142135623730951f;
                    v1 = (v1 > 0.f ? v1 : 0.2f * v1) * 1.4142135623730951f;
                }
                h2[i] = __floats2bfloat162_rn(v0, v1);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = packed[i];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

// ---- split-K finish: y = act(acc + noise_w * noise + bias) in bf16 from the fp32 accumulation buffer
__global__ void __launch_bounds__(256) conv_finish_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ y, long pixels,
                                                         long pixels_per_image, int Cout, const float* __restrict__ bias,
                                                         const float* __restrict__ noise, const float* __restrict__ noise_w,
                                                         int activate) {
    const int cv = Cout / 8;
    const long total = pixels * cv;
    const float nw = (noise && noise_w) ? noise_w[0] : 0.f;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * 8;
        const long p = idx / cv;
        const float add = noise ? nw * noise[p % pixels_per_image] : 0.f;
        const float4 a = __ldg(reinterpret_cast<const float4*>(acc + p * Cout + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(acc + p * Cout + c + 4));
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint4 packed;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&packed);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = v[i] + add;
            if (bias) t += bias[c + i];
            if (activate) t = (t > 0.f ? t : 0.2f * t) * 1.4142135623730951f;
            v[i] = t;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(y + p * Cout + c) = packed;
    }
}

// ---- KRSC weight -> the weight of the data-gradient convolution: out[ci][taps-1-t][co] = in[co][t][ci]
__global__ void __launch_bounds__(256) weight_flip_transpose_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                                   int Cout, int Cin, int taps) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int t = blockIdx.z;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        tile[r][tx] = (co < Cout && ci < Cin) ? in[((size_t)co * taps + t) * Cin + ci] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cin && co < Cout) out[((size_t)ci * taps + (taps - 1 - t)) * Cout + co] = tile[tx][r];
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && ptr)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static bool make_map_4d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t b0, uint32_t b1,
                        uint32_t b2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {d0 * 2, d0 * d1 * 2, d0 * d1 * d2 * 2};
    cuuint32_t box[4] = {b0, b1, b2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_map_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};  // bytes, dims 1..2
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
static int launch_s(const CUtensorMap& mx, const CUtensorMap& mw, const ConvParams& p, cudaStream_t s) {
    constexpr int smem = STAGES * (A_BYTES + BN * BK * 2) + 1024;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AGR_ERR_CUDA;
        attr = true;
    }
    dim3 grid(p.N * (p.H / TILE_H) * (p.W / TILE_W), p.Cout / BN, p.splits > 1 ? p.splits : 1);
    conv_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, smem, s>>>(mx, mw, p);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

template <int BN>
static int launch(const void* x, const void* w, const ConvParams& p, int w_cin_total, cudaStream_t s) {
    CUtensorMap mx, mw;
    if (!make_map_4d(&mx, x, (uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N, BK, TILE_W, TILE_H)) return AGR_ERR_CUDA;
    if (!make_map_3d(&mw, w, (uint64_t)w_cin_total, (uint64_t)p.taps, (uint64_t)p.Cout, BK, 1, BN)) return AGR_ERR_CUDA;
    const long tiles = (long)p.N * (p.H / TILE_H) * (p.W / TILE_W) * (p.Cout / BN) * (p.splits > 1 ? p.splits : 1);
    if (BN == 64 || tiles <= 148) return launch_s<BN, 4>(mx, mw, p, s);
    return launch_s<BN, 3>(mx, mw, p, s);
}

}  // namespace tc
}  // namespace agr

extern "C" {

int agr_conv2d_tc_supported(int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize) {
    using namespace agr::tc;
    if (ksize != 1 && ksize != 3) return 0;
    if (H < TILE_H || W < TILE_W || (H % TILE_H) || (W % TILE_W)) return 0;
    if (Cin % BK) return 0;
    if (Cout % 64) return 0;
    return 1;
}

int agr_conv2d_tc_forward(const void* x, const void* w_krsc, void* y, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                          int32_t ksize, const float* bias, const float* noise, const float* noise_w, int32_t activate,
                          void* cuda_stream) {
    using namespace agr::tc;
    if (!x || !w_krsc || !y || N < 1 || !agr_conv2d_tc_supported(H, W, Cin, Cout, ksize)) return AGR_ERR_INVALID_ARGUMENT;
    ConvParams p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize; p.taps = ksize * ksize;
    p.bias = bias; p.noise = noise; p.noise_w = noise_w; p.activate = activate; p.y = static_cast<__nv_bfloat16*>(y);
    p.residual = nullptr; p.y_f32 = nullptr; p.w_cin_offset = 0; p.splits = 1;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (Cout % 128 == 0) return launch<128>(x, w_krsc, p, Cin, s);
    return launch<64>(x, w_krsc, p, Cin, s);
}

int agr_conv2d_tc_splits(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize) {
    using namespace agr::tc;
    if (!agr_conv2d_tc_supported(H, W, Cin, Cout, ksize)) return 1;
    const int BN = (Cout % 128 == 0) ? 128 : 64;
    const long tiles = (long)N * (H / TILE_H) * (W / TILE_W) * (Cout / BN);
    const int num_kb = ksize * ksize * (Cin / BK);
    // fill ~2 CTAs per SM, but keep >= 4 (tap, channel-block) steps per CTA so the pipeline still overlaps
    long s = (2 * 148) / tiles;
    if (s > num_kb / 4) s = num_kb / 4;
    if (s > 32) s = 32;
    return s < 2 ? 1 : (int)s;
}

int agr_conv2d_tc_forward_splitk(const void* x, const void* w_krsc, void* y, float* workspace, int32_t splits, int32_t N, int32_t H,
                                 int32_t W, int32_t Cin, int32_t Cout, int32_t ksize, const float* bias, const float* noise,
                                 const float* noise_w, int32_t activate, void* cuda_stream) {
    using namespace agr::tc;
    if (!x || !w_krsc || !y || !workspace || N < 1 || !agr_conv2d_tc_supported(H, W, Cin, Cout, ksize)) return AGR_ERR_INVALID_ARGUMENT;
    if (splits < 2 || splits > ksize * ksize * (Cin / BK)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const long pixels = (long)N * H * W;
    if (cudaMemsetAsync(workspace, 0, (size_t)pixels * Cout * sizeof(float), s) != cudaSuccess) return AGR_ERR_CUDA;
    ConvParams p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize; p.taps = ksize * ksize;
    p.bias = nullptr; p.noise = nullptr; p.noise_w = nullptr; p.activate = 0; p.y = nullptr;
    p.residual = nullptr; p.y_f32 = workspace; p.w_cin_offset = 0; p.splits = splits;
    const int st = (Cout % 128 == 0) ? launch<128>(x, w_krsc, p, Cin, s) : launch<64>(x, w_krsc, p, Cin, s);
    if (st != AGR_OK) return st;
    const long total = pixels * (Cout / 8);
    long g = (total + 255) / 256; if (g > 148 * 8) g = 148 * 8;
    conv_finish_kernel<<<(unsigned)g, 256, 0, s>>>(workspace, static_cast<__nv_bfloat16*>(y), pixels, (long)H * W, Cout, bias, noise, noise_w, activate);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_conv2d_tc_forward_split(const void* x, const void* w_krsc, void* y, int32_t out_fp32, int32_t N, int32_t H, int32_t W,
                                int32_t Cin, int32_t Cout, int32_t ksize, int32_t w_cin_total, int32_t w_cin_offset,
                                const float* residual, const float* bias, int32_t activate, void* cuda_stream) {
    using namespace agr::tc;
    if (!x || !w_krsc || !y || N < 1 || !agr_conv2d_tc_supported(H, W, Cin, Cout, ksize)) return AGR_ERR_INVALID_ARGUMENT;
    if (w_cin_offset < 0 || (w_cin_offset % BK) || w_cin_offset + Cin > w_cin_total) return AGR_ERR_INVALID_ARGUMENT;
    ConvParams p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize; p.taps = ksize * ksize;
    p.bias = bias; p.noise = nullptr; p.noise_w = nullptr; p.activate = activate;
    p.y = out_fp32 ? nullptr : static_cast<__nv_bfloat16*>(y);
    p.y_f32 = out_fp32 ? static_cast<float*>(y) : nullptr;
    p.residual = residual; p.w_cin_offset = w_cin_offset; p.splits = 1;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (Cout % 128 == 0) return launch<128>(x, w_krsc, p, w_cin_total, s);
    return launch<64>(x, w_krsc, p, w_cin_total, s);
}

int agr_weight_flip_transpose(const void* w_krsc, void* w_out, int32_t Cout, int32_t Cin, int32_t ksize, void* cuda_stream) {
    if (!w_krsc || !w_out || Cout < 1 || Cin < 1 || ksize < 1) return AGR_ERR_INVALID_ARGUMENT;
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, ksize * ksize);
    agr::tc::weight_flip_transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const __nv_bfloat16*>(w_krsc), static_cast<__nv_bfloat16*>(w_out), Cout, Cin, ksize * ksize);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // extern "C"
