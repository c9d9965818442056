// Implicit-GEMM convolutions on the Blackwell tensor cores (path 1 of include/agr_conv.h):
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), operands staged in shared memory by TMA, warp-specialised
// (1 TMA producer warp, 1 MMA issuer warp, 4 epilogue warps), mbarrier pipelines.
//
// Replaces the reference's dense contractions, which are cuDNN calls (conv2d_gradfix.py:34,66 via
// dual_styleunet.py:114,275-296 and their autograd backward) followed by separate noise / bias / activation passes
// (dual_styleunet.py:598-604): here noise injection + bias + leaky-ReLU run in the epilogue, straight out of TMEM.
//
// (1) conv_tc_kernel — forward form (forward of every layer, and every data gradient through the adjoint geometry):
//   Y[g*os + phase][co] = sum_{tap} sum_{ci} X[g*is + off(tap)][ci] * Wt[co][tap][ci]
//   M = 128 grid points = one 16 (w) x 8 (h) patch,   N = BN output channels,   K = taps * Cin in 64-blocks.
//   "im2col" never exists in memory: for each tap the A tile is ONE TMA box {64 ch, 16 w, 8 h} of the activation fetched
//   at the tap's shift — with element strides {1,2,2} for the stride-2 layers — TMA zero-fills the out-of-image part
//   (= the conv padding) and writes the 128B-swizzled K-major layout tcgen05.mma reads directly.  B tiles are
//   {64 ci, 1 tap, BN co} boxes of the KRSC weight.  A transposed convolution is stride^2 such sub-convolutions (one per
//   output phase, blockIdx.z) that scatter to every stride-th output pixel: no zero insertion, no wasted MACs.
//
// (2) conv_wgrad_tc_kernel — weight gradient: dW[co][tap][ci] = sum_pixels X[p(tap)][ci] * dY[p][co], a GEMM with
//   M = Cin, N = Cout, K = pixels.  NHWC tensors are channel-contiguous, i.e. MN-major for both operands: a TMA box
//   {64 ch, 16 w, R h} with SWIZZLE_128B lands as 16R pixel rows of 128 B, the canonical UMMA MN-major SW128 layout with
//   SBO = 1024 B (8-pixel groups) and LBO = the distance between consecutive 64-channel boxes.  A CTA owns a pixel
//   slice x (Cin tile, Cout tile) x a GROUP of up to 3 taps that differ only by whole rows: the shifted operand is
//   fetched once with R = 8 + halo rows and each tap's operand is the same box at a start address 16 pixels (2048 B,
//   swizzle-phase preserving) further on.  fp32 tiles are RED-added into dW (split-K over pixels is inherent: K is up
//   to 16 * 512^2), 128 B coalesced per instruction because TMEM lanes = consecutive ci.
#include "conv_common.cuh"

namespace agr {
namespace tc {

constexpr int TILE_W = 16, TILE_H = 8, BM = TILE_W * TILE_H;  // 128 grid points per tile
constexpr int BK = 64;                                       // channels per k-block (128 B of bf16)
constexpr int A_BYTES = BM * BK * 2;                         // 16 KB
constexpr int NUM_THREADS = 192;                             // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue
constexpr int MAX_SMEM = 227 * 1024;

struct ConvParams {
    int N, GH, GW;          // images, compute grid (output-phase coordinates)
    int OH, OW;             // output tensor
    int Cin, Cout;
    int in_stride, out_stride;
    TapList taps;
    const float* bias;      // (Cout) or null
    const float* noise;     // (OH*OW) or null
    const float* noise_w;   // (1) or null
    const float* residual;  // (OH, OW, Cout) fp32 added before bias/activation, shared by the N images; or null
    float* y_f32;           // when set, the raw fp32 accumulator is stored here instead of bf16 in y
    __nv_bfloat16* y;       // (N, OH, OW, Cout)
    int w_cin_offset;       // first input channel of the weight slice this call contracts with (split contraction)
    int activate;           // 0 | 1 lrelu*sqrt2 | 2 lrelu
};

// B_MN: the weight operand is read straight from the KRSC tensor of the layer whose DATA GRADIENT this launch computes
// (w[co][tap][ci]: the contraction index co is the slow one), i.e. as an MN-major B operand: boxes {64 ci, 1 tap, 64 co}
// land as 64 K-rows of 128 B, N atoms of 64 ci are LBO = 8 KB apart, a K step of 16 rows is +2048 B — the layout the
// weight-gradient kernel below uses for both of its operands.  No transposed copy of the weight exists.
template <int BN, int STAGES, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, (STAGES * (A_BYTES + BN * BK * 2) + 1024) * 2 <= MAX_SMEM ? 2 : 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ ConvParams p) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = (p.GW + TILE_W - 1) / TILE_W;
    const int tiles_img = tiles_w * ((p.GH + TILE_H - 1) / TILE_H);
    const int img = blockIdx.x / tiles_img;
    const int tile_m = blockIdx.x - img * tiles_img;
    const int h0 = (tile_m / tiles_w) * TILE_H, w0 = (tile_m % tiles_w) * TILE_W;
    const int n0 = blockIdx.y * BN;
    const int phase = blockIdx.z;
    const int t_begin = p.taps.begin[phase];
    const int kchunks = p.Cin / BK;
    const int num_kb = (p.taps.begin[phase + 1] - t_begin) * kchunks;   // >= 1 (build_taps rejects empty phases)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(&tmem_base_smem);   // BN fp32 columns (power of two >= 32)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ================= TMA producer: the whole warp runs the loop, the copies are predicated on one elected lane =================
        const uint32_t el = elect_one();
        const uint32_t smem_base = smem_u32(smem);
        for (int it = 0; it < num_kb; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait_warp(&empty_bar[s], ph ^ 1);
            const int tap = t_begin + it / kchunks, ck = it % kchunks;
            const uint32_t a_dst = smem_base + s * STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            const uint32_t bar = smem_u32(&full_bar[s]);
            mbar_expect_tx_p(&full_bar[s], STAGE_BYTES, el);
            // OOB -> zeros = padding; with in_stride 2 the box holds every other pixel from its origin
            tma_load_4d_p(a_dst, &map_x, bar, ck * BK, p.in_stride * w0 + p.taps.dx[tap], p.in_stride * h0 + p.taps.dy[tap], img, el);
            if (B_MN) {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                    tma_load_3d_p(b_dst + j * (64 * BK * 2), &map_w, bar, p.w_cin_offset + n0 + 64 * j, p.taps.wt[tap], ck * BK, el);
            } else {
                tma_load_3d_p(b_dst, &map_w, bar, p.w_cin_offset + ck * BK, p.taps.wt[tap], n0, el);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: convergent loop, tcgen05.mma / commit predicated on one elected lane =================
        const uint32_t el = elect_one();
        const uint32_t idesc = umma_idesc2(BM, BN, 0, B_MN ? 1 : 0);
        const uint32_t smem_base = smem_u32(smem);
        for (int it = 0; it < num_kb; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait_warp(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_base + s * STAGE_BYTES, 16);
            const uint32_t b_lo = umma_desc_lo(smem_base + s * STAGE_BYTES + A_BYTES, B_MN ? 64 * BK * 2 : 16);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16: +32 B along a K-major row, +16 rows (2048 B) of an MN-major tile
                umma_f16_p(tmem_base, a_lo + 2 * k, b_lo + (B_MN ? 128 : 2) * k, idesc, (it | k) != 0 ? 1u : 0u, el);
            umma_commit_p(&empty_bar[s], el);      // frees the smem stage once these MMAs have read it
        }
        umma_commit_p(&tmem_full_bar, el);         // accumulator complete
    } else {
        // ================= epilogue: TMEM -> registers -> (+residual, +noise, +bias, lrelu) -> global =================
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;             // grid point within the tile
        const int oy = (h0 + row / TILE_W) * p.out_stride + p.taps.py[phase];
        const int ox = (w0 + row % TILE_W) * p.out_stride + p.taps.px[phase];
        const bool valid = oy < p.OH && ox < p.OW;
        const size_t pix = (size_t)oy * p.OW + ox;   // noise / residual are per output pixel, shared by the N images
        const float add = (valid && p.noise && p.noise_w) ? p.noise_w[0] * p.noise[pix] : 0.f;
        const size_t opix = (size_t)img * p.OH * p.OW + pix;
        __nv_bfloat16* out = p.y + opix * p.Cout + n0;
        const float* res = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;
        float* out32 = p.y_f32 ? p.y_f32 + opix * p.Cout + n0 : nullptr;
        conv_epilogue_row<BN>(tmem_base + ((uint32_t)(q * 32) << 16), valid, p.bias ? p.bias + n0 : nullptr, res, out32, out, add,
                              p.noise != nullptr && p.noise_w != nullptr, p.activate);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---- weight gradient ---------------------------------------------------------------------------------------------------
struct WgradGroup {
    int8_t dx, dy;        // origin shift of the shifted operand's box for this group
    int8_t ntaps;         // 1..3 taps that share the box
    int8_t r[3];          // row offset of each tap inside the box (units of 16-pixel rows)
    int8_t wt[3];         // tap index in dW
};

struct WgradParams {
    int N, GH, GW;          // pixel grid of the plain (unshifted) operand
    int Cin, Cout, taps;    // taps = k*k
    int stride;             // coordinate multiplier of the shifted operand's box origin (= its TMA element stride)
    int x_shifted;          // 1: X is the shifted / strided operand (convolution); 0: dY is (transposed convolution)
    int rows_x, rows_y;     // box heights (8, or 8 + halo for the shifted operand)
    int slices, stages;
    int ci_total, ci_offset;
    int n_groups;
    WgradGroup groups[16];
    float* dw;              // (Cout, taps, ci_total) fp32, accumulated
};

template <int MT, int NT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ WgradParams p) {
    constexpr int ROW_BYTES = TILE_W * BK * 2;               // 2048 B: 16 pixels x 64 channels
    constexpr int MAX_STAGES = 4;
    constexpr int TMEM_COLS = NT == 128 ? 512 : 256;        // 3 taps x NT fp32 columns, power of two
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = (p.GW + TILE_W - 1) / TILE_W, tiles_h = (p.GH + TILE_H - 1) / TILE_H;
    const int boxes_img = tiles_w * tiles_h;
    const long total_boxes = (long)p.N * boxes_img;
    // grid = (tap groups, channel tiles, pixel slices): CTAs that read the SAME pixel boxes are neighbours in launch order, so
    // the operands of a slice come from HBM once and from L2 for the other groups / tiles
    const int box0 = (int)((blockIdx.z * total_boxes) / p.slices), box1 = (int)(((blockIdx.z + 1) * total_boxes) / p.slices);
    const int num_kb = box1 - box0;                        // host keeps slices <= total_boxes
    const int n_tiles = p.Cout / NT;
    const int ci0 = (blockIdx.y / n_tiles) * MT, co0 = (blockIdx.y % n_tiles) * NT;
    const WgradGroup grp = p.groups[blockIdx.x];
    const int a_box = p.rows_x * ROW_BYTES, b_box = p.rows_y * ROW_BYTES;     // one 64-channel box of each operand
    const int a_bytes = (MT / 64) * a_box, stage_bytes = a_bytes + (NT / 64) * b_box;
    const int STAGES = p.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_dy);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // producer: convergent loop, copies predicated on one elected lane (see conv_common.cuh "warp-uniform issue")
        const uint32_t el = elect_one();
        const uint32_t smem_base = smem_u32(smem);
        const int sx = p.x_shifted ? p.stride : 1, sy = p.x_shifted ? 1 : p.stride;
        const int xdx = p.x_shifted ? grp.dx : 0, xdy = p.x_shifted ? grp.dy : 0;
        const int ydx = p.x_shifted ? 0 : grp.dx, ydy = p.x_shifted ? 0 : grp.dy;
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < num_kb; ++it) {
            mbar_wait_warp(&empty_bar[s], ph ^ 1);
            const int b = box0 + it;
            const int img = b / boxes_img, t = b - img * boxes_img;
            const int h0 = (t / tiles_w) * TILE_H, w0 = (t % tiles_w) * TILE_W;
            const uint32_t a_dst = smem_base + s * stage_bytes;
            const uint32_t b_dst = a_dst + a_bytes;
            const uint32_t bar = smem_u32(&full_bar[s]);
            mbar_expect_tx_p(&full_bar[s], stage_bytes, el);
#pragma unroll
            for (int j = 0; j < MT / 64; ++j) tma_load_4d_p(a_dst + j * a_box, &map_x, bar, ci0 + 64 * j, sx * w0 + xdx, sx * h0 + xdy, img, el);
#pragma unroll
            for (int j = 0; j < NT / 64; ++j) tma_load_4d_p(b_dst + j * b_box, &map_dy, bar, co0 + 64 * j, sy * w0 + ydx, sy * h0 + ydy, img, el);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t el = elect_one();
        const uint32_t idesc = umma_idesc(MT, NT, 1);   // a_major = b_major = MN
        const uint32_t smem_base = smem_u32(smem);
        int s = 0; uint32_t ph = 0;
        uint32_t acc = 0;
        for (int it = 0; it < num_kb; ++it) {
            mbar_wait_warp(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_base + s * stage_bytes;
            const uint32_t b_addr = a_addr + a_bytes;
            for (int t = 0; t < grp.ntaps; ++t) {
                const uint32_t a_lo = umma_desc_lo(a_addr + (p.x_shifted ? grp.r[t] * ROW_BYTES : 0), a_box);
                const uint32_t b_lo = umma_desc_lo(b_addr + (p.x_shifted ? 0 : grp.r[t] * ROW_BYTES), b_box);
#pragma unroll
                for (int k = 0; k < BM / 16; ++k)   // 16 pixels per MMA = two 8-row groups = 2048 B = +128 in the (>>4) address field
                    umma_f16_p(tmem_base + t * NT, a_lo + 128 * k, b_lo + 128 * k, idesc, (acc | (uint32_t)k) != 0 ? 1u : 0u, el);
            }
            acc = 1;
            umma_commit_p(&empty_bar[s], el);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_p(&tmem_full_bar, el);
    } else {
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        // accumulator row m (= input channel) sits in TMEM lane m for M = 128, lane (m % 16) + 32 * (m / 16) for M = 64
        const int row = MT == 128 ? q * 32 + lane : q * 16 + lane;
        const bool live = (MT == 128 || lane < 16) && num_kb > 0;
        for (int t = 0; t < grp.ntaps; ++t) {
            float* out = p.dw + ((size_t)co0 * p.taps + grp.wt[t]) * p.ci_total + p.ci_offset + ci0 + row;
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NT + c0), r);
                tmem_wait_ld();
                if (live) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(out + (size_t)(c0 + i) * p.taps * p.ci_total, __uint_as_float(r[i]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
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

// NHWC activation (C, W, H, N): box of 64 channels x bw x bh pixels taken every `es`-th pixel (element strides)
static bool make_map_act(CUtensorMap* m, const void* base, uint64_t C_, uint64_t W_, uint64_t H_, uint64_t N_, uint32_t bw, uint32_t bh, uint32_t es) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {C_, W_, H_, N_};
    cuuint64_t strides[3] = {C_ * 2, C_ * W_ * 2, C_ * W_ * H_ * 2};
    cuuint32_t box[4] = {BK, bw * es, bh * es, 1};
    cuuint32_t estr[4] = {1, es, es, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_map_w(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};  // bytes, dims 1..2
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES, bool B_MN>
static int launch_s(const CUtensorMap& mx, const CUtensorMap& mw, const ConvParams& p, cudaStream_t s) {
    constexpr int smem = STAGES * (A_BYTES + BN * BK * 2) + 1024;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AGR_ERR_CUDA;
        attr = true;
    }
    dim3 grid(p.N * ((p.GH + TILE_H - 1) / TILE_H) * ((p.GW + TILE_W - 1) / TILE_W), p.Cout / BN, p.taps.n_phase);
    conv_tc_kernel<BN, STAGES, B_MN><<<grid, NUM_THREADS, smem, s>>>(mx, mw, p);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

bool forward_supported(const AgrConvGeom& g) {
    if (!geom_ok(g)) return false;
    if (g.Cin % BK || g.Cout % 64) return false;
    TapList t; int is, os, GH, GW;
    return build_taps(g, &t, &is, &os, &GH, &GW);
}

int conv_tc_generation();
bool use_v2(const AgrConvGeom& g);
int launch_forward_v2(const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, bool w_mn, cudaStream_t s);

// w_mn == false: `w` is the K-major operand of this geometry, w[Cout][tap][cin_total].  w_mn == true: `g` is the adjoint of a
// layer and `w` is THAT layer's KRSC weight [g.Cin = its Cout][tap][cin_total = its input channels]: this launch's output
// channels are the weight's fastest index (ep.w_cin_offset / w_cin_total select the slice of them).
int launch_forward(const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, bool w_mn, cudaStream_t s) {
    if (use_v2(g)) return launch_forward_v2(g, x, w, y, ep, w_mn, s);   // conv_tc2.cu: tap groups + CTA pairs
    ConvParams p;
    if (!forward_supported(g) || !build_taps(g, &p.taps, &p.in_stride, &p.out_stride, &p.GH, &p.GW)) return AGR_ERR_INVALID_ARGUMENT;
    const int wc = w_mn ? g.Cout : g.Cin;     // extent of this launch along the weight's fastest (channel) index
    const int cin_total = ep.w_cin_total > 0 ? ep.w_cin_total : wc;
    if (ep.w_cin_offset < 0 || (ep.w_cin_offset % BK) || ep.w_cin_offset + wc > cin_total) return AGR_ERR_INVALID_ARGUMENT;
    p.N = g.N; p.OH = g.OH; p.OW = g.OW; p.Cin = g.Cin; p.Cout = g.Cout;
    p.bias = ep.out_fp32 ? nullptr : ep.bias; p.noise = ep.out_fp32 ? nullptr : ep.noise; p.noise_w = ep.out_fp32 ? nullptr : ep.noise_w;
    p.residual = ep.residual; p.activate = ep.out_fp32 ? 0 : ep.activate; p.w_cin_offset = ep.w_cin_offset;
    p.y = ep.out_fp32 ? nullptr : static_cast<__nv_bfloat16*>(y);
    p.y_f32 = ep.out_fp32 ? static_cast<float*>(y) : nullptr;
    CUtensorMap mx, mw;
    if (!make_map_act(&mx, x, (uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N, TILE_W, TILE_H, (uint32_t)p.in_stride)) return AGR_ERR_CUDA;
    const int BN = (g.Cout % 128 == 0) ? 128 : 64;
    if (w_mn) {   // KRSC weight of the adjoint layer: dims {its Cin = cin_total, taps, its Cout = g.Cin}; boxes {64 ci, 1, 64 co}
        if (!make_map_w(&mw, w, (uint64_t)cin_total, (uint64_t)(g.ksize * g.ksize), (uint64_t)g.Cin, BK, 1, BK)) return AGR_ERR_CUDA;
    } else if (!make_map_w(&mw, w, (uint64_t)cin_total, (uint64_t)(g.ksize * g.ksize), (uint64_t)g.Cout, BK, 1, (uint32_t)BN)) return AGR_ERR_CUDA;
    const long ctas = (long)p.N * ((p.GH + TILE_H - 1) / TILE_H) * ((p.GW + TILE_W - 1) / TILE_W) * (g.Cout / BN) * p.taps.n_phase;
    // 4 stages when the grid is a single wave (1 CTA/SM anyway), 3 stages (96 KB at BN=128) otherwise so that two CTAs
    // per SM overlap one's epilogue with the other's mainloop
    if (w_mn) {
        if (BN == 64) return launch_s<64, 4, true>(mx, mw, p, s);
        if (ctas <= 148) return launch_s<128, 4, true>(mx, mw, p, s);
        return launch_s<128, 3, true>(mx, mw, p, s);
    }
    if (BN == 64) return launch_s<64, 4, false>(mx, mw, p, s);
    if (ctas <= 148) return launch_s<128, 4, false>(mx, mw, p, s);
    return launch_s<128, 3, false>(mx, mw, p, s);
}

static int g_wgrad_ctas = 2 * 148, g_wgrad_min_boxes = 16;   // measured best of {296,592,1184} x {8,16,32}: profiles/r02_conv_generations.txt

bool wgrad_supported(const AgrConvGeom& g) {
    if (!geom_ok(g)) return false;
    return g.Cin % 64 == 0 && g.Cout % 64 == 0;
}

template <int MT, int NT>
static int launch_w(const CUtensorMap& mx, const CUtensorMap& mdy, const WgradParams& p, int smem, cudaStream_t s) {
    static int attr = 0;
    if (attr < smem) {
        if (cudaFuncSetAttribute(conv_wgrad_tc_kernel<MT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AGR_ERR_CUDA;
        attr = smem;
    }
    dim3 grid((unsigned)p.n_groups, (unsigned)((p.Cin / MT) * (p.Cout / NT)), (unsigned)p.slices);
    conv_wgrad_tc_kernel<MT, NT><<<grid, NUM_THREADS, smem, s>>>(mx, mdy, p);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int launch_wgrad(const AgrConvGeom& g, const void* x, const void* dy, float* dw, int ci_total, int ci_offset, cudaStream_t s) {
    if (!wgrad_supported(g)) return AGR_ERR_INVALID_ARGUMENT;
    WgradParams p;
    const int k = g.ksize, st = g.stride;
    p.N = g.N; p.Cin = g.Cin; p.Cout = g.Cout; p.taps = k * k; p.stride = st; p.dw = dw; p.ci_total = ci_total; p.ci_offset = ci_offset;
    p.x_shifted = g.transposed ? 0 : 1;
    // plain operand: dY for a convolution (grid = output), X for a transposed convolution (grid = input); in both the
    // shifted operand is read at  stride * g + (k - pad)
    p.GH = g.transposed ? g.H : g.OH; p.GW = g.transposed ? g.W : g.OW;
    // groups: same kx, ky in one residue class mod stride -> rows (ky - ky_min) / stride apart, at most 3 per group
    int ng = 0, halo = 0;
    for (int kx = 0; kx < k; ++kx)
        for (int c = 0; c < st; ++c) {
            int n_in = 0;
            for (int ky = c; ky < k; ky += st) {
                if (n_in == 0) {
                    if (ng >= 16) return AGR_ERR_INVALID_ARGUMENT;
                    p.groups[ng].dx = (int8_t)(kx - g.pad); p.groups[ng].dy = (int8_t)(ky - g.pad); p.groups[ng].ntaps = 0;
                }
                WgradGroup& G = p.groups[ng];
                G.r[G.ntaps] = (int8_t)n_in; G.wt[G.ntaps] = (int8_t)(ky * k + kx); ++G.ntaps;
                if (n_in > halo) halo = n_in;
                if (++n_in == 3) { ++ng; n_in = 0; }
            }
            if (n_in) ++ng;
        }
    p.n_groups = ng;
    p.rows_x = TILE_H + (p.x_shifted ? halo : 0);
    p.rows_y = TILE_H + (p.x_shifted ? 0 : halo);
    const int MT = (g.Cin % 128 == 0) ? 128 : 64, NT = (g.Cout % 128 == 0) ? 128 : 64;
    const int stage_bytes = ((MT / 64) * p.rows_x + (NT / 64) * p.rows_y) * TILE_W * BK * 2;
    int stages = (MAX_SMEM - 1024) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) return AGR_ERR_INVALID_ARGUMENT;
    p.stages = stages;
    const long boxes = (long)g.N * ((p.GH + TILE_H - 1) / TILE_H) * ((p.GW + TILE_W - 1) / TILE_W);
    const long tiles = (long)(g.Cin / MT) * (g.Cout / NT) * ng;
    // split-K over pixel boxes: never more CTAs than the target (an extra partial wave costs a whole wave: 1 CTA / SM), and
    // enough boxes per CTA that the fp32 reduction epilogue (3 taps x MT x NT REDs per CTA) stays a small part of its work
    long slices = g_wgrad_ctas / tiles;
    if (slices > boxes / g_wgrad_min_boxes) slices = boxes / g_wgrad_min_boxes;
    if (slices < 1) slices = 1;
    p.slices = (int)slices;
    CUtensorMap mx, mdy;
    if (!make_map_act(&mx, x, (uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N, TILE_W, (uint32_t)p.rows_x, p.x_shifted ? (uint32_t)st : 1u)) return AGR_ERR_CUDA;
    if (!make_map_act(&mdy, dy, (uint64_t)g.Cout, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N, TILE_W, (uint32_t)p.rows_y, p.x_shifted ? 1u : (uint32_t)st)) return AGR_ERR_CUDA;
    const int smem = stages * stage_bytes + 1024;
    if (MT == 128 && NT == 128) return launch_w<128, 128>(mx, mdy, p, smem, s);
    if (MT == 128) return launch_w<128, 64>(mx, mdy, p, smem, s);
    if (NT == 128) return launch_w<64, 128>(mx, mdy, p, smem, s);
    return launch_w<64, 64>(mx, mdy, p, smem, s);
}

}  // namespace tc
}  // namespace agr

extern "C" int agr_conv2d_set_wgrad_split(int32_t max_ctas, int32_t min_boxes) {
    if (max_ctas > 0) agr::tc::g_wgrad_ctas = max_ctas;
    if (min_boxes > 0) agr::tc::g_wgrad_min_boxes = min_boxes;
    return agr::tc::g_wgrad_ctas;
}
