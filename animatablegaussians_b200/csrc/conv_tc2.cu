// Forward-form implicit-GEMM convolution, second generation (include/agr_conv.h path 1; forward of every layer and every
// data gradient through the adjoint geometry).  Same arithmetic and epilogue as conv_tc_kernel (conv_tc.cu); what changed
// is how many bytes travel L2 -> shared memory per FLOP, which is what bounded that kernel: at 128x128x64 per k-block it
// needs 32 KB per 256 MMA-cycles = 128 B/clk/SM against the ~43 B/clk/SM the L2 sustains chip-wide (B300_MICROARCH.md
// "LTS throughput cap").
//
//   * TAP GROUPS.  Taps of one output phase that share dx and whose dy are consecutive multiples of the input stride read
//     the SAME pixel box shifted by whole rows.  One TMA box {64 ch, 16 w, 8 + halo h} per group replaces up to three
//     {64,16,8} boxes; tap t of the group is the same box at a start address t * 16 pixels * 128 B = t * 2048 B further on
//     (a multiple of the 1024-B swizzle atom, so the SWIZZLE_128B phase is preserved and plain descriptors work).
//     3x3 stride-1: 3 boxes of 10 rows instead of 9 of 8 (A traffic / 2.4).
//   * CTA PAIRS (cta_group::2).  Two CTAs of a cluster (adjacent pixel tiles, same channel tile) run ONE tcgen05.mma of
//     M = 256: each CTA stages its own pixel box and HALF of the weight tile (rows rank*BN/2 ...), the leader CTA issues the
//     MMAs, each CTA's TMEM receives its own 128 x BN accumulator.  Weight traffic per CTA halves; BN = 256 becomes possible.
//   Per 3x3 layer and 128 x 128 outputs x 64 input channels: 288 KB (conv_tc_kernel) -> 132 KB (pair, BN 128) / 102 KB (pair, BN 256).
//
// Reference call sites replaced: network/styleunet/conv2d_gradfix.py:34,66 (F.conv2d / F.conv_transpose2d from
// dual_styleunet.py:114,275-296) and their cuDNN backward-data kernels; epilogue = dual_styleunet.py:598-604.
#include <cstdlib>
#include "conv_common.cuh"

namespace agr {
namespace tc {

namespace v2 {

constexpr int TILE_W = 16, TILE_H = 8, BM = TILE_W * TILE_H;
constexpr int BK = 64;
constexpr int ROW_BYTES = TILE_W * BK * 2;                   // 2048 B: one row of 16 pixels x 64 channels
constexpr int MAX_ROWS = TILE_H + 2;
constexpr int A_SLOT = MAX_ROWS * ROW_BYTES;                 // 20 KB
constexpr int NUM_THREADS = 192;                             // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue
constexpr int MAX_SMEM = 227 * 1024;

struct TapGroup {
    int8_t dx, dy;        // shift of the group's box origin (first tap)
    int8_t ntaps;         // 1..3; tap t reads the box from row t on
    int8_t wt[3];         // weight tap index (ky * k + kx) of each tap
};

struct Params {
    int N, GH, GW;          // images, compute grid (output-phase coordinates)
    int OH, OW;
    int Cin, Cout;
    int in_stride, out_stride;
    int n_phase;
    int8_t py[4], px[4];
    int8_t gbegin[5];       // groups of phase p: [gbegin[p], gbegin[p+1])
    TapGroup groups[16];
    int rows;               // box height = 8 + halo
    int stages;
    const float* bias;
    const float* noise;
    const float* noise_w;
    const float* residual;
    float* y_f32;
    __nv_bfloat16* y;
    int w_cin_offset;
    int activate;
};

// ---- PTX for the pair mode ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t saddr) {   // shared::cluster address of the same offset in CTA rank 0
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(saddr));
    return r;
}
// TMA loads of a CTA pair: data lands in the issuing CTA's shared memory, the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS) : "memory");
}

// BN: output channels of the tile (of the pair's tile in pair mode).  PAIR: two CTAs per tcgen05.mma (cluster 2x1x1).
template <int BN, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ Params p) {
    constexpr int BNL = PAIR ? BN / 2 : BN;                  // weight rows this CTA stages
    constexpr int B_BYTES = BNL * BK * 2;
    constexpr int STAGE_BYTES = A_SLOT + 3 * B_BYTES;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr int MAX_STAGES = 6;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // cluster (2,1,1): rank = blockIdx.x & 1
    const int tiles_w = (p.GW + TILE_W - 1) / TILE_W;
    const int tiles_img = tiles_w * ((p.GH + TILE_H - 1) / TILE_H);
    const int img = blockIdx.x / tiles_img;
    const int tile_m = blockIdx.x - img * tiles_img;
    const int h0 = (tile_m / tiles_w) * TILE_H, w0 = (tile_m % tiles_w) * TILE_W;
    const int n0 = blockIdx.y * BN;
    const int phase = blockIdx.z;
    const int g_begin = p.gbegin[phase], n_groups = p.gbegin[phase + 1] - g_begin;
    const int kchunks = p.Cin / BK;
    const int num_it = n_groups * kchunks;                   // >= 1
    const int STAGES = p.stages;
    const int a_bytes = p.rows * ROW_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair<TMEM_COLS>(&tmem_base_smem); else tmem_alloc<TMEM_COLS>(&tmem_base_smem);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();     // the peer's barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ================= TMA producer (one elected lane of EACH CTA) =================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < num_it; ++it) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                const TapGroup grp = p.groups[g_begin + it / kchunks];
                const int ck = it % kchunks;
                unsigned char* a_dst = smem + s * STAGE_BYTES;
                unsigned char* b_dst = a_dst + A_SLOT;
                const int x0 = p.in_stride * w0 + grp.dx, y0 = p.in_stride * h0 + grp.dy;
                if (PAIR) {
                    // both CTAs' bytes are counted on the leader's barrier; the leader alone arrives on it
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)(a_bytes + grp.ntaps * B_BYTES));
                    const uint32_t bar = mapa_rank0(smem_u32(&full_bar[s]));
                    tma_load_4d_pair(a_dst, &map_x, bar, ck * BK, x0, y0, img);
                    for (int t = 0; t < grp.ntaps; ++t)
                        tma_load_3d_pair(b_dst + t * B_BYTES, &map_w, bar, p.w_cin_offset + ck * BK, grp.wt[t], n0 + (int)rank * BNL);
                } else {
                    mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + grp.ntaps * B_BYTES));
                    tma_load_4d(a_dst, &map_x, &full_bar[s], ck * BK, x0, y0, img);
                    for (int t = 0; t < grp.ntaps; ++t)
                        tma_load_3d(b_dst + t * B_BYTES, &map_w, &full_bar[s], p.w_cin_offset + ck * BK, grp.wt[t], n0);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected lane; in pair mode of the leader CTA only) =================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc(PAIR ? 2 * BM : BM, BN, 0);
            int s = 0; uint32_t ph = 0;
            uint32_t acc = 0;
            for (int it = 0; it < num_it; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const int ntaps = p.groups[g_begin + it / kchunks].ntaps;
                const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_SLOT;
                for (int t = 0; t < ntaps; ++t) {
                    const uint64_t adesc = umma_desc(a_addr + t * ROW_BYTES, 16), bdesc = umma_desc(b_addr + t * B_BYTES, 16);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {  // UMMA_K = 16 bf16 = 32 B -> +2 in the (>>4) start-address field
                        if (PAIR) umma_f16_pair(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                        else umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                        acc = 1;
                    }
                }
                if (PAIR) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (PAIR) umma_commit_pair(&tmem_full_bar); else umma_commit(&tmem_full_bar);
        }
    } else {
        // ================= epilogue: TMEM -> registers -> (+residual, +noise, +bias, lrelu) -> global =================
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;             // grid point within the tile
        const int oy = (h0 + row / TILE_W) * p.out_stride + p.py[phase];
        const int ox = (w0 + row % TILE_W) * p.out_stride + p.px[phase];
        const bool valid = oy < p.OH && ox < p.OW;
        const size_t pix = (size_t)oy * p.OW + ox;   // noise / residual are per output pixel, shared by the N images
        const float add = (valid && p.noise && p.noise_w) ? p.noise_w[0] * p.noise[pix] : 0.f;
        const size_t opix = (size_t)img * p.OH * p.OW + pix;
        __nv_bfloat16* out = p.y + opix * p.Cout + n0;
        const float* res = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;
        float* out32 = p.y_f32 ? p.y_f32 + opix * p.Cout + n0 : nullptr;
        const float slope_gain = p.activate == 1 ? 1.4142135623730951f : 1.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_wait_ld();
            if (!valid) continue;
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
            if (out32) {   // fp32 partial result (no epilogue math): consumed as `residual` by the second half
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    reinterpret_cast<float4*>(out32 + c0)[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                                           __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                continue;
            }
            uint4 packed[4];
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(packed);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float v0 = __uint_as_float(r[2 * i]) + add, v1 = __uint_as_float(r[2 * i + 1]) + add;
                if (p.bias) { v0 += p.bias[n0 + c0 + 2 * i]; v1 += p.bias[n0 + c0 + 2 * i + 1]; }
                if (p.activate) {
                    v0 = (v0 > 0.f ? v0 : 0.2f * v0) * slope_gain;
                    v1 = (v1 > 0.f ? v1 : 0.2f * v1) * slope_gain;
                }
                h2[i] = __floats2bfloat162_rn(v0, v1);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = packed[i];
        }
        tc_fence_before();
    }
    __syncthreads();
    if (PAIR) cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the pair's MMAs / multicast arrivals are in flight
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
// Taps of the geometry -> groups (same phase, same dx, dy consecutive multiples of the input stride, at most 3).
static bool build_groups(const AgrConvGeom& g, Params* p) {
    TapList t;
    if (!build_taps(g, &t, &p->in_stride, &p->out_stride, &p->GH, &p->GW)) return false;
    p->n_phase = t.n_phase;
    int ng = 0, halo = 0;
    for (int ph = 0; ph < t.n_phase; ++ph) {
        p->py[ph] = t.py[ph]; p->px[ph] = t.px[ph]; p->gbegin[ph] = (int8_t)ng;
        bool used[16] = {false};
        for (int a = t.begin[ph]; a < t.begin[ph + 1]; ++a) {
            if (used[a]) continue;
            // the run through tap a: smallest dy first
            int first = a;
            for (bool moved = true; moved;) {
                moved = false;
                for (int b = t.begin[ph]; b < t.begin[ph + 1]; ++b)
                    if (!used[b] && t.dx[b] == t.dx[first] && t.dy[b] == t.dy[first] - p->in_stride) { first = b; moved = true; }
            }
            int cur = first;
            while (cur >= 0) {
                if (ng >= 16) return false;
                TapGroup& G = p->groups[ng++];
                G.dx = t.dx[cur]; G.dy = t.dy[cur]; G.ntaps = 0; G.wt[0] = G.wt[1] = G.wt[2] = 0;
                while (cur >= 0 && G.ntaps < 3) {
                    G.wt[G.ntaps++] = t.wt[cur];
                    used[cur] = true;
                    int next = -1;
                    for (int b = t.begin[ph]; b < t.begin[ph + 1]; ++b)
                        if (!used[b] && t.dx[b] == t.dx[cur] && t.dy[b] == t.dy[cur] + p->in_stride) { next = b; break; }
                    cur = next;
                }
                if (G.ntaps - 1 > halo) halo = G.ntaps - 1;
            }
        }
        if (ng == p->gbegin[ph]) return false;
    }
    p->gbegin[t.n_phase] = (int8_t)ng;
    p->rows = TILE_H + halo;
    return true;
}

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
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool PAIR>
static int launch_k(const CUtensorMap& mx, const CUtensorMap& mw, Params& p, long tiles, cudaStream_t s) {
    constexpr int BNL = PAIR ? BN / 2 : BN;
    constexpr int stage_bytes = A_SLOT + 3 * BNL * BK * 2;
    int stages = (MAX_SMEM - 1024) / stage_bytes;
    if (stages > 6) stages = 6;
    // a grid of several waves: two CTAs per SM (one's epilogue under the other's mainloop) when two 2-stage rings fit
    const long ctas = tiles * (p.Cout / BN) * p.n_phase;
    constexpr int half_smem = (MAX_SMEM - 2048) / 2;
    if (ctas > 2 * 148 && 2 * stage_bytes + 1024 <= half_smem) stages = (half_smem - 1024) / stage_bytes;
    if (stages < 2) return AGR_ERR_INVALID_ARGUMENT;
    p.stages = stages;
    const int smem = stages * stage_bytes + 1024;
    static int attr = 0;
    if (attr < smem) {
        if (cudaFuncSetAttribute(conv_tc2_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AGR_ERR_CUDA;
        attr = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)tiles, (unsigned)(p.Cout / BN), (unsigned)p.n_phase);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BN, PAIR>, mx, mw, p) != cudaSuccess) return AGR_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // namespace v2

// AGR_CONV_TC: 1 = conv_tc_kernel (first generation), 2 = tap groups without pairs, 3 (default) = tap groups + CTA pairs
static int g_generation = -1;

int conv_tc_generation() {
    if (g_generation < 0) {
        const char* e = getenv("AGR_CONV_TC");
        g_generation = e ? atoi(e) : 3;
        if (g_generation < 1 || g_generation > 3) g_generation = 3;
    }
    return g_generation;
}

int launch_forward_v2(const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, cudaStream_t s) {
    using namespace v2;
    Params p;
    if (!geom_ok(g) || g.Cin % BK || g.Cout % 64 || !build_groups(g, &p)) return AGR_ERR_INVALID_ARGUMENT;
    const int cin_total = ep.w_cin_total > 0 ? ep.w_cin_total : g.Cin;
    if (ep.w_cin_offset < 0 || (ep.w_cin_offset % BK) || ep.w_cin_offset + g.Cin > cin_total) return AGR_ERR_INVALID_ARGUMENT;
    p.N = g.N; p.OH = g.OH; p.OW = g.OW; p.Cin = g.Cin; p.Cout = g.Cout;
    p.bias = ep.out_fp32 ? nullptr : ep.bias; p.noise = ep.out_fp32 ? nullptr : ep.noise; p.noise_w = ep.out_fp32 ? nullptr : ep.noise_w;
    p.residual = ep.residual; p.activate = ep.out_fp32 ? 0 : ep.activate; p.w_cin_offset = ep.w_cin_offset;
    p.y = ep.out_fp32 ? nullptr : static_cast<__nv_bfloat16*>(y);
    p.y_f32 = ep.out_fp32 ? static_cast<float*>(y) : nullptr;
    const long tiles = (long)p.N * ((p.GH + TILE_H - 1) / TILE_H) * ((p.GW + TILE_W - 1) / TILE_W);
    const bool pair = conv_tc_generation() >= 3 && (tiles % 2 == 0);
    // channel tile: as wide as the layer allows while the grid still covers the SMs (a pair mode CTA stages BN/2 weight rows)
    int BN = 64;
    if (g.Cout % 128 == 0) BN = 128;
    if (pair && g.Cout % 256 == 0 && tiles * (g.Cout / 256) * p.n_phase >= 120) BN = 256;
    CUtensorMap mx, mw;
    if (!make_map_act(&mx, x, (uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N, TILE_W, (uint32_t)p.rows, (uint32_t)p.in_stride)) return AGR_ERR_CUDA;
    const int bnl = pair ? BN / 2 : BN;
    if (!make_map_w(&mw, w, (uint64_t)cin_total, (uint64_t)(g.ksize * g.ksize), (uint64_t)g.Cout, BK, 1, (uint32_t)bnl)) return AGR_ERR_CUDA;
    if (pair) {
        if (BN == 256) return launch_k<256, true>(mx, mw, p, tiles, s);
        if (BN == 128) return launch_k<128, true>(mx, mw, p, tiles, s);
        return launch_k<64, true>(mx, mw, p, tiles, s);
    }
    if (BN == 128) return launch_k<128, false>(mx, mw, p, tiles, s);
    return launch_k<64, false>(mx, mw, p, tiles, s);
}

}  // namespace tc
}  // namespace agr

extern "C" int agr_conv2d_set_generation(int32_t generation) {
    if (generation >= 1 && generation <= 3) agr::tc::g_generation = generation;
    return agr::tc::conv_tc_generation();
}
