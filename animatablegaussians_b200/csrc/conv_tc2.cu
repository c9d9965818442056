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
// B_MN: weight operand read MN-major from the adjoint layer's KRSC tensor (see conv_tc_kernel in conv_tc.cu).
template <int BN, bool PAIR, bool B_MN>
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
        // ================= TMA producer of EACH CTA: convergent loop, copies predicated on one elected lane =================
        const uint32_t el = elect_one();
        const uint32_t smem_base = smem_u32(smem);
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < num_it; ++it) {
            mbar_wait_warp(&empty_bar[s], ph ^ 1);
            const TapGroup grp = p.groups[g_begin + it / kchunks];
            const int ck = it % kchunks;
            const uint32_t a_dst = smem_base + s * STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_SLOT;
            const int x0 = p.in_stride * w0 + grp.dx, y0 = p.in_stride * h0 + grp.dy;
            if (PAIR) {
                // both CTAs' bytes are counted on the leader's barrier; the leader alone arrives on it
                mbar_expect_tx_p(&full_bar[s], 2u * (uint32_t)(a_bytes + grp.ntaps * B_BYTES), rank == 0 ? el : 0u);
                const uint32_t bar = mapa_rank0(smem_u32(&full_bar[s]));
                tma_load_4d_pair_p(a_dst, &map_x, bar, ck * BK, x0, y0, img, el);
                for (int t = 0; t < grp.ntaps; ++t) {
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BNL / 64; ++j)
                            tma_load_3d_pair_p(b_dst + t * B_BYTES + j * (64 * BK * 2), &map_w, bar, p.w_cin_offset + n0 + (int)rank * BNL + 64 * j, grp.wt[t], ck * BK, el);
                    } else {
                        tma_load_3d_pair_p(b_dst + t * B_BYTES, &map_w, bar, p.w_cin_offset + ck * BK, grp.wt[t], n0 + (int)rank * BNL, el);
                    }
                }
            } else {
                const uint32_t bar = smem_u32(&full_bar[s]);
                mbar_expect_tx_p(&full_bar[s], (uint32_t)(a_bytes + grp.ntaps * B_BYTES), el);
                tma_load_4d_p(a_dst, &map_x, bar, ck * BK, x0, y0, img, el);
                for (int t = 0; t < grp.ntaps; ++t) {
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BNL / 64; ++j)
                            tma_load_3d_p(b_dst + t * B_BYTES + j * (64 * BK * 2), &map_w, bar, p.w_cin_offset + n0 + 64 * j, grp.wt[t], ck * BK, el);
                    } else {
                        tma_load_3d_p(b_dst + t * B_BYTES, &map_w, bar, p.w_cin_offset + ck * BK, grp.wt[t], n0, el);
                    }
                }
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (in pair mode of the leader CTA only): convergent loop, predicated issue =================
        if (rank == 0) {
            const uint32_t el = elect_one();
            const uint32_t idesc = umma_idesc2(PAIR ? 2 * BM : BM, BN, 0, B_MN ? 1 : 0);
            const uint32_t smem_base = smem_u32(smem);
            int s = 0; uint32_t ph = 0;
            uint32_t acc = 0;
            for (int it = 0; it < num_it; ++it) {
                mbar_wait_warp(&full_bar[s], ph);
                tc_fence_after();
                const int ntaps = p.groups[g_begin + it / kchunks].ntaps;
                const uint32_t a_addr = smem_base + s * STAGE_BYTES;
                const uint32_t b_addr = a_addr + A_SLOT;
                for (int t = 0; t < ntaps; ++t) {
                    const uint32_t a_lo = umma_desc_lo(a_addr + t * ROW_BYTES, 16), b_lo = umma_desc_lo(b_addr + t * B_BYTES, B_MN ? 64 * BK * 2 : 16);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {  // UMMA_K = 16: +32 B along a K-major row, +16 rows (2048 B) of an MN-major tile
                        if (PAIR) umma_f16_pair_p(tmem_base, a_lo + 2 * k, b_lo + (B_MN ? 128 : 2) * k, idesc, acc, el);
                        else umma_f16_p(tmem_base, a_lo + 2 * k, b_lo + (B_MN ? 128 : 2) * k, idesc, acc, el);
                        acc = 1;
                    }
                }
                if (PAIR) umma_commit_pair_p(&empty_bar[s], el); else umma_commit_p(&empty_bar[s], el);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (PAIR) umma_commit_pair_p(&tmem_full_bar, el); else umma_commit_p(&tmem_full_bar, el);
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
        conv_epilogue_row<BN>(tmem_base + ((uint32_t)(q * 32) << 16), valid, p.bias ? p.bias + n0 : nullptr, res, out32, out, add,
                              p.noise != nullptr && p.noise_w != nullptr, p.activate);
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

template <int BN, bool PAIR, bool B_MN>
static int launch_k(const CUtensorMap& mx, const CUtensorMap& mw, Params& p, long tiles, cudaStream_t s) {
    constexpr int BNL = PAIR ? BN / 2 : BN;
    constexpr int stage_bytes = A_SLOT + 3 * BNL * BK * 2;
    int stages = (MAX_SMEM - 1024) / stage_bytes;
    if (stages > 6) stages = 6;
    // a grid of several waves: two CTAs per SM (one's epilogue under the other's mainloop) when two 2-stage rings fit
    const long ctas = tiles * (p.Cout / BN) * p.n_phase;
    constexpr int half_smem = (MAX_SMEM - 2048) / 2;
    // ... never for CTA pairs: two pairs resident on the same two SMs would interleave their cta_group::2 TMEM allocations
    // (each needs the allocation permit of BOTH SMs), and a pair kernel keeps the whole shared memory for its ring anyway
    if (!PAIR && ctas > 2 * 148 && 2 * stage_bytes + 1024 <= half_smem) stages = (half_smem - 1024) / stage_bytes;
    if (stages < 2) return AGR_ERR_INVALID_ARGUMENT;
    p.stages = stages;
    const int smem = stages * stage_bytes + 1024;
    static int attr = 0;
    if (attr < smem) {
        if (cudaFuncSetAttribute(conv_tc2_kernel<BN, PAIR, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AGR_ERR_CUDA;
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
    if (cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BN, PAIR, B_MN>, mx, mw, p) != cudaSuccess) return AGR_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // namespace v2

// ---- persistent variant for the 64-output-channel layers -------------------------------------------------------------------
// The layers with Cout = 64 at 512^2 x 16 views have 32 768 pixel tiles with a SHORT contraction each (9 taps x 64 or 128
// channels): one tile per CTA is bound by the per-CTA latency chain (launch, barrier init, TMEM allocation, the first TMA
// round trip, the epilogue) and not by any throughput.  Here one CTA per SM walks a strided list of tiles with
//   * the whole weight (taps x Cin x 64, 72 or 144 KB) resident in shared memory, loaded once;
//   * the producer warp running ahead across tile boundaries through a ring of haloed pixel boxes (tap groups, see above);
//   * TWO accumulators in TMEM (2 x 64 columns): the epilogue of tile i overlaps the MMAs of tile i + 1.
// Steady state per tile: 3 boxes x 20 KB from L2, 36 tcgen05.mma (128 x 64 x 16), one 128 x 64 epilogue.
namespace v3 {
using namespace v2;

constexpr int W_TILE = 64 * BK * 2;      // 8 KB: 64 output channels x 64 input channels of one tap
constexpr int MAX_STAGES3 = 8;

struct Params3 {
    Params c;               // geometry, tap groups, epilogue (c.stages = A ring depth)
    int kk;                 // taps of the weight (ksize^2)
    int n_wtiles;           // kk * Cin / 64
    int total_tiles;        // N * tiles per image * phases
};

template <bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ Params3 q) {
    constexpr int BN = 64;
    const Params& p = q.c;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* w_smem = smem;                                  // [wt][ck] tiles of 8 KB
    unsigned char* a_ring = smem + (size_t)q.n_wtiles * W_TILE;    // stages x A_SLOT
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES3], empty_bar[MAX_STAGES3], w_bar, tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = (p.GW + TILE_W - 1) / TILE_W;
    const int tiles_img = tiles_w * ((p.GH + TILE_H - 1) / TILE_H);
    const int kchunks = p.Cin / BK;
    const int STAGES = p.stages;
    const int a_bytes = p.rows * ROW_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&w_bar, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc<2 * BN>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ================= producer: the weight once, then pixel boxes across all of this CTA's tiles =================
        const uint32_t el = elect_one();
        const uint32_t w_base = smem_u32(w_smem), a_base = smem_u32(a_ring);
        mbar_expect_tx_p(&w_bar, (uint32_t)(q.n_wtiles * W_TILE), el);
        for (int wt = 0; wt < q.kk; ++wt)
            for (int ck = 0; ck < kchunks; ++ck) {
                const uint32_t dst = w_base + (uint32_t)((wt * kchunks + ck) * W_TILE);
                if (B_MN) tma_load_3d_p(dst, &map_w, smem_u32(&w_bar), p.w_cin_offset, wt, ck * BK, el);
                else tma_load_3d_p(dst, &map_w, smem_u32(&w_bar), p.w_cin_offset + ck * BK, wt, 0, el);
            }
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x) {
            const int phase = tile % p.n_phase, rest = tile / p.n_phase;
            const int img = rest / tiles_img, tile_m = rest - img * tiles_img;
            const int h0 = (tile_m / tiles_w) * TILE_H, w0 = (tile_m % tiles_w) * TILE_W;
            for (int g = p.gbegin[phase]; g < p.gbegin[phase + 1]; ++g) {
                const int x0 = p.in_stride * w0 + p.groups[g].dx, y0 = p.in_stride * h0 + p.groups[g].dy;
                for (int ck = 0; ck < kchunks; ++ck) {
                    mbar_wait_warp(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx_p(&full_bar[s], (uint32_t)a_bytes, el);
                    tma_load_4d_p(a_base + (uint32_t)(s * A_SLOT), &map_x, smem_u32(&full_bar[s]), ck * BK, x0, y0, img, el);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: convergent loop, predicated issue =================
        const uint32_t el = elect_one();
        const uint32_t idesc = umma_idesc2(BM, BN, 0, B_MN ? 1 : 0);
        mbar_wait_warp(&w_bar, 0);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(w_smem), a_base = smem_u32(a_ring);
        int s = 0; uint32_t ph = 0;
        int tl = 0;
        for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x, ++tl) {
            const int phase = tile % p.n_phase;
            const int buf = tl & 1;
            mbar_wait_warp(&tempty_bar[buf], (uint32_t)(((tl >> 1) & 1) ^ 1));     // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
            uint32_t acc = 0;
            for (int g = p.gbegin[phase]; g < p.gbegin[phase + 1]; ++g) {
                const TapGroup grp = p.groups[g];
                for (int ck = 0; ck < kchunks; ++ck) {
                    mbar_wait_warp(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = a_base + (uint32_t)(s * A_SLOT);
                    for (int t = 0; t < grp.ntaps; ++t) {
                        const uint32_t a_lo = umma_desc_lo(a_addr + t * ROW_BYTES, 16);
                        const uint32_t b_lo = umma_desc_lo(w_addr + (uint32_t)((grp.wt[t] * kchunks + ck) * W_TILE), B_MN ? W_TILE : 16);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_f16_p(d_tmem, a_lo + 2 * k, b_lo + (B_MN ? 128 : 2) * k, idesc, acc, el);
                            acc = 1;
                        }
                    }
                    umma_commit_p(&empty_bar[s], el);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
            umma_commit_p(&tfull_bar[buf], el);
        }
    } else {
        // ================= epilogue warps: drain accumulator (tl & 1) while the next tile's MMAs fill the other =================
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        int tl = 0;
        for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x, ++tl) {
            const int phase = tile % p.n_phase, rest = tile / p.n_phase;
            const int img = rest / tiles_img, tile_m = rest - img * tiles_img;
            const int h0 = (tile_m / tiles_w) * TILE_H, w0 = (tile_m % tiles_w) * TILE_W;
            const int buf = tl & 1;
            mbar_wait_warp(&tfull_bar[buf], (uint32_t)((tl >> 1) & 1));
            tc_fence_after();
            const int oy = (h0 + row / TILE_W) * p.out_stride + p.py[phase];
            const int ox = (w0 + row % TILE_W) * p.out_stride + p.px[phase];
            const bool valid = oy < p.OH && ox < p.OW;
            const size_t pix = (size_t)oy * p.OW + ox;
            const float add = (valid && p.noise && p.noise_w) ? p.noise_w[0] * p.noise[pix] : 0.f;
            const size_t opix = (size_t)img * p.OH * p.OW + pix;
            __nv_bfloat16* out = p.y + opix * p.Cout;
            const float* res = p.residual ? p.residual + pix * p.Cout : nullptr;
            float* out32 = p.y_f32 ? p.y_f32 + opix * p.Cout : nullptr;
            conv_epilogue_row<BN>(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * BN), valid, p.bias, res, out32, out, add,
                                  p.noise != nullptr && p.noise_w != nullptr, p.activate);
            tc_fence_before();                 // this warp's tcgen05.ld of the accumulator are complete (wait::ld above)
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

template <bool B_MN>
static int launch_p(const CUtensorMap& mx, const CUtensorMap& mw, Params3& q, cudaStream_t s) {
    const int w_bytes = q.n_wtiles * W_TILE;
    int stages = (MAX_SMEM - 1024 - w_bytes) / A_SLOT;
    if (stages > MAX_STAGES3) stages = MAX_STAGES3;
    if (stages < 3) return AGR_ERR_INVALID_ARGUMENT;
    q.c.stages = stages;
    const int smem = w_bytes + stages * A_SLOT + 1024;
    static int attr = 0;
    if (attr < smem) {
        if (cudaFuncSetAttribute(conv_tc3_kernel<B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return AGR_ERR_CUDA;
        attr = smem;
    }
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); static int cached = 0; if (!cached) cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev); if (cached > 0) sms = cached; }
    const int grid = q.total_tiles < sms ? q.total_tiles : sms;
    conv_tc3_kernel<B_MN><<<grid, NUM_THREADS, smem, s>>>(mx, mw, q);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // namespace v3

// AGR_CONV_TC / agr_conv2d_set_generation: 1 = conv_tc_kernel always, 2 = tap groups always, 3 = tap groups + CTA pairs always,
// 0 (default) = by shape: measured on B200 (tools/bench_conv.py, profiles/r02_conv_generations.txt) the pair kernel wins where
// the contraction is long and there are enough pixel tiles to fill the SMs with pairs; elsewhere one box per tap with two
// CTAs per SM is as fast or faster (short contractions are bound by per-CTA latency, not by L2 -> shared-memory bytes).
static int g_generation = -1;

int conv_tc_generation() {
    if (g_generation < 0) {
        const char* e = getenv("AGR_CONV_TC");
        g_generation = e ? atoi(e) : 0;
        if (g_generation < 0 || g_generation > 3) g_generation = 0;
    }
    return g_generation;
}

// The persistent 64-channel kernel (AGR_CONV_PERSISTENT=0 switches it off) serves single-phase geometries only: 0.55 vs 0.60 ms
// on the 16 x 512^2 64 -> 64 layer, 40.85 vs 41.3 ms per train step; the transposed 128 -> 64 layer (four output phases of 1-4
// taps) is faster on one tile per CTA (0.52 vs 0.62 ms).  History: with independent per-lane mbarrier polling this kernel hung
// the captured step in 3 of 5 runs (a lane lagging a full ring cycle sees its parity again); with the lockstep wait
// (conv_common.cuh mbar_wait_warp) 0 of 6 — profiles/r02_hang_hunt.txt.
static int g_persistent = -1;

// Persistent kernel: Cout = 64, the whole weight fits next to a >= 3-deep box ring, and there are several waves of tiles.
static bool use_v3(const AgrConvGeom& g) {
    if (g_persistent < 0) { const char* e = getenv("AGR_CONV_PERSISTENT"); g_persistent = e ? atoi(e) : 1; }
    if (!g_persistent || conv_tc_generation() != 0) return false;
    if (g.Cout != 64 || g.Cin % 64 || g.Cin > 128) return false;
    if (g.transposed && g.stride > 1) return false;      // several output phases: see above
    const int s = g.transposed ? g.stride : 1;
    const long tiles = (long)g.N * (((g.OH + s - 1) / s + v2::TILE_H - 1) / v2::TILE_H) * (((g.OW + s - 1) / s + v2::TILE_W - 1) / v2::TILE_W) * s * s;
    return tiles >= 4 * 148 && g.ksize * g.ksize * (g.Cin / 64) * v3::W_TILE + 3 * v2::A_SLOT + 1024 <= v2::MAX_SMEM;
}

bool use_v2(const AgrConvGeom& g) {
    if (use_v3(g)) return true;
    const int gen = conv_tc_generation();
    if (gen) return gen >= 2;
    if (g.stride != 1 || g.ksize != 3) return false;   // (the data gradient of a stride-1 layer arrives as transposed, stride 1)
    const long tiles = (long)g.N * ((g.OH + v2::TILE_H - 1) / v2::TILE_H) * ((g.OW + v2::TILE_W - 1) / v2::TILE_W);
    return g.Cin >= 256 && g.Cout % 128 == 0 && tiles >= 128 && tiles % 2 == 0;
}

int launch_forward_v2(const AgrConvGeom& g, const void* x, const void* w, void* y, const AgrConvEpilogue& ep, bool w_mn, cudaStream_t s) {
    using namespace v2;
    Params p;
    if (!geom_ok(g) || g.Cin % BK || g.Cout % 64 || !build_groups(g, &p)) return AGR_ERR_INVALID_ARGUMENT;
    const int wc = w_mn ? g.Cout : g.Cin;
    const int cin_total = ep.w_cin_total > 0 ? ep.w_cin_total : wc;
    if (ep.w_cin_offset < 0 || (ep.w_cin_offset % BK) || ep.w_cin_offset + wc > cin_total) return AGR_ERR_INVALID_ARGUMENT;
    p.N = g.N; p.OH = g.OH; p.OW = g.OW; p.Cin = g.Cin; p.Cout = g.Cout;
    p.bias = ep.out_fp32 ? nullptr : ep.bias; p.noise = ep.out_fp32 ? nullptr : ep.noise; p.noise_w = ep.out_fp32 ? nullptr : ep.noise_w;
    p.residual = ep.residual; p.activate = ep.out_fp32 ? 0 : ep.activate; p.w_cin_offset = ep.w_cin_offset;
    p.y = ep.out_fp32 ? nullptr : static_cast<__nv_bfloat16*>(y);
    p.y_f32 = ep.out_fp32 ? static_cast<float*>(y) : nullptr;
    const long tiles = (long)p.N * ((p.GH + TILE_H - 1) / TILE_H) * ((p.GW + TILE_W - 1) / TILE_W);
    if (use_v3(g)) {
        v3::Params3 q;
        q.c = p; q.kk = g.ksize * g.ksize; q.n_wtiles = q.kk * (g.Cin / BK); q.total_tiles = (int)(tiles * p.n_phase);
        CUtensorMap mx3, mw3;
        if (!make_map_act(&mx3, x, (uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N, TILE_W, (uint32_t)p.rows, (uint32_t)p.in_stride)) return AGR_ERR_CUDA;
        if (!make_map_w(&mw3, w, (uint64_t)cin_total, (uint64_t)q.kk, (uint64_t)(w_mn ? g.Cin : g.Cout), BK, 1, BK)) return AGR_ERR_CUDA;
        return w_mn ? v3::launch_p<true>(mx3, mw3, q, s) : v3::launch_p<false>(mx3, mw3, q, s);
    }
    // channel tile: as wide as the layer allows while the grid still covers the SMs (a pair mode CTA stages BN/2 weight rows)
    int BN = 64;
    if (g.Cout % 128 == 0) BN = 128;
    // an MN-major weight tile is made of whole 64-channel atoms: a pair needs BN >= 128
    const bool pair = conv_tc_generation() != 2 && (tiles % 2 == 0) && !(w_mn && BN < 128);
    if (pair && g.Cout % 256 == 0 && tiles * (g.Cout / 256) * p.n_phase >= 120) BN = 256;
    CUtensorMap mx, mw;
    if (!make_map_act(&mx, x, (uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N, TILE_W, (uint32_t)p.rows, (uint32_t)p.in_stride)) return AGR_ERR_CUDA;
    const int bnl = pair ? BN / 2 : BN;
    if (w_mn) {
        if (!make_map_w(&mw, w, (uint64_t)cin_total, (uint64_t)(g.ksize * g.ksize), (uint64_t)g.Cin, BK, 1, BK)) return AGR_ERR_CUDA;
    } else if (!make_map_w(&mw, w, (uint64_t)cin_total, (uint64_t)(g.ksize * g.ksize), (uint64_t)g.Cout, BK, 1, (uint32_t)bnl)) return AGR_ERR_CUDA;
    if (w_mn) {
        if (pair) return BN == 256 ? launch_k<256, true, true>(mx, mw, p, tiles, s) : launch_k<128, true, true>(mx, mw, p, tiles, s);
        return BN == 128 ? launch_k<128, false, true>(mx, mw, p, tiles, s) : launch_k<64, false, true>(mx, mw, p, tiles, s);
    }
    if (pair) {
        if (BN == 256) return launch_k<256, true, false>(mx, mw, p, tiles, s);
        if (BN == 128) return launch_k<128, true, false>(mx, mw, p, tiles, s);
        return launch_k<64, true, false>(mx, mw, p, tiles, s);
    }
    if (BN == 128) return launch_k<128, false, false>(mx, mw, p, tiles, s);
    return launch_k<64, false, false>(mx, mw, p, tiles, s);
}

}  // namespace tc
}  // namespace agr

extern "C" int agr_conv2d_set_generation(int32_t generation) {
    if (generation >= 0 && generation <= 3) agr::tc::g_generation = generation;
    return agr::tc::conv_tc_generation();
}
