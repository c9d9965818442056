// Shared pieces of the convolution kernels (conv_tc.cu: tcgen05 paths, conv_direct.cu: CUDA-core paths, conv_api.cu: the
// C entry points of include/agr_conv.h): PTX wrappers for mbarrier / TMA / tcgen05, and the tap tables both paths use.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_conv.h"
#include "../../include/agr_rasterizer.h"
#include "../../include/agr_styleunet.h"

namespace agr {
namespace tc {

// ---- geometry -> tap tables -------------------------------------------------------------------------------------------
// Forward-form kernels compute, for every output phase (py,px) of an output-stride-`os` grid,
//   y[g*os + (py,px)] = sum_{taps t of the phase} x[g*is + (dy_t,dx_t)] * w[:, wt_t, :]
// which covers convolutions (one phase, is = stride) and transposed convolutions (os = stride, stride^2 phases,
// is = 1; phase (py,px) owns the taps ky = (py+pad) mod s, +s, ... with dy = (py + pad - ky)/s).
struct TapList {
    int n_phase;
    int8_t py[4], px[4];
    int8_t begin[5];               // taps of phase p: [begin[p], begin[p+1])
    int8_t dx[16], dy[16], wt[16];
};

inline bool build_taps(const AgrConvGeom& g, TapList* t, int* in_stride, int* out_stride, int* GH, int* GW) {
    const int k = g.ksize, s = g.stride, pad = g.pad;
    if (k < 1 || k > 4 || s < 1 || s > 2 || pad < 0 || pad > 3) return false;
    int n = 0;
    if (!g.transposed) {
        t->n_phase = 1; t->py[0] = t->px[0] = 0; t->begin[0] = 0;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) { t->dx[n] = (int8_t)(kx - pad); t->dy[n] = (int8_t)(ky - pad); t->wt[n] = (int8_t)(ky * k + kx); ++n; }
        t->begin[1] = (int8_t)n;
        *in_stride = s; *out_stride = 1; *GH = g.OH; *GW = g.OW;
        return true;
    }
    t->n_phase = s * s;
    for (int py = 0; py < s; ++py)
        for (int px = 0; px < s; ++px) {
            const int ph = py * s + px;
            t->py[ph] = (int8_t)py; t->px[ph] = (int8_t)px; t->begin[ph] = (int8_t)n;
            for (int ky = (py + pad) % s; ky < k; ky += s)
                for (int kx = (px + pad) % s; kx < k; kx += s) {
                    t->dy[n] = (int8_t)((py + pad - ky) / s); t->dx[n] = (int8_t)((px + pad - kx) / s); t->wt[n] = (int8_t)(ky * k + kx);
                    ++n;
                }
            if (n == t->begin[ph]) return false;   // a phase without taps (k < stride): not a layer of this path
        }
    t->begin[t->n_phase] = (int8_t)n;
    *in_stride = 1; *out_stride = s; *GH = (g.OH + s - 1) / s; *GW = (g.OW + s - 1) / s;
    return true;
}

inline bool geom_ok(const AgrConvGeom& g) {
    if (g.N < 1 || g.H < 1 || g.W < 1 || g.OH < 1 || g.OW < 1 || g.Cin < 1 || g.Cout < 1) return false;
    if (g.ksize < 1 || g.ksize > 4 || g.stride < 1 || g.stride > 2 || g.pad < 0 || g.pad > 3) return false;
    return true;
}

// the adjoint geometry: the data gradient of `g` is the forward of adjoint(g) applied to dy with the transposed weight
inline AgrConvGeom adjoint(const AgrConvGeom& g) {
    AgrConvGeom a = g;
    a.H = g.OH; a.W = g.OW; a.Cin = g.Cout; a.OH = g.H; a.OW = g.W; a.Cout = g.Cin; a.transposed = !g.transposed;
    return a;
}

#ifdef __CUDACC__
// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// UMMA shared-memory descriptor, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
// K-major operand (rows of 128 B = 64 bf16 along K): SBO = 1024 B between 8-row groups, LBO unused (1).
// MN-major operand (rows of 128 B = 64 bf16 along M/N, one row per K index): canonical layout
//   ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>):
//   SBO = 1024 B between 8-K groups, LBO = byte distance between consecutive 64-wide M/N atoms.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10, a_major @15, b_major @16
// (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t umma_idesc2(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t umma_idesc(int M, int N, int mn_major) { return umma_idesc2(M, N, mn_major, mn_major); }
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS) : "memory");
}
#endif  // __CUDACC__

}  // namespace tc
}  // namespace agr
