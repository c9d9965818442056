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
// Wait of a whole warp on one barrier phase, in LOCKSTEP: every lane polls, the warp leaves together once all lanes have seen
// the phase complete (a warp vote per iteration).  32 lanes polling the same parity INDEPENDENTLY is not safe for barriers
// that go through many phases: a lane that falls a full ring cycle behind its siblings sees the parity it waits for come
// round again and can be left waiting when the kernel ends — the intermittent hang of the first warp-uniform version
// (profiles/r02_hang_hunt.txt).  The vote also keeps the loop exit warp-uniform, so ptxas keeps the code behind it on the
// uniform datapath (a `lane == 0` poll + __syncwarp() made it fall back to vector registers + R2UR per MMA operand).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!__all_sync(0xffffffffu, ok));
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
// ---- warp-uniform issue ---------------------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit / cp.async.bulk.tensor are issued by ONE thread, but the loop around them should be run by the
// WHOLE warp with the asynchronous instruction predicated on an elect.sync flag: inside `if (lane == 0) { ... }` ptxas cannot
// prove the operands warp-uniform, keeps descriptors and loop counters in vector registers and wraps every UTCHMMA / UTMALDG
// in an ELECT + R2UR + BRA.U.ANY "waterfall" (~20 instructions and a dependent chain per MMA: measured 190 cycles per
// 128x64x16 MMA whose tensor time is 32 cycles — profiles/r02_ncu_conv64*.txt).  Convergent code keeps everything in uniform
// registers and the MMAs issue back to back.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
constexpr uint32_t UMMA_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B (bits 32..63)
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ void umma_f16_p(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t el) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\tmov.b64 db, {%2, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(el), "r"(UMMA_DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_f16_pair_p(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t el) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\tmov.b64 db, {%2, %6};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(el), "r"(UMMA_DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint64_t* bar, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(el) : "memory");
}
__device__ __forceinline__ void umma_commit_pair_p(uint64_t* bar, uint32_t el) {   // arrives at this offset in BOTH CTAs of the pair
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}"
                 ::"r"(smem_u32(bar)), "r"(el), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_p(uint64_t* bar, uint32_t bytes, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
                 ::"r"(smem_u32(bar)), "r"(bytes), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_4d_p(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
                 "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_3d_p(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                 "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(el) : "memory");
}
// CTA-pair forms: the data lands in the issuing CTA's shared memory, the bytes are counted on the barrier `bar` (a
// shared::cluster address, normally the leader CTA's)
__device__ __forceinline__ void tma_load_4d_pair_p(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
                 "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair_p(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint32_t el) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                 "@q cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(el) : "memory");
}
// ---- epilogue of the forward-form kernels: one accumulator row (TMEM lane) per thread, BN columns in chunks of 32 ----------
//   y = act(acc + residual + noise + bias) -> bf16, or the raw fp32 accumulator (out32).  Launches without any of that (every
//   data gradient) take the `plain` path: tcgen05.ld, 16 packs, 4 x 16-byte stores per chunk.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

template <int BN>
__device__ __forceinline__ void conv_epilogue_row(uint32_t taddr, bool valid, const float* __restrict__ bias, const float* __restrict__ res,
                                                  float* __restrict__ out32, __nv_bfloat16* __restrict__ out, float add, bool has_noise, int activate) {
    const bool plain = !bias && !res && !out32 && !activate && !has_noise;
    const float gain = activate == 1 ? 1.4142135623730951f : 1.f, neg_slope = activate == 3 ? 0.f : 0.2f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_wait_ld();
        if (!valid) continue;
        uint4 packed[4];
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(packed);
        if (plain) {
#pragma unroll
            for (int i = 0; i < 16; ++i) h2[i] = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        } else {
            if (res) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 rr = __ldg(reinterpret_cast<const float4*>(res + c0) + i);
                    r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) + rr.x);
                    r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + rr.y);
                    r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + rr.z);
                    r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + rr.w);
                }
            }
            if (out32) {   // fp32 partial result (no epilogue math): consumed as `residual` by the second half of a split contraction
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_global_256(out32 + c0 + 8 * i, make_uint4(r[8 * i], r[8 * i + 1], r[8 * i + 2], r[8 * i + 3]),
                                  make_uint4(r[8 * i + 4], r[8 * i + 5], r[8 * i + 6], r[8 * i + 7]));
                continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 b = make_float4(add, add, add, add);
                if (bias) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + i);
                    b.x += bb.x; b.y += bb.y; b.z += bb.z; b.w += bb.w;
                }
                float v0 = __uint_as_float(r[4 * i]) + b.x, v1 = __uint_as_float(r[4 * i + 1]) + b.y;
                float v2 = __uint_as_float(r[4 * i + 2]) + b.z, v3 = __uint_as_float(r[4 * i + 3]) + b.w;
                if (activate) {
                    v0 = (v0 > 0.f ? v0 : neg_slope * v0) * gain; v1 = (v1 > 0.f ? v1 : neg_slope * v1) * gain;
                    v2 = (v2 > 0.f ? v2 : neg_slope * v2) * gain; v3 = (v3 > 0.f ? v3 : neg_slope * v3) * gain;
                }
                h2[2 * i] = __floats2bfloat162_rn(v0, v1);
                h2[2 * i + 1] = __floats2bfloat162_rn(v2, v3);
            }
        }
        // two 256-bit stores (STG.256: whole 32-byte sectors) instead of four 128-bit ones
        st_global_256(out + c0, packed[0], packed[1]);
        st_global_256(out + c0 + 16, packed[2], packed[3]);
    }
}

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
