// Kernel parameter blocks, launcher prototypes and the sm_100a PTX helpers
// (mbarrier + cp.async.bulk) used by the blend kernels.
#pragma once
#include <cstdio>
#include "raster_common.cuh"

#ifndef AGR_BATCH
#define AGR_BATCH 128  // instance records staged per shared-memory batch (128 * 48 B = 6 KB)
#endif
#ifndef AGR_STAGES
#define AGR_STAGES 4   // ring depth of the producer/consumer pipeline of the blend kernels (24 KB)
#endif

namespace agr {

// ---------------------------------------------------------------- PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "AGR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra AGR_DONE_%=;\n\t"
        "bra AGR_WAIT_%=;\n\t"
        "AGR_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// single probe (the instruction itself blocks for a bounded, implementation-defined time, then reports false)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// bytes must be a multiple of 16; src/dst 16-byte aligned.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    mbar_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- parameter blocks ----
struct PreprocessFwdParams {
    int P, V, W, H;
    int sh_degree, sh_coeffs;
    float scale_modifier;
    int prefiltered;
    uint32_t grid_x, grid_y;
    const float* means3D; const float* scales; const float* rotations; const float* opacities;
    const float* cov3D_precomp; const float* shs; const float* colors_precomp;
    const float* viewmatrix; const float* projmatrix; const float* campos;
    int* radii;
    GeomRec* ws_rec; uint32_t* ws_tiles; float* ws_rgb; uint8_t* ws_clamped;
};

struct DuplicateParams {
    int P, V;
    uint32_t grid_x, grid_y;
    uint64_t capacity;
    const GeomRec* ws_rec; const uint64_t* ws_offsets;
    uint64_t* keys; uint32_t* vals;
};

struct GatherParams {
    uint32_t R; int P; uint32_t tiles_per_view;
    const int64_t* dev_count;  // sync-free mode: R lives here (R above = capacity = launch bound)
    const uint64_t* keys_sorted; const uint32_t* vals_sorted;
    const GeomRec* ws_rec;
    const float* colors; size_t colors_view_stride;
    uint2* ranges; InstRec* stream;
};

struct BlendFwdParams {
    int W, H; uint32_t grid_x, tiles_per_view, num_tiles_total;
    const uint2* ranges; const InstRec* stream;
    const float* background; int bg_view_stride;
    float* out_color; float* out_depth; float* out_alpha;
    uint32_t* n_contrib; uint32_t* tile_last;
};

struct BlendBwdParams {
    int W, H, P; uint32_t grid_x, tiles_per_view, num_tiles_total;
    const uint2* ranges; const InstRec* stream;
    const float* background; int bg_view_stride;
    const float* out_alpha; const uint32_t* n_contrib; const uint32_t* tile_last;
    const float* dL_dcolor; const float* dL_ddepth; const float* dL_dalpha;
    float* acc;  // (V,P,AGR_ACC_STRIDE)
};

struct PreprocessBwdParams {
    int P, V, W, H;
    int sh_degree, sh_coeffs;
    float scale_modifier;
    const float* means3D; const float* scales; const float* rotations;
    const float* cov3D_precomp; const float* shs;
    const float* viewmatrix; const float* projmatrix; const float* campos;
    const int* radii; const uint8_t* ws_clamped;
    const float* acc;
    int colors_per_view;  // dL_dcolors layout: 1 -> (V,P,3), 0 -> (P,3) summed over views
    float* dL_dmeans3D; float* dL_dmeans2D; float* dL_dcolors; float* dL_dopacity;
    float* dL_dcov3D; float* dL_dsh; float* dL_dscales; float* dL_drotations;
};

void launch_preprocess_fwd(const PreprocessFwdParams&, const ViewScalars&, cudaStream_t);
void launch_duplicate(const DuplicateParams&, cudaStream_t);
void launch_ranges_gather(const GatherParams&, cudaStream_t);
void launch_finalize_count(const uint64_t* offsets_last, uint64_t capacity, uint64_t* keys, int64_t* status, cudaStream_t);
void launch_blend_fwd(const BlendFwdParams&, cudaStream_t);
void launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t);
void launch_blend_bwd(const BlendBwdParams&, cudaStream_t);
void launch_preprocess_bwd(const PreprocessBwdParams&, const ViewScalars&, cudaStream_t);

// CUB-backed primitives (raster_binning.cu)
cudaError_t inclusive_scan_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint64_t* out, size_t n, cudaStream_t);
cudaError_t sort_pairs_u64_u32(void* tmp, size_t tmp_bytes, const uint64_t* kin, uint64_t* kout,
                               const uint32_t* vin, uint32_t* vout, size_t n, int end_bit, cudaStream_t);

}  // namespace agr
