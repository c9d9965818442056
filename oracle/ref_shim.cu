// TEST INFRASTRUCTURE — not product code.
//
// Thin C-ABI shim around the UNMODIFIED reference rasterizer
// (gaussians/diff_gaussian_rasterization_depth_alpha/cuda_rasterizer/rasterizer.h:20-91,
// CudaRasterizer::Rasterizer::{forward,backward,markVisible}).  It is compiled by
// oracle/build_ref.py together with the reference's own forward.cu / backward.cu /
// rasterizer_impl.cu *from where they lie under /root/reference* into
// oracle/_ref/libref_rasterizer.so.  No reference source is copied into this repo.
//
// The shim replaces rasterize_points.cu (the torch glue, rasterize_points.cu:35-208):
// the three growable byte buffers the reference obtains through std::function
// callbacks are grow-only cudaMalloc blocks kept in a handle, so that the reference
// arm of bench.py does not pay an allocation per call (the real glue uses the torch
// caching allocator, which is equally cheap in steady state).
#include <cstdint>
#include <cstdio>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer.h"

namespace {
struct Block {
    char* ptr = nullptr;
    size_t cap = 0;
    char* obtain(size_t n) {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            size_t want = n + n / 4 + 256;
            if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; return nullptr; }
            cap = want;
        }
        return ptr;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};
struct Handle { Block geom, binning, img; };
}  // namespace

extern "C" {

void* refrast_create() { return new Handle(); }

void refrast_destroy(void* h) {
    Handle* hd = static_cast<Handle*>(h);
    hd->geom.release(); hd->binning.release(); hd->img.release();
    delete hd;
}

// Mirrors RasterizeGaussiansCUDA (rasterize_points.cu:35-119). Outputs must be
// zero-filled by the caller (the reference glue uses torch::full(0)).
int refrast_forward(void* h, int P, int D, int M,
                    const float* background, int W, int H,
                    const float* means3D, const float* shs, const float* colors_precomp,
                    const float* opacities, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp,
                    const float* viewmatrix, const float* projmatrix, const float* campos,
                    float tan_fovx, float tan_fovy, int prefiltered,
                    float* out_color, float* out_depth, float* out_alpha, int* radii)
{
    Handle* hd = static_cast<Handle*>(h);
    if (P == 0) return 0;
    std::function<char*(size_t)> g = [hd](size_t n) { return hd->geom.obtain(n); };
    std::function<char*(size_t)> b = [hd](size_t n) { return hd->binning.obtain(n); };
    std::function<char*(size_t)> i = [hd](size_t n) { return hd->img.obtain(n); };
    return CudaRasterizer::Rasterizer::forward(
        g, b, i, P, D, M, background, W, H, means3D, shs, colors_precomp, opacities,
        scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
        tan_fovx, tan_fovy, prefiltered != 0, out_color, out_depth, out_alpha, radii, false);
}

// Mirrors RasterizeGaussiansBackwardCUDA (rasterize_points.cu:121-208). All gradient
// outputs must be zero-filled by the caller (torch::zeros in the reference glue).
void refrast_backward(void* h, int P, int D, int M, int R,
                      const float* background, int W, int H,
                      const float* means3D, const float* shs, const float* colors_precomp,
                      const float* alphas, const float* scales, float scale_modifier,
                      const float* rotations, const float* cov3D_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* campos,
                      float tan_fovx, float tan_fovy, const int* radii,
                      const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas,
                      float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                      float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                      float* dL_dscale, float* dL_drot)
{
    Handle* hd = static_cast<Handle*>(h);
    if (P == 0) return;
    CudaRasterizer::Rasterizer::backward(
        P, D, M, R, background, W, H, means3D, shs, colors_precomp, alphas, scales,
        scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
        tan_fovx, tan_fovy, radii, hd->geom.ptr, hd->binning.ptr, hd->img.ptr,
        dL_dpix, dL_dpix_depth, dL_dalphas, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
        dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, false);
}

// Mirrors markVisible (rasterize_points.cu:210-229).
void refrast_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    if (P == 0) return;
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

int refrast_last_error() { return (int)cudaGetLastError(); }

}  // extern "C"
