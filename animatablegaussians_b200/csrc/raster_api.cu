// extern "C" entry points of the rasterizer (include/agr_rasterizer.h).
// Host-side orchestration only: workspace carving, stage ordering, error mapping.
// Replaces CudaRasterizer::Rasterizer::forward/backward (RAST/cuda_rasterizer/rasterizer_impl.cu:197-447)
// and the torch glue RAST/rasterize_points.cu:35-229.
#include <cstring>
#include "../../include/agr_rasterizer.h"
#include "raster_kernels.cuh"

namespace {
thread_local cudaError_t g_last_err = cudaSuccess;

inline bool cuda_fail(cudaError_t e) {
    if (e != cudaSuccess) { g_last_err = e; return true; }
    return false;
}
// debug == true mirrors CHECK_CUDA (auxiliary.h:166-173): synchronise + check after every stage.
inline bool stage_fail(bool debug, cudaStream_t s) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(s);
    return cuda_fail(e);
}
inline int bits_for(uint32_t n) {  // number of bits to represent values in [0, n)
    int b = 0;
    while (((uint64_t)1 << b) < (uint64_t)n) ++b;
    return b < 1 ? 1 : b;
}
}  // namespace

extern "C" {

int agr_last_cuda_error(void) { return (int)g_last_err; }
const char* agr_last_cuda_error_string(void) { return cudaGetErrorString(g_last_err); }
const char* agr_version(void) { return "agr-b200 0.1 (sm_100a)"; }

int agr_raster_workspace(int32_t P, int32_t V, int32_t width, int32_t height, int32_t sh_coeffs,
                         int64_t capacity, AgrRasterWorkspace* out) {
    if (!out || P < 0 || V < 1 || V > AGR_MAX_VIEWS || width < 1 || height < 1 || capacity < 0)
        return AGR_ERR_INVALID_ARGUMENT;
    const size_t Pn = P > 0 ? (size_t)P : 1;
    out->geom_bytes = agr::carve_geom(nullptr, Pn, V, sh_coeffs).total;
    out->image_bytes = agr::carve_image(nullptr, V, width, height).total;
    out->binning_bytes = agr::carve_binning(nullptr, capacity > 0 ? (size_t)capacity : 1).total;
    out->backward_bytes = Pn * (size_t)V * AGR_ACC_STRIDE * sizeof(float) + 256;
    return AGR_OK;
}

int agr_raster_forward(const AgrRasterForwardArgs* a, void* cuda_stream) {
    using namespace agr;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (!a || a->P < 0 || a->V < 1 || a->V > AGR_MAX_VIEWS || a->width < 1 || a->height < 1) return AGR_ERR_INVALID_ARGUMENT;
    if (!a->out_color || !a->out_depth || !a->out_alpha) return AGR_ERR_INVALID_ARGUMENT;
    const bool sync_free = a->num_rendered == nullptr;
    if (sync_free && a->capacity < 1) return AGR_ERR_INVALID_ARGUMENT;
    const int P = a->P, V = a->V, W = a->width, H = a->height;
    const uint32_t gx = (W + AGR_TILE_X - 1) / AGR_TILE_X, gy = (H + AGR_TILE_Y - 1) / AGR_TILE_Y;
    const uint32_t tiles = gx * gy;
    if (!sync_free) *a->num_rendered = 0;
    if (P > 0) {
        if (!a->means3D || !a->opacities || !a->radii || !a->viewmatrix || !a->projmatrix || !a->background ||
            !a->tan_fovx || !a->tan_fovy)
            return AGR_ERR_INVALID_ARGUMENT;
        if ((a->colors_precomp == nullptr) == (a->shs == nullptr)) return AGR_ERR_INVALID_ARGUMENT;
        if (a->shs && (!a->campos || a->sh_coeffs < 1)) return AGR_ERR_INVALID_ARGUMENT;
        const bool has_sr = a->scales && a->rotations;
        if (has_sr == (a->cov3D_precomp != nullptr)) return AGR_ERR_INVALID_ARGUMENT;
    }
    AgrRasterWorkspace need;
    agr_raster_workspace(P, V, W, H, a->shs ? a->sh_coeffs : 0, a->capacity, &need);
    if (a->geom_bytes < need.geom_bytes || a->image_bytes < need.image_bytes || a->binning_bytes < need.binning_bytes)
        return AGR_ERR_WORKSPACE;
    const bool debug = a->debug != 0;

    ImageWs iw = carve_image(a->image_ws, V, W, H);
    if (cuda_fail(cudaMemsetAsync(iw.ranges, 0, (size_t)V * tiles * sizeof(uint2), s))) return AGR_ERR_CUDA;

    int64_t R = 0;
    GeomWs gw{};
    BinWs bw{};
    if (P > 0) {
        gw = carve_geom(a->geom_ws, P, V, a->shs ? a->sh_coeffs : 0);
        ViewScalars vs;
        for (int v = 0; v < V; ++v) { vs.tan_fovx[v] = a->tan_fovx[v]; vs.tan_fovy[v] = a->tan_fovy[v]; }
        PreprocessFwdParams pp{};
        pp.P = P; pp.V = V; pp.W = W; pp.H = H; pp.sh_degree = a->sh_degree; pp.sh_coeffs = a->sh_coeffs;
        pp.scale_modifier = a->scale_modifier; pp.prefiltered = a->prefiltered; pp.grid_x = gx; pp.grid_y = gy;
        pp.means3D = a->means3D; pp.scales = a->scales; pp.rotations = a->rotations; pp.opacities = a->opacities;
        pp.cov3D_precomp = a->cov3D_precomp; pp.shs = a->shs; pp.colors_precomp = a->colors_precomp;
        pp.viewmatrix = a->viewmatrix; pp.projmatrix = a->projmatrix; pp.campos = a->campos;
        pp.radii = a->radii; pp.ws_rec = gw.rec; pp.ws_tiles = gw.tiles; pp.ws_rgb = gw.rgb; pp.ws_clamped = gw.clamped;
        launch_preprocess_fwd(pp, vs, s);
        if (stage_fail(debug, s)) return AGR_ERR_CUDA;

        const size_t n = (size_t)P * V;
        if (cuda_fail(inclusive_scan_u32(gw.scan_tmp, gw.scan_tmp_bytes, gw.tiles, gw.offsets, n, s))) return AGR_ERR_CUDA;
        if (stage_fail(debug, s)) return AGR_ERR_CUDA;

        // number of (tile, Gaussian) instances; the reference does the same blocking read
        // (rasterizer_impl.cu:281-282) to size its binning buffer.
        int64_t* status = a->device_status ? a->device_status : iw.status;
        if (sync_free) {
            if (a->capacity >= ((int64_t)1 << 32)) return AGR_ERR_INVALID_ARGUMENT;
            bw = carve_binning(a->binning_ws, (size_t)a->capacity);
            launch_finalize_count(gw.offsets + n - 1, (uint64_t)a->capacity, bw.keys_in, status, s);
            if (stage_fail(debug, s)) return AGR_ERR_CUDA;
            R = a->capacity;  // launch bound; the real count is read on the device
        } else {
            uint64_t r64 = 0;
            if (cuda_fail(cudaMemcpyAsync(&r64, gw.offsets + n - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, s))) return AGR_ERR_CUDA;
            if (cuda_fail(cudaStreamSynchronize(s))) return AGR_ERR_CUDA;
            R = (int64_t)r64;
            *a->num_rendered = R;
            if (R > a->capacity || R >= ((int64_t)1 << 32)) return AGR_ERR_BINNING_CAPACITY;
        }

        if (R > 0) {
            bw = carve_binning(a->binning_ws, (size_t)a->capacity);
            DuplicateParams dp{};
            dp.P = P; dp.V = V; dp.grid_x = gx; dp.grid_y = gy; dp.capacity = (uint64_t)a->capacity;
            dp.ws_rec = gw.rec; dp.ws_offsets = gw.offsets; dp.keys = bw.keys_in; dp.vals = bw.vals_in;
            launch_duplicate(dp, s);
            if (stage_fail(debug, s)) return AGR_ERR_CUDA;

            const int end_bit = 32 + bits_for(tiles * (uint32_t)V);
            if (cuda_fail(sort_pairs_u64_u32(bw.sort_tmp, bw.sort_tmp_bytes, bw.keys_in, bw.keys_out, bw.vals_in, bw.vals_out,
                                             (size_t)R, end_bit, s)))
                return AGR_ERR_CUDA;
            if (stage_fail(debug, s)) return AGR_ERR_CUDA;

            GatherParams gp{};
            gp.R = (uint32_t)R; gp.P = P; gp.tiles_per_view = tiles; gp.dev_count = sync_free ? status : nullptr;
            gp.keys_sorted = bw.keys_out; gp.vals_sorted = bw.vals_out; gp.ws_rec = gw.rec;
            gp.colors = a->colors_precomp ? a->colors_precomp : gw.rgb;
            gp.colors_view_stride = a->colors_precomp ? (size_t)a->colors_view_stride : (size_t)P * 3;
            gp.ranges = iw.ranges; gp.stream = bw.stream;
            launch_ranges_gather(gp, s);
            if (stage_fail(debug, s)) return AGR_ERR_CUDA;
        }
    }

    BlendFwdParams bp{};
    bp.W = W; bp.H = H; bp.grid_x = gx; bp.tiles_per_view = tiles; bp.num_tiles_total = tiles * V;
    bp.ranges = iw.ranges; bp.stream = bw.stream; bp.background = a->background; bp.bg_view_stride = a->bg_view_stride;
    bp.out_color = a->out_color; bp.out_depth = a->out_depth; bp.out_alpha = a->out_alpha;
    bp.n_contrib = iw.n_contrib; bp.tile_last = iw.tile_last;
    launch_blend_fwd(bp, s);
    if (stage_fail(debug, s)) return AGR_ERR_CUDA;
    return AGR_OK;
}

int agr_raster_backward(const AgrRasterBackwardArgs* a, void* cuda_stream) {
    using namespace agr;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (!a || a->P < 0 || a->V < 1 || a->V > AGR_MAX_VIEWS) return AGR_ERR_INVALID_ARGUMENT;
    const int P = a->P, V = a->V, W = a->width, H = a->height;
    if (P == 0) return AGR_OK;
    if (!a->dL_dmeans3D || !a->dL_dmeans2D || !a->dL_dopacity || !a->backward_ws) return AGR_ERR_INVALID_ARGUMENT;
    if (a->scales && (!a->dL_dscales || !a->dL_drotations)) return AGR_ERR_INVALID_ARGUMENT;
    if (a->shs && !a->dL_dsh) return AGR_ERR_INVALID_ARGUMENT;
    const size_t acc_bytes = (size_t)P * V * AGR_ACC_STRIDE * sizeof(float);
    if (a->backward_bytes < acc_bytes) return AGR_ERR_WORKSPACE;
    const bool debug = a->debug != 0;
    const uint32_t gx = (W + AGR_TILE_X - 1) / AGR_TILE_X, gy = (H + AGR_TILE_Y - 1) / AGR_TILE_Y;
    const uint32_t tiles = gx * gy;

    GeomWs gw = carve_geom(const_cast<void*>(a->geom_ws), P, V, a->shs ? a->sh_coeffs : 0);
    ImageWs iw = carve_image(const_cast<void*>(a->image_ws), V, W, H);
    BinWs bw = carve_binning(const_cast<void*>(a->binning_ws), (size_t)(a->capacity > 0 ? a->capacity : 1));
    float* acc = static_cast<float*>(a->backward_ws);
    if (cuda_fail(cudaMemsetAsync(acc, 0, acc_bytes, s))) return AGR_ERR_CUDA;

    if (a->num_rendered != 0) {
        BlendBwdParams bp{};
        bp.W = W; bp.H = H; bp.P = P; bp.grid_x = gx; bp.tiles_per_view = tiles; bp.num_tiles_total = tiles * V;
        bp.ranges = iw.ranges; bp.stream = bw.stream; bp.background = a->background; bp.bg_view_stride = a->bg_view_stride;
        bp.out_alpha = a->out_alpha; bp.n_contrib = iw.n_contrib; bp.tile_last = iw.tile_last;
        bp.dL_dcolor = a->dL_dout_color; bp.dL_ddepth = a->dL_dout_depth; bp.dL_dalpha = a->dL_dout_alpha;
        bp.acc = acc;
        launch_blend_bwd(bp, s);
        if (stage_fail(debug, s)) return AGR_ERR_CUDA;
    }

    ViewScalars vs;
    for (int v = 0; v < V; ++v) { vs.tan_fovx[v] = a->tan_fovx[v]; vs.tan_fovy[v] = a->tan_fovy[v]; }
    PreprocessBwdParams pp{};
    pp.P = P; pp.V = V; pp.W = W; pp.H = H; pp.sh_degree = a->sh_degree; pp.sh_coeffs = a->sh_coeffs;
    pp.scale_modifier = a->scale_modifier;
    pp.means3D = a->means3D; pp.scales = a->scales; pp.rotations = a->rotations; pp.cov3D_precomp = a->cov3D_precomp;
    pp.shs = a->shs; pp.viewmatrix = a->viewmatrix; pp.projmatrix = a->projmatrix; pp.campos = a->campos;
    pp.radii = a->radii; pp.ws_clamped = gw.clamped; pp.acc = acc;
    pp.colors_per_view = a->colors_view_stride != 0 ? 1 : 0;
    pp.dL_dmeans3D = a->dL_dmeans3D; pp.dL_dmeans2D = a->dL_dmeans2D; pp.dL_dcolors = a->dL_dcolors;
    pp.dL_dopacity = a->dL_dopacity; pp.dL_dcov3D = a->dL_dcov3D; pp.dL_dsh = a->dL_dsh;
    pp.dL_dscales = a->dL_dscales; pp.dL_drotations = a->dL_drotations;
    launch_preprocess_bwd(pp, vs, s);
    if (stage_fail(debug, s)) return AGR_ERR_CUDA;
    return AGR_OK;
}

int agr_raster_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                            uint8_t* present, void* cuda_stream) {
    (void)projmatrix;  // the reference computes p_proj but only tests z_view (auxiliary.h:149-154)
    if (P < 0) return AGR_ERR_INVALID_ARGUMENT;
    if (P == 0) return AGR_OK;
    if (!means3D || !viewmatrix || !present) return AGR_ERR_INVALID_ARGUMENT;
    agr::launch_mark_visible(P, means3D, viewmatrix, present, static_cast<cudaStream_t>(cuda_stream));
    if (cuda_fail(cudaGetLastError())) return AGR_ERR_CUDA;
    return AGR_OK;
}

}  // extern "C"
