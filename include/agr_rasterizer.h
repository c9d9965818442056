/*
 * agr_rasterizer.h — C ABI of the B200-native differentiable 3D-Gaussian rasterizer
 * (RGB + depth + alpha, forward + backward, one or many views per call).
 *
 * This is the drop-in boundary for the reference's native module
 *   diff_gaussian_rasterization_depth_alpha._C          (RAST/ext.cpp:15-19)
 * whose three entry points are
 *   rasterize_gaussians           (RAST/rasterize_points.h:18-38,  rasterize_points.cu:35-119)
 *   rasterize_gaussians_backward  (RAST/rasterize_points.h:40-65,  rasterize_points.cu:121-208)
 *   mark_visible                  (RAST/rasterize_points.h:67-68,  rasterize_points.cu:210-229)
 * and, below them, CudaRasterizer::Rasterizer::{forward,backward,markVisible}
 * (RAST/cuda_rasterizer/rasterizer.h:20-91).  RAST = gaussians/diff_gaussian_rasterization_depth_alpha.
 *
 * Conventions
 *   - plain C types only: raw device pointers, ints, floats; no torch / C++ types.
 *   - the CALLER owns every buffer (inputs, outputs, workspaces); the library never
 *     allocates, frees or retains device memory.  The reference grows three byte
 *     buffers through callbacks (rasterize_points.cu:27-33); here the caller sizes
 *     them up front with agr_raster_workspace().
 *   - every call takes the cudaStream_t to run on (passed as void*); the reference
 *     launches on the legacy default stream (forward.cu:398,439).
 *   - return value: 0 = ok, otherwise an AgrStatus; no C++ exception crosses the ABI.
 *     (reference: AT_ERROR / std::runtime_error, rasterize_points.cu:58, rasterizer_impl.cu:245)
 *   - matrices are 16 floats in the reference's memory order (row-major torch tensors
 *     holding the TRANSPOSED matrix, i.e. column-major; gaussian_renderer.py:49-51).
 *   - V views are rendered in one call ("view batch").  The reference renders one view
 *     per call; V = 1 reproduces it exactly.  Gaussian geometry (means3D, scales,
 *     rotations, opacities, cov3D_precomp, shs) is shared by all views of a call;
 *     precomputed colours may be shared (colors_view_stride = 0) or per view.
 */
#ifndef AGR_RASTERIZER_H_
#define AGR_RASTERIZER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGR_MAX_VIEWS 32  /* views per call; larger batches are split by the host mirror */

typedef enum AgrStatus {
    AGR_OK = 0,
    AGR_ERR_INVALID_ARGUMENT = 1,
    AGR_ERR_BINNING_CAPACITY = 2, /* binning workspace too small; *num_rendered holds the need */
    AGR_ERR_CUDA = 3,             /* see agr_last_cuda_error() */
    AGR_ERR_WORKSPACE = 4         /* geom/img/bwd workspace smaller than agr_raster_workspace() says */
} AgrStatus;

/* Byte sizes of the caller-allocated workspaces.
 * capacity = maximum number of (tile, Gaussian) instances over all V views
 * (the reference's num_rendered, rasterizer_impl.cu:281-286). */
typedef struct AgrRasterWorkspace {
    size_t geom_bytes;     /* per-(view,Gaussian) state   (reference GeometryState, rasterizer_impl.h:29-44) */
    size_t image_bytes;    /* per-pixel / per-tile state  (reference ImageState,    rasterizer_impl.h:46-52) */
    size_t binning_bytes;  /* per-instance state + sort   (reference BinningState,  rasterizer_impl.h:54-64) */
    size_t backward_bytes; /* gradient accumulators used by agr_raster_backward only */
} AgrRasterWorkspace;

int agr_raster_workspace(int32_t P, int32_t V, int32_t width, int32_t height, int32_t sh_coeffs,
                         int64_t capacity, AgrRasterWorkspace* out);

/* Arguments of one forward call. Replaces the 19 positional arguments of
 * RasterizeGaussiansCUDA (rasterize_points.h:18-38). NULL pointer == "absent" input,
 * like the empty tensors of the reference (__init__.py:200-210). */
typedef struct AgrRasterForwardArgs {
    int32_t P;              /* Gaussians */
    int32_t V;              /* views in this call (1..AGR_MAX_VIEWS) */
    int32_t width, height;  /* image size, shared by all views */
    int32_t sh_degree;      /* D */
    int32_t sh_coeffs;      /* M = sh.size(1), 0 if no SH */
    float scale_modifier;
    int32_t prefiltered;
    int32_t debug;          /* 1: synchronise + check after every stage (auxiliary.h:166-173) */

    const float* background;     /* (3) device, or (V,3) if bg_view_stride = 3 */
    int32_t bg_view_stride;
    const float* means3D;        /* (P,3) */
    const float* shs;            /* (P,M,3) or NULL */
    const float* colors_precomp; /* (P,3) / (V,P,3) or NULL */
    int64_t colors_view_stride;  /* floats between views: 0 (shared) or P*3 */
    const float* opacities;      /* (P,1) */
    const float* scales;         /* (P,3) or NULL */
    const float* rotations;      /* (P,4) or NULL — used UN-normalised (forward.cu:127) */
    const float* cov3D_precomp;  /* (P,6) or NULL */
    const float* viewmatrix;     /* (V,16) device */
    const float* projmatrix;     /* (V,16) device */
    const float* campos;         /* (V,3)  device (SH only) */
    const float* tan_fovx;       /* (V) HOST floats (the reference passes Python floats) */
    const float* tan_fovy;       /* (V) HOST floats */

    float* out_color;  /* (V,3,H,W) */
    float* out_depth;  /* (V,1,H,W) */
    float* out_alpha;  /* (V,1,H,W) */
    int32_t* radii;    /* (V,P) */

    void* geom_ws;    size_t geom_bytes;
    void* image_ws;   size_t image_bytes;
    void* binning_ws; size_t binning_bytes;
    int64_t capacity;           /* instances the binning workspace was sized for */
    int64_t* num_rendered;      /* HOST out: instances emitted (reference return value, costs one stream sync like
                                 * rasterizer_impl.cu:282).  NULL selects the SYNC-FREE mode: the count stays on the
                                 * device, nothing blocks (CUDA-graph capturable); `capacity` instances are sorted and
                                 * an overflow is reported through device_status instead of AGR_ERR_BINNING_CAPACITY. */
    int64_t* device_status;     /* DEVICE out (2 x int64) or NULL: [0] = instances emitted, [1] = 1 if > capacity */
} AgrRasterForwardArgs;

int agr_raster_forward(const AgrRasterForwardArgs* args, void* cuda_stream);

/* Arguments of one backward call. Replaces RasterizeGaussiansBackwardCUDA
 * (rasterize_points.h:40-65). Workspaces are the ones the matching forward filled. */
typedef struct AgrRasterBackwardArgs {
    int32_t P, V, width, height, sh_degree, sh_coeffs;
    float scale_modifier;
    int32_t debug;

    const float* background; int32_t bg_view_stride;
    const float* means3D;
    const float* shs;
    const float* colors_precomp; int64_t colors_view_stride;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    const float* tan_fovx;   /* HOST */
    const float* tan_fovy;   /* HOST */
    const int32_t* radii;    /* (V,P) from forward */
    const float* out_alpha;  /* (V,1,H,W) from forward (backward.cu:463 uses T_final = 1 - alpha) */

    const float* dL_dout_color;  /* (V,3,H,W) */
    const float* dL_dout_depth;  /* (V,1,H,W) */
    const float* dL_dout_alpha;  /* (V,1,H,W) */

    /* gradient outputs; fully overwritten (no need to zero). Shared inputs get the SUM over views. */
    float* dL_dmeans3D;   /* (P,3) */
    float* dL_dmeans2D;   /* (V,P,3): x,y in NDC-scaled units (backward.cu:490-491,589-590), z = 0 */
    float* dL_dcolors;    /* (P,3) if colors_view_stride == 0 else (V,P,3); NULL with SH */
    float* dL_dopacity;   /* (P,1) */
    float* dL_dcov3D;     /* (P,6): meaningful with cov3D_precomp, else internal scratch; may be NULL */
    float* dL_dsh;        /* (P,M,3) or NULL */
    float* dL_dscales;    /* (P,3) or NULL */
    float* dL_drotations; /* (P,4) or NULL */

    const void* geom_ws;    size_t geom_bytes;
    const void* image_ws;   size_t image_bytes;
    const void* binning_ws; size_t binning_bytes;
    void* backward_ws;      size_t backward_bytes;
    int64_t capacity;
    int64_t num_rendered;   /* R returned by the forward; pass -1 after a sync-free forward */
} AgrRasterBackwardArgs;

int agr_raster_backward(const AgrRasterBackwardArgs* args, void* cuda_stream);

/* Frustum test of checkFrustum (rasterizer_impl.cu:54-66): present[i] = z_view > 0.2. */
int agr_raster_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                            const float* projmatrix, uint8_t* present, void* cuda_stream);

/* Last CUDA error code seen by this library on the calling thread's device (cudaError_t). */
int agr_last_cuda_error(void);
const char* agr_last_cuda_error_string(void);
const char* agr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGR_RASTERIZER_H_ */
