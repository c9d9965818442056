// Shared device helpers + workspace layout of the B200 rasterizer.
//
// Arithmetic conventions follow the reference kernels so that discrete decisions
// (radius, tile rectangle, alpha < 1/255 rejection, T < 1e-4 termination) agree:
//   RAST/cuda_rasterizer/auxiliary.h:41-97  (ndc2Pix, getRect, transformPoint*)
//   RAST/cuda_rasterizer/forward.cu:74-152  (computeCov2D, computeCov3D)
// GLM (column-major mat3, used by the reference) is replaced by the tiny M3 below; its
// product keeps GLM's expression order (third_party/glm/glm/detail/type_mat3x3.inl:486-518)
// so nvcc contracts the same FMAs.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define AGR_TILE_X 16
#define AGR_TILE_Y 16
#define AGR_TILE_PIX (AGR_TILE_X * AGR_TILE_Y)
#define AGR_MAXV 32

namespace agr {

// ------------------------------------------------------------------ small math ----
struct M3 {  // column-major: c[col][row], like glm::mat3
    float c[3][3];
};

__device__ __forceinline__ M3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2,
                                      float c0, float c1, float c2) {
    M3 r;
    r.c[0][0] = a0; r.c[0][1] = a1; r.c[0][2] = a2;
    r.c[1][0] = b0; r.c[1][1] = b1; r.c[1][2] = b2;
    r.c[2][0] = c0; r.c[2][1] = c1; r.c[2][2] = c2;
    return r;
}

__device__ __forceinline__ M3 m3_mul(const M3& A, const M3& B) {
    M3 R;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            R.c[j][i] = A.c[0][i] * B.c[j][0] + A.c[1][i] * B.c[j][1] + A.c[2][i] * B.c[j][2];
        }
    }
    return R;
}

__device__ __forceinline__ M3 m3_transpose(const M3& A) {
    M3 R;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) R.c[j][i] = A.c[i][j];
    return R;
}

__device__ __forceinline__ M3 m3_scale(const M3& A, float s) {  // s * A
    M3 R;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) R.c[j][i] = A.c[j][i] * s;
    return R;
}

__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* __restrict__ m) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}

__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* __restrict__ m) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
                       m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}

__device__ __forceinline__ float3 xform_vec_4x3_transpose(const float3& p, const float* __restrict__ m) {
    return make_float3(m[0] * p.x + m[1] * p.y + m[2] * p.z,
                       m[4] * p.x + m[5] * p.y + m[6] * p.z,
                       m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

__device__ __forceinline__ float ndc_to_pix(float v, int S) {  // auxiliary.h:41-44 (double on purpose)
    return ((v + 1.0) * S - 1.0) * 0.5;
}

struct TileRect { uint32_t x0, y0, x1, y1; };

// auxiliary.h:46-56
__device__ __forceinline__ TileRect tile_rect(float px, float py, int max_radius, uint32_t gx, uint32_t gy) {
    TileRect r;
    r.x0 = min(gx, (uint32_t)max((int)0, (int)((px - max_radius) / AGR_TILE_X)));
    r.y0 = min(gy, (uint32_t)max((int)0, (int)((py - max_radius) / AGR_TILE_Y)));
    r.x1 = min(gx, (uint32_t)max((int)0, (int)((px + max_radius + AGR_TILE_X - 1) / AGR_TILE_X)));
    r.y1 = min(gy, (uint32_t)max((int)0, (int)((py + max_radius + AGR_TILE_Y - 1) / AGR_TILE_Y)));
    return r;
}

// Sigma = (S R)^T (S R) with the reference's un-normalised quaternion (forward.cu:118-152).
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float4 q,
                                                     float* cov6) {
    M3 S = m3_cols(mod * sx, 0.f, 0.f, 0.f, mod * sy, 0.f, 0.f, 0.f, mod * sz);
    float r = q.x, x = q.y, y = q.z, z = q.w;
    M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 M = m3_mul(S, R);
    M3 Sigma = m3_mul(m3_transpose(M), M);
    cov6[0] = Sigma.c[0][0]; cov6[1] = Sigma.c[0][1]; cov6[2] = Sigma.c[0][2];
    cov6[3] = Sigma.c[1][1]; cov6[4] = Sigma.c[1][2]; cov6[5] = Sigma.c[2][2];
}

// EWA projection (forward.cu:74-113). Returns (cov.x, cov.y, cov.z) incl. the 0.3 low-pass.
// Also hands back T (= W*J) for the backward pass.
__device__ __forceinline__ float3 cov2d_project(const float3& mean, float focal_x, float focal_y,
                                                float tan_fovx, float tan_fovy, const float* cov3D,
                                                const float* __restrict__ view, M3* T_out, float3* t_out,
                                                float* xmul, float* ymul) {
    float3 t = xform_point_4x3(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;
    if (xmul) *xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    if (ymul) *ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;

    M3 J = m3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
                   0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
                   0.f, 0.f, 0.f);
    M3 W = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    M3 T = m3_mul(W, J);
    M3 Vrk = m3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
    if (T_out) *T_out = T;
    if (t_out) *t_out = t;
    return make_float3(cov.c[0][0] + 0.3f, cov.c[0][1], cov.c[1][1] + 0.3f);
}

// Conservative footprint of one projected Gaussian.  A pixel at offset d contributes only if
//   power = -1/2 d^T A d <= 0  and  o * exp(power) >= 1/255   (forward.cu:339-349),  A = [[a,b],[b,c]] the conic,
// i.e. inside the ellipse 1/2 d^T A d <= tau, tau = ln(255 o).  Its bounding box has half-sizes
//   hx = sqrt(2 tau (A^-1)_xx) = sqrt(2 tau c / det),  hy = sqrt(2 tau a / det).
// Inflated by 1e-3 relative + 0.01 px (orders of magnitude above the fp32 rounding of `power`), +inf when the conic
// is not positive definite, negative ("never") when o < 1/255.  Returned as half2 bits, rounded up.
__device__ __forceinline__ uint32_t footprint_half_extent(float a, float b, float c, float o) {
    float hx, hy;
    const float det = a * c - b * b;
    if (o < 1.0f / 255.0f) { hx = -1.f; hy = -1.f; }
    else if (!(det > 0.f) || !(a > 0.f) || !(c > 0.f)) { hx = hy = __int_as_float(0x7f800000); }
    else {
        const float tau2 = 2.0f * logf(255.0f * o) + 1e-4f;
        hx = sqrtf(fmaxf(tau2 * c / det, 0.f)) * 1.001f + 0.01f;
        hy = sqrtf(fmaxf(tau2 * a / det, 0.f)) * 1.001f + 0.01f;
    }
    const __half2 h = __halves2half2(__float2half_ru(hx), __float2half_ru(hy));
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_extent(float bits) {
    const uint32_t u = __float_as_uint(bits);
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

// ------------------------------------------------------------------ SH constants ----
__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                    -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                    0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                    -0.5900435899266435f};

// ------------------------------------------------------------------ data layout ----
// Per-(view, Gaussian) record written by preprocess (32 B, one 2x float4 store):
//   a = (px, py, conic_a, conic_b)   b = (conic_c, opacity, z_view, radius as int bits)
struct __align__(16) GeomRec { float4 a, b; };

// Per-instance record of the depth-sorted, tile-major attribute stream (48 B) that the
// blend kernels stream through shared memory:
//   q0 = (px, py, conic_a, conic_b)  q1 = (conic_c, opacity, r, g)  q2 = (b, z_view, id bits, extent)
// extent = half2 (hx, hy), rounded UP: half-sizes of the axis-aligned box around the ellipse outside of which the
// Gaussian cannot reach alpha >= 1/255 (see footprint_half_extent); lets a warp skip Gaussians that miss its pixels.
struct __align__(16) InstRec { float4 q0, q1, q2; };

// Per-(view, Gaussian) gradient accumulator filled by the blend backward (64 B):
//   [0..2] dL/drgb  [3] dL/dz_view  [4..5] dL/dmean2D  [6..8] dL/dconic(a,b,c)  [9] dL/dopacity
#define AGR_ACC_STRIDE 16

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct GeomWs {  // carved from geom_ws
    GeomRec* rec;          // V*P
    uint32_t* tiles;       // V*P   tiles touched
    uint64_t* offsets;     // V*P   inclusive scan of tiles (64-bit: cannot wrap)
    float* rgb;            // V*P*3 (SH only)
    uint8_t* clamped;      // V*P*3 (SH only)
    void* scan_tmp; size_t scan_tmp_bytes;
    size_t total;
};
struct ImageWs {
    uint2* ranges;         // V*tiles
    uint32_t* n_contrib;   // V*H*W
    uint32_t* tile_last;   // V*tiles: max n_contrib in tile (lets the backward skip the dead tail)
    int64_t* status;       // 2: instances emitted, overflow flag (sync-free mode)
    size_t total;
};
struct BinWs {
    uint64_t* keys_in;     // capacity
    uint64_t* keys_out;    // capacity
    uint32_t* vals_in;     // capacity
    uint32_t* vals_out;    // capacity
    InstRec* stream;       // capacity
    void* sort_tmp; size_t sort_tmp_bytes;
    size_t total;
};

GeomWs carve_geom(void* base, size_t P, size_t V, size_t M);
ImageWs carve_image(void* base, size_t V, size_t W, size_t H);
BinWs carve_binning(void* base, size_t capacity);
size_t scan_temp_bytes(size_t n);
size_t sort_temp_bytes(size_t n);

struct ViewScalars {  // passed by value as kernel parameter
    float tan_fovx[AGR_MAXV];
    float tan_fovy[AGR_MAXV];
};

}  // namespace agr
