// Backward kernels of the B200 rasterizer.
//
// Reference behaviour being reproduced (RAST = gaussians/diff_gaussian_rasterization_depth_alpha):
//   renderCUDA (bwd)     RAST/cuda_rasterizer/backward.cu:415-601
//   computeCov2DCUDA     RAST/cuda_rasterizer/backward.cu:144-274
//   preprocessCUDA (bwd) RAST/cuda_rasterizer/backward.cu:346-412
//   computeCov3D (bwd)   RAST/cuda_rasterizer/backward.cu:278-341
//   computeColorFromSH   RAST/cuda_rasterizer/backward.cu:20-139
// Design differences (results unchanged up to fp32 summation order, which the reference's
// own atomics already leave unspecified):
//   * the blend backward reduces the ten per-(pixel,Gaussian) gradient terms across the warp
//     with a 12-shuffle transposed butterfly and issues ONE red.global per term per warp
//     (10 lanes, one 64-byte accumulator line per Gaussian) instead of 10 atomics per pixel;
//     warps none of whose pixels is touched by the Gaussian skip it with one vote;
//   * the dead tail of each tile list (behind every pixel's last contributor) is never read;
//   * computeCov2DCUDA + preprocessCUDA + computeCov3D are ONE pass, one thread per Gaussian
//     looping over the view batch, so shared-parameter gradients are summed in registers
//     and written once (no atomics, no zero-fill of the outputs).
#include "raster_kernels.cuh"

namespace agr {

// ------------------------------------------------------------------ blend (backward)
template <int BATCH, int STAGES>
__global__ void __launch_bounds__(AGR_TILE_PIX + 32, 4) blend_bwd_kernel(BlendBwdParams p) {
    __shared__ __align__(128) InstRec s_rec[STAGES][BATCH];
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES];
    constexpr uint32_t NCONS = AGR_TILE_PIX / 32;

    const uint32_t tile_lin = blockIdx.x;
    const uint32_t L = p.tile_last[tile_lin];  // entries [0, L) of this tile's list can matter
    if (L == 0) return;
    const uint32_t v = tile_lin / p.tiles_per_view;
    const uint32_t t = tile_lin - v * p.tiles_per_view;
    const uint32_t tile_y = t / p.grid_x, tile_x = t - tile_y * p.grid_x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t px = tile_x * AGR_TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = tile_y * AGR_TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)p.W && py < (uint32_t)p.H;
    const float2 pixf = make_float2((float)px, (float)py);
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix_id = (size_t)p.W * py + px;

    const uint2 range = p.ranges[tile_lin];
    const InstRec* src = p.stream + range.x;
    const int rounds = (int)((L + BATCH - 1) / BATCH);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], NCONS); }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == NCONS) {
        // ================= producer warp: streams the list back to front through the ring =================
        if (lane == 0) {
            for (int i = 0; i < rounds; ++i) {
                const int s = i % STAGES;
                if (i >= STAGES) mbar_wait(&s_empty[s], ((i / STAGES) & 1) ^ 1);
                const int hi = (int)L - i * BATCH, lo = max(0, hi - BATCH);
                bulk_load(&s_rec[s][0], src + lo, (uint32_t)(hi - lo) * sizeof(InstRec), &s_full[s]);
            }
        }
        return;  // consumers wait on every full barrier, so no copy outlives the CTA
    }

    const float T_final = inside ? (1 - p.out_alpha[v * HW + pix_id]) : 0;
    float T = T_final;
    const uint32_t last_contributor = inside ? p.n_contrib[v * HW + pix_id] : 0;

    float accum_rec0 = 0.f, accum_rec1 = 0.f, accum_rec2 = 0.f;
    float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLp_depth = 0.f, dLp_alpha = 0.f;
    float accum_depth_rec = 0.f, accum_alpha_rec = 0.f;
    if (inside) {
        const float* g = p.dL_dcolor + (size_t)v * 3 * HW;
        dLp0 = g[0 * HW + pix_id]; dLp1 = g[1 * HW + pix_id]; dLp2 = g[2 * HW + pix_id];
        dLp_depth = p.dL_ddepth[v * HW + pix_id];
        dLp_alpha = p.dL_dalpha[v * HW + pix_id];
    }
    float last_alpha = 0.f, last_c0 = 0.f, last_c1 = 0.f, last_c2 = 0.f, last_depth = 0.f;
    const float* bg = p.background + (size_t)v * p.bg_view_stride;
    float bg_dot_dpixel = 0;
    bg_dot_dpixel += bg[0] * dLp0;
    bg_dot_dpixel += bg[1] * dLp1;
    bg_dot_dpixel += bg[2] * dLp2;
    const float ddelx_dx = 0.5 * p.W;
    const float ddely_dy = 0.5 * p.H;

    float* acc_view = p.acc + (size_t)v * p.P * AGR_ACC_STRIDE;
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last_contributor);
    const float bx0 = (float)(tile_x * AGR_TILE_X + (warp & 1) * 8), bx1 = bx0 + 7.f;
    const float by0 = (float)(tile_y * AGR_TILE_Y + (warp >> 1) * 4), by1 = by0 + 3.f;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;

    for (int i = 0; i < rounds; ++i) {
        const int buf = i % STAGES;
        const int hi = (int)L - i * BATCH;
        const int lo = max(0, hi - BATCH);
        mbar_wait(&s_full[buf], (i / STAGES) & 1);

        const int n = hi - lo;
        for (int c = ((n - 1) / 32) * 32; c >= 0; c -= 32) {
          // sub-tile culling (see blend_fwd): lane l tests record c+l — inside the warp's reach (k < warp_last) and its
          // footprint box overlapping the warp's 8x4 pixel block; survivors are replayed back to front.
          const int jl = c + (int)lane;
          bool hit = false;
          if (jl < n && (uint32_t)(lo + jl) < warp_last) {
              const float4 t0 = s_rec[buf][jl].q0;
              const float2 ext = unpack_extent(s_rec[buf][jl].q2.w);
              hit = (t0.x + ext.x >= bx0) && (t0.x - ext.x <= bx1) && (t0.y + ext.y >= by0) && (t0.y - ext.y <= by1);
          }
          uint32_t mask = __ballot_sync(0xffffffffu, hit);
          while (mask) {
            const int b = 31 - __clz(mask);
            mask &= ~(1u << b);
            const int j = c + b;
            const uint32_t k = (uint32_t)(lo + j);  // position in the tile list == reference `contributor`
            const float4 q0 = s_rec[buf][j].q0;
            const float4 q1 = s_rec[buf][j].q1;
            const float2 d = make_float2(q0.x - pixf.x, q0.y - pixf.y);
            const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
            bool valid = (k < last_contributor) && !(power > 0.0f);
            const float G = expf(power);
            const float alpha = min(0.99f, q1.y * G);
            valid = valid && !(alpha < 1.0f / 255.0f);
            if (!__any_sync(0xffffffffu, valid)) continue;

            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f, v9 = 0.f;
            const float4 q2 = s_rec[buf][j].q2;
            if (valid) {
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                float dL_dopa = 0.0f;
                {
                    const float c = q1.z;
                    accum_rec0 = last_alpha * last_c0 + (1.f - last_alpha) * accum_rec0;
                    last_c0 = c;
                    dL_dopa += (c - accum_rec0) * dLp0;
                    v0 = dchannel_dcolor * dLp0;
                }
                {
                    const float c = q1.w;
                    accum_rec1 = last_alpha * last_c1 + (1.f - last_alpha) * accum_rec1;
                    last_c1 = c;
                    dL_dopa += (c - accum_rec1) * dLp1;
                    v1 = dchannel_dcolor * dLp1;
                }
                {
                    const float c = q2.x;
                    accum_rec2 = last_alpha * last_c2 + (1.f - last_alpha) * accum_rec2;
                    last_c2 = c;
                    dL_dopa += (c - accum_rec2) * dLp2;
                    v2 = dchannel_dcolor * dLp2;
                }
                const float c_d = q2.y;
                accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                last_depth = c_d;
                dL_dopa += (c_d - accum_depth_rec) * dLp_depth;
                v3 = dchannel_dcolor * dLp_depth;

                accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                dL_dopa += (1 - accum_alpha_rec) * dLp_alpha;

                dL_dopa *= T;
                last_alpha = alpha;
                dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                const float dL_dG = q1.y * dL_dopa;
                const float gdx = G * d.x;
                const float gdy = G * d.y;
                const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
                const float dG_ddely = -gdy * q1.x - gdx * q0.w;
                v4 = dL_dG * dG_ddelx * ddelx_dx;
                v5 = dL_dG * dG_ddely * ddely_dy;
                v6 = -0.5f * gdx * d.x * dL_dG;
                v7 = -0.5f * gdx * d.y * dL_dG;
                v8 = -0.5f * gdy * d.y * dL_dG;
                v9 = G * dL_dopa;
            }
            // transposed butterfly: 8 "A" values (v0..v7) + 2 "B" values (v8,v9), 12 shuffles
            float w0, w1, w2, w3, s2;
            {
                float send, keep;
                send = b4 ? v0 : v4; keep = b4 ? v4 : v0; w0 = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                send = b4 ? v1 : v5; keep = b4 ? v5 : v1; w1 = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                send = b4 ? v2 : v6; keep = b4 ? v6 : v2; w2 = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                send = b4 ? v3 : v7; keep = b4 ? v7 : v3; w3 = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                send = b4 ? v8 : v9; keep = b4 ? v9 : v8; s2 = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            float u0, u1;
            {
                float send, keep;
                send = b3 ? w0 : w2; keep = b3 ? w2 : w0; u0 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                send = b3 ? w1 : w3; keep = b3 ? w3 : w1; u1 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
            }
            float s;
            {
                const float send = b2 ? u0 : u1, keep = b2 ? u1 : u0;
                s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
            }
            float z;
            {
                const float send = b1 ? s : s2, keep = b1 ? s2 : s;
                z = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                z += __shfl_xor_sync(0xffffffffu, z, 1);
            }
            // lane & 3 == 0 -> A slot (b4*4 + b3*2 + b2) ; lane & 15 == 2 -> B slot 8 + b4
            const uint32_t gid = __float_as_uint(q2.z);
            float* dst = acc_view + (size_t)gid * AGR_ACC_STRIDE;
            if ((lane & 3) == 0) {
                const int slot = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
                atomicAdd(dst + slot, z);
            } else if ((lane & 15) == 2) {
                atomicAdd(dst + 8 + (b4 ? 1 : 0), z);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
    }
}

// ------------------------------------------------------------------ SH backward
// backward.cu:20-139; dL_dsh accumulates over the view batch; returns the dL/dmean part.
__device__ __forceinline__ float3 sh_backward(int deg, int M, const float* __restrict__ sh, float3 pos, float3 campos,
                                              const uint8_t* clamped3, float3 dL_dcolor, float* __restrict__ dL_dsh,
                                              bool first) {
    float3 dir_orig = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    float3 dir = make_float3(dir_orig.x / len, dir_orig.y / len, dir_orig.z / len);
    float dRGB[3] = {dL_dcolor.x * (clamped3[0] ? 0 : 1), dL_dcolor.y * (clamped3[1] ? 0 : 1), dL_dcolor.z * (clamped3[2] ? 0 : 1)};
    float x = dir.x, y = dir.y, z = dir.z;
    float coef[16];
    coef[0] = kSH_C0;
    int ncoef = 1;
    if (deg > 0) {
        coef[1] = -kSH_C1 * y; coef[2] = kSH_C1 * z; coef[3] = -kSH_C1 * x;
        ncoef = 4;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            coef[4] = kSH_C2[0] * xy; coef[5] = kSH_C2[1] * yz; coef[6] = kSH_C2[2] * (2.f * zz - xx - yy);
            coef[7] = kSH_C2[3] * xz; coef[8] = kSH_C2[4] * (xx - yy);
            ncoef = 9;
            if (deg > 2) {
                coef[9] = kSH_C3[0] * y * (3.f * xx - yy); coef[10] = kSH_C3[1] * xy * z;
                coef[11] = kSH_C3[2] * y * (4.f * zz - xx - yy);
                coef[12] = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                coef[13] = kSH_C3[4] * x * (4.f * zz - xx - yy); coef[14] = kSH_C3[5] * z * (xx - yy);
                coef[15] = kSH_C3[6] * x * (xx - 3.f * yy);
                ncoef = 16;
            }
        }
    }
    for (int k = 0; k < M; ++k) {
        for (int c = 0; c < 3; ++c) {
            const float val = (k < ncoef) ? coef[k] * dRGB[c] : 0.f;
            if (first) dL_dsh[k * 3 + c] = val; else if (k < ncoef) dL_dsh[k * 3 + c] += val;
        }
    }
    float dL_ddir[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (deg > 0) {
            dx = -kSH_C1 * sh[3 * 3 + c]; dy = -kSH_C1 * sh[1 * 3 + c]; dz = kSH_C1 * sh[2 * 3 + c];
            if (deg > 1) {
                dx += kSH_C2[0] * y * sh[4 * 3 + c] + kSH_C2[2] * 2.f * -x * sh[6 * 3 + c] + kSH_C2[3] * z * sh[7 * 3 + c] + kSH_C2[4] * 2.f * x * sh[8 * 3 + c];
                dy += kSH_C2[0] * x * sh[4 * 3 + c] + kSH_C2[1] * z * sh[5 * 3 + c] + kSH_C2[2] * 2.f * -y * sh[6 * 3 + c] + kSH_C2[4] * 2.f * -y * sh[8 * 3 + c];
                dz += kSH_C2[1] * y * sh[5 * 3 + c] + kSH_C2[2] * 2.f * 2.f * z * sh[6 * 3 + c] + kSH_C2[3] * x * sh[7 * 3 + c];
                if (deg > 2) {
                    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    dx += (kSH_C3[0] * sh[9 * 3 + c] * 3.f * 2.f * xy + kSH_C3[1] * sh[10 * 3 + c] * yz +
                           kSH_C3[2] * sh[11 * 3 + c] * -2.f * xy + kSH_C3[3] * sh[12 * 3 + c] * -3.f * 2.f * xz +
                           kSH_C3[4] * sh[13 * 3 + c] * (-3.f * xx + 4.f * zz - yy) + kSH_C3[5] * sh[14 * 3 + c] * 2.f * xz +
                           kSH_C3[6] * sh[15 * 3 + c] * 3.f * (xx - yy));
                    dy += (kSH_C3[0] * sh[9 * 3 + c] * 3.f * (xx - yy) + kSH_C3[1] * sh[10 * 3 + c] * xz +
                           kSH_C3[2] * sh[11 * 3 + c] * (-3.f * yy + 4.f * zz - xx) + kSH_C3[3] * sh[12 * 3 + c] * -3.f * 2.f * yz +
                           kSH_C3[4] * sh[13 * 3 + c] * -2.f * xy + kSH_C3[5] * sh[14 * 3 + c] * -2.f * yz +
                           kSH_C3[6] * sh[15 * 3 + c] * -3.f * 2.f * xy);
                    dz += (kSH_C3[1] * sh[10 * 3 + c] * xy + kSH_C3[2] * sh[11 * 3 + c] * 4.f * 2.f * yz +
                           kSH_C3[3] * sh[12 * 3 + c] * 3.f * (2.f * zz - xx - yy) + kSH_C3[4] * sh[13 * 3 + c] * 4.f * 2.f * xz +
                           kSH_C3[5] * sh[14 * 3 + c] * (xx - yy));
                }
            }
        }
        dL_ddir[0] += dx * dRGB[c]; dL_ddir[1] += dy * dRGB[c]; dL_ddir[2] += dz * dRGB[c];
    }
    // dnormvdv (auxiliary.h:107-117)
    const float3 vv = dir_orig;
    const float sum2 = vv.x * vv.x + vv.y * vv.y + vv.z * vv.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - vv.x * vv.x) * dL_ddir[0] - vv.y * vv.x * dL_ddir[1] - vv.z * vv.x * dL_ddir[2]) * invsum32;
    r.y = (-vv.x * vv.y * dL_ddir[0] + (sum2 - vv.y * vv.y) * dL_ddir[1] - vv.z * vv.y * dL_ddir[2]) * invsum32;
    r.z = (-vv.x * vv.z * dL_ddir[0] - vv.y * vv.z * dL_ddir[1] + (sum2 - vv.z * vv.z) * dL_ddir[2]) * invsum32;
    return r;
}

// ------------------------------------------------------------------ preprocess (backward), fused
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(PreprocessBwdParams p, ViewScalars vs) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.P) return;

    const float3 mean = make_float3(p.means3D[3 * g], p.means3D[3 * g + 1], p.means3D[3 * g + 2]);
    float cov6[6];
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    float3 scl = make_float3(0.f, 0.f, 0.f);
    if (p.cov3D_precomp != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; ++i) cov6[i] = p.cov3D_precomp[6 * g + i];
    } else {
        q = reinterpret_cast<const float4*>(p.rotations)[g];
        scl = make_float3(p.scales[3 * g], p.scales[3 * g + 1], p.scales[3 * g + 2]);
        cov3d_from_scale_rot(scl.x, scl.y, scl.z, p.scale_modifier, q, cov6);
    }

    float3 dmean_sum = make_float3(0.f, 0.f, 0.f);
    float dcov_sum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dopac_sum = 0.f;
    float3 dcol_sum = make_float3(0.f, 0.f, 0.f);
    bool any_view = false, sh_first = true;

    for (int v = 0; v < p.V; ++v) {
        const size_t vg = (size_t)v * p.P + g;
        float* d2 = p.dL_dmeans2D + 3 * vg;
        const bool vis = p.radii[vg] > 0;
        if (!vis) {
            d2[0] = 0.f; d2[1] = 0.f; d2[2] = 0.f;
            if (p.dL_dcolors && p.colors_per_view) { p.dL_dcolors[3 * vg] = 0.f; p.dL_dcolors[3 * vg + 1] = 0.f; p.dL_dcolors[3 * vg + 2] = 0.f; }
            continue;
        }
        const float4* a4 = reinterpret_cast<const float4*>(p.acc + vg * AGR_ACC_STRIDE);
        const float4 A0 = a4[0], A1 = a4[1], A2 = a4[2];
        // A0 = (dr, dg, db, ddepth)  A1 = (dmx, dmy, dconic_a, dconic_b)  A2 = (dconic_c, dopacity, -, -)
        const float* view = p.viewmatrix + 16 * v;
        const float* proj = p.projmatrix + 16 * v;
        const float tan_fovx = vs.tan_fovx[v], tan_fovy = vs.tan_fovy[v];
        const float h_y = p.H / (2.0f * tan_fovy);
        const float h_x = p.W / (2.0f * tan_fovx);

        // ---- computeCov2DCUDA (backward.cu:144-274)
        const float3 dL_dconic = make_float3(A1.z, A1.w, A2.x);
        M3 T; float3 t; float x_grad_mul, y_grad_mul;
        const float3 c2 = cov2d_project(mean, h_x, h_y, tan_fovx, tan_fovy, cov6, view, &T, &t, &x_grad_mul, &y_grad_mul);
        const float a = c2.x, b = c2.y, c = c2.z;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dcov[6];
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
            dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);
            dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
            dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
            dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
            dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
            dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
            dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) dcov[i] = 0;
        }
        // Vrk[i][j] symmetric
        const float V00 = cov6[0], V01 = cov6[1], V02 = cov6[2], V11 = cov6[3], V12 = cov6[4], V22 = cov6[5];
        const float dL_dT00 = 2 * (T.c[0][0] * V00 + T.c[0][1] * V01 + T.c[0][2] * V02) * dL_da + (T.c[1][0] * V00 + T.c[1][1] * V01 + T.c[1][2] * V02) * dL_db;
        const float dL_dT01 = 2 * (T.c[0][0] * V01 + T.c[0][1] * V11 + T.c[0][2] * V12) * dL_da + (T.c[1][0] * V01 + T.c[1][1] * V11 + T.c[1][2] * V12) * dL_db;
        const float dL_dT02 = 2 * (T.c[0][0] * V02 + T.c[0][1] * V12 + T.c[0][2] * V22) * dL_da + (T.c[1][0] * V02 + T.c[1][1] * V12 + T.c[1][2] * V22) * dL_db;
        const float dL_dT10 = 2 * (T.c[1][0] * V00 + T.c[1][1] * V01 + T.c[1][2] * V02) * dL_dc + (T.c[0][0] * V00 + T.c[0][1] * V01 + T.c[0][2] * V02) * dL_db;
        const float dL_dT11 = 2 * (T.c[1][0] * V01 + T.c[1][1] * V11 + T.c[1][2] * V12) * dL_dc + (T.c[0][0] * V01 + T.c[0][1] * V11 + T.c[0][2] * V12) * dL_db;
        const float dL_dT12 = 2 * (T.c[1][0] * V02 + T.c[1][1] * V12 + T.c[1][2] * V22) * dL_dc + (T.c[0][0] * V02 + T.c[0][1] * V12 + T.c[0][2] * V22) * dL_db;
        // W (glm literal, backward.cu:182-185): W[c][r] = view[4*r... ] i.e. W.c[0] = (view[0],view[4],view[8])
        const float W00 = view[0], W01 = view[4], W02 = view[8];
        const float W10 = view[1], W11 = view[5], W12 = view[9];
        const float W20 = view[2], W21 = view[6], W22 = view[10];
        const float dL_dJ00 = W00 * dL_dT00 + W01 * dL_dT01 + W02 * dL_dT02;
        const float dL_dJ02 = W20 * dL_dT00 + W21 * dL_dT01 + W22 * dL_dT02;
        const float dL_dJ11 = W10 * dL_dT10 + W11 * dL_dT11 + W12 * dL_dT12;
        const float dL_dJ12 = W20 * dL_dT10 + W21 * dL_dT11 + W22 * dL_dT12;
        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
        float3 dmean = xform_vec_4x3_transpose(make_float3(dL_dtx, dL_dty, dL_dtz), view);

        // ---- preprocessCUDA bwd (backward.cu:372-403)
        const float4 m_hom = xform_point_4x4(mean, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        const float gx = A1.x, gy = A1.y;
        float3 dm1;
        dm1.x = (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
        dm1.y = (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
        dm1.z = (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;
        dmean.x += dm1.x; dmean.y += dm1.y; dmean.z += dm1.z;
        const float mul3 = view[2] * mean.x + view[6] * mean.y + view[10] * mean.z + view[14];
        const float dd = A0.w;
        float3 dm2;
        dm2.x = (view[2] - view[3] * mul3) * dd;
        dm2.y = (view[6] - view[7] * mul3) * dd;
        dm2.z = (view[10] - view[11] * mul3) * dd;
        dmean.x += dm2.x; dmean.y += dm2.y; dmean.z += dm2.z;

        if (p.shs != nullptr) {
            const float3 campos = make_float3(p.campos[3 * v], p.campos[3 * v + 1], p.campos[3 * v + 2]);
            const float3 dsh = sh_backward(p.sh_degree, p.sh_coeffs, p.shs + (size_t)g * p.sh_coeffs * 3, mean, campos,
                                           p.ws_clamped + 3 * vg, make_float3(A0.x, A0.y, A0.z),
                                           p.dL_dsh + (size_t)g * p.sh_coeffs * 3, sh_first);
            sh_first = false;
            dmean.x += dsh.x; dmean.y += dsh.y; dmean.z += dsh.z;
        }

        d2[0] = gx; d2[1] = gy; d2[2] = 0.f;
        if (p.dL_dcolors) {
            if (p.colors_per_view) { p.dL_dcolors[3 * vg] = A0.x; p.dL_dcolors[3 * vg + 1] = A0.y; p.dL_dcolors[3 * vg + 2] = A0.z; }
            else { dcol_sum.x += A0.x; dcol_sum.y += A0.y; dcol_sum.z += A0.z; }
        }
        if (!any_view) {
            dmean_sum = dmean;
#pragma unroll
            for (int i = 0; i < 6; ++i) dcov_sum[i] = dcov[i];
            dopac_sum = A2.y;
            any_view = true;
        } else {
            dmean_sum.x += dmean.x; dmean_sum.y += dmean.y; dmean_sum.z += dmean.z;
#pragma unroll
            for (int i = 0; i < 6; ++i) dcov_sum[i] += dcov[i];
            dopac_sum += A2.y;
        }
    }

    p.dL_dmeans3D[3 * g] = dmean_sum.x; p.dL_dmeans3D[3 * g + 1] = dmean_sum.y; p.dL_dmeans3D[3 * g + 2] = dmean_sum.z;
    p.dL_dopacity[g] = dopac_sum;
    if (p.dL_dcolors && !p.colors_per_view) { p.dL_dcolors[3 * g] = dcol_sum.x; p.dL_dcolors[3 * g + 1] = dcol_sum.y; p.dL_dcolors[3 * g + 2] = dcol_sum.z; }
    if (p.dL_dcov3D) {
#pragma unroll
        for (int i = 0; i < 6; ++i) p.dL_dcov3D[6 * g + i] = dcov_sum[i];
    }
    if (p.shs != nullptr && sh_first) {  // no visible view: dL_dsh = 0
        for (int k = 0; k < p.sh_coeffs * 3; ++k) p.dL_dsh[(size_t)g * p.sh_coeffs * 3 + k] = 0.f;
    }

    // ---- computeCov3D bwd (backward.cu:278-341), once on the view-summed dL/dSigma
    if (p.scales != nullptr) {
        float3 dscale = make_float3(0.f, 0.f, 0.f);
        float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
        if (any_view) {
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                           2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                           2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
            const float3 s = make_float3(p.scale_modifier * scl.x, p.scale_modifier * scl.y, p.scale_modifier * scl.z);
            M3 S = m3_cols(s.x, 0.f, 0.f, 0.f, s.y, 0.f, 0.f, 0.f, s.z);
            M3 M = m3_mul(S, R);
            M3 dL_dSigma = m3_cols(dcov_sum[0], 0.5f * dcov_sum[1], 0.5f * dcov_sum[2],
                                   0.5f * dcov_sum[1], dcov_sum[3], 0.5f * dcov_sum[4],
                                   0.5f * dcov_sum[2], 0.5f * dcov_sum[4], dcov_sum[5]);
            M3 dL_dM = m3_mul(m3_scale(M, 2.0f), dL_dSigma);  // 2.0f * M * dL_dSigma == (2M) * dL_dSigma
            M3 Rt = m3_transpose(R);
            M3 dL_dMt = m3_transpose(dL_dM);
            // glm::dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z
            dscale.x = Rt.c[0][0] * dL_dMt.c[0][0] + Rt.c[0][1] * dL_dMt.c[0][1] + Rt.c[0][2] * dL_dMt.c[0][2];
            dscale.y = Rt.c[1][0] * dL_dMt.c[1][0] + Rt.c[1][1] * dL_dMt.c[1][1] + Rt.c[1][2] * dL_dMt.c[1][2];
            dscale.z = Rt.c[2][0] * dL_dMt.c[2][0] + Rt.c[2][1] * dL_dMt.c[2][1] + Rt.c[2][2] * dL_dMt.c[2][2];
#pragma unroll
            for (int i = 0; i < 3; ++i) { dL_dMt.c[0][i] *= s.x; dL_dMt.c[1][i] *= s.y; dL_dMt.c[2][i] *= s.z; }
            drot.x = 2 * z * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * y * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) + 2 * x * (dL_dMt.c[1][2] - dL_dMt.c[2][1]);
            drot.y = 2 * y * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * z * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) + 2 * r * (dL_dMt.c[1][2] - dL_dMt.c[2][1]) - 4 * x * (dL_dMt.c[2][2] + dL_dMt.c[1][1]);
            drot.z = 2 * x * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * r * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) + 2 * z * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * y * (dL_dMt.c[2][2] + dL_dMt.c[0][0]);
            drot.w = 2 * r * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * x * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) + 2 * y * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * z * (dL_dMt.c[1][1] + dL_dMt.c[0][0]);
        }
        p.dL_dscales[3 * g] = dscale.x; p.dL_dscales[3 * g + 1] = dscale.y; p.dL_dscales[3 * g + 2] = dscale.z;
        reinterpret_cast<float4*>(p.dL_drotations)[g] = drot;
    }
}

void launch_blend_bwd(const BlendBwdParams& p, cudaStream_t s) {
    blend_bwd_kernel<AGR_BATCH, AGR_STAGES><<<p.num_tiles_total, AGR_TILE_PIX + 32, 0, s>>>(p);
}
void launch_preprocess_bwd(const PreprocessBwdParams& p, const ViewScalars& vs, cudaStream_t s) {
    preprocess_bwd_kernel<<<(p.P + 255) / 256, 256, 0, s>>>(p, vs);
}

}  // namespace agr
