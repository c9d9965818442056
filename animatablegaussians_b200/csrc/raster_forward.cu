// Forward kernels of the B200 rasterizer: preprocess -> duplicate -> (sort) -> ranges+gather -> blend.
//
// Reference behaviour being reproduced (RAST = gaussians/diff_gaussian_rasterization_depth_alpha):
//   preprocessCUDA      RAST/cuda_rasterizer/forward.cu:155-256
//   duplicateWithKeys   RAST/cuda_rasterizer/rasterizer_impl.cu:70-111
//   identifyTileRanges  RAST/cuda_rasterizer/rasterizer_impl.cu:116-138
//   renderCUDA          RAST/cuda_rasterizer/forward.cu:261-381
// Design differences (B200-first, results unchanged):
//   * a view batch: one thread per Gaussian walks all V views, so xyz/scale/quat are read
//     once and the view-independent 3-D covariance is built once per step, not per view;
//   * per-(view,Gaussian) state is one 32-byte record (two 128-bit stores) instead of six arrays;
//   * after the sort, attributes are gathered ONCE into a tile-major, depth-sorted 48-byte
//     stream; both blend passes then read it with contiguous bulk copies instead of
//     re-gathering id -> xy/conic/rgb/depth per batch (forward.cu:321-324,359,361).
#include "raster_kernels.cuh"

namespace agr {

// --------------------------------------------------------------------------- SH -> RGB
// forward.cu:20-71. `sh` points at this Gaussian's (M,3) coefficients.
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* __restrict__ sh, float3 pos, float3 campos,
                                            uint8_t* clamped3) {
    float3 dir = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float result = kSH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            float x = dir.x, y = dir.y, z = dir.z;
            result = result - kSH_C1 * y * sh[1 * 3 + c] + kSH_C1 * z * sh[2 * 3 + c] - kSH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z;
                float xy = x * y, yz = y * z, xz = x * z;
                result = result + kSH_C2[0] * xy * sh[4 * 3 + c] + kSH_C2[1] * yz * sh[5 * 3 + c] +
                         kSH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + kSH_C2[3] * xz * sh[7 * 3 + c] +
                         kSH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    result = result + kSH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] +
                             kSH_C3[1] * xy * z * sh[10 * 3 + c] +
                             kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                             kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                             kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
                             kSH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] +
                             kSH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
                }
            }
        }
        result += 0.5f;
        clamped3[c] = (result < 0);
        res[c] = fmaxf(result, 0.0f);
    }
    return make_float3(res[0], res[1], res[2]);
}

// --------------------------------------------------------------------------- preprocess
__global__ void __launch_bounds__(256) preprocess_fwd_kernel(PreprocessFwdParams p, ViewScalars vs) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.P) return;

    const float3 p_orig = make_float3(p.means3D[3 * g], p.means3D[3 * g + 1], p.means3D[3 * g + 2]);
    float cov6[6];
    if (p.cov3D_precomp != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; ++i) cov6[i] = p.cov3D_precomp[6 * g + i];
    } else {
        const float4 q = reinterpret_cast<const float4*>(p.rotations)[g];
        cov3d_from_scale_rot(p.scales[3 * g], p.scales[3 * g + 1], p.scales[3 * g + 2], p.scale_modifier, q, cov6);
    }
    const float opacity = p.opacities[g];

    for (int v = 0; v < p.V; ++v) {
        const size_t vg = (size_t)v * p.P + g;
        const float* view = p.viewmatrix + 16 * v;
        const float* proj = p.projmatrix + 16 * v;
        const float tan_fovx = vs.tan_fovx[v], tan_fovy = vs.tan_fovy[v];
        const float focal_y = p.H / (2.0f * tan_fovy);
        const float focal_x = p.W / (2.0f * tan_fovx);

        int radius_out = 0;
        uint32_t tiles_out = 0;
        GeomRec rec;
        rec.a = make_float4(0.f, 0.f, 0.f, 0.f);
        rec.b = make_float4(0.f, 0.f, 0.f, 0.f);

        // near culling (auxiliary.h:139-164): only z_view <= 0.2 rejects
        const float3 p_view = xform_point_4x3(p_orig, view);
        bool alive = p_view.z > 0.2f;
        if (!alive && p.prefiltered) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }
        if (alive) {
            const float4 p_hom = xform_point_4x4(p_orig, proj);
            const float p_w = 1.0f / (p_hom.w + 0.0000001f);
            const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);
            const float3 cov = cov2d_project(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov6, view,
                                             nullptr, nullptr, nullptr, nullptr);
            const float det = (cov.x * cov.z - cov.y * cov.y);
            if (det != 0.0f) {
                const float det_inv = 1.f / det;
                const float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);
                const float mid = 0.5f * (cov.x + cov.z);
                const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
                const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
                const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
                const float px = ndc_to_pix(p_proj.x, p.W);
                const float py = ndc_to_pix(p_proj.y, p.H);
                const TileRect r = tile_rect(px, py, (int)my_radius, p.grid_x, p.grid_y);
                const uint32_t area = (r.x1 - r.x0) * (r.y1 - r.y0);
                if (area != 0) {
                    if (p.colors_precomp == nullptr) {
                        uint8_t cl[3];
                        const float3 campos = make_float3(p.campos[3 * v], p.campos[3 * v + 1], p.campos[3 * v + 2]);
                        const float3 c = sh_to_rgb(p.sh_degree, p.shs + (size_t)g * p.sh_coeffs * 3, p_orig, campos, cl);
                        p.ws_rgb[3 * vg + 0] = c.x; p.ws_rgb[3 * vg + 1] = c.y; p.ws_rgb[3 * vg + 2] = c.z;
                        p.ws_clamped[3 * vg + 0] = cl[0]; p.ws_clamped[3 * vg + 1] = cl[1]; p.ws_clamped[3 * vg + 2] = cl[2];
                    }
                    radius_out = (int)my_radius;
                    tiles_out = area;
                    rec.a = make_float4(px, py, conic.x, conic.y);
                    rec.b = make_float4(conic.z, opacity, p_view.z, __int_as_float(radius_out));
                }
            }
        }
        p.radii[vg] = radius_out;
        p.ws_tiles[vg] = tiles_out;
        p.ws_rec[vg] = rec;
    }
}

// --------------------------------------------------------------------------- duplicate
// One thread per (view, Gaussian): emit key = (view*tiles + tile) << 32 | depth bits, value = Gaussian id.
__global__ void __launch_bounds__(256) duplicate_kernel(DuplicateParams p) {
    const size_t vg = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vg >= (size_t)p.V * p.P) return;
    const GeomRec rec = p.ws_rec[vg];
    const int radius = __float_as_int(rec.b.w);
    if (radius <= 0) return;
    const uint32_t v = (uint32_t)(vg / p.P);
    const uint32_t g = (uint32_t)(vg - (size_t)v * p.P);
    uint64_t off = (vg == 0) ? 0ull : p.ws_offsets[vg - 1];
    const TileRect r = tile_rect(rec.a.x, rec.a.y, radius, p.grid_x, p.grid_y);
    const uint32_t depth_bits = __float_as_uint(rec.b.z);
    const uint32_t tile_base = v * p.grid_x * p.grid_y;
    for (uint32_t y = r.y0; y < r.y1; ++y) {
        for (uint32_t x = r.x0; x < r.x1; ++x) {
            if (off < p.capacity) {  // capacity was verified on the host; belt and braces
                uint64_t key = tile_base + y * p.grid_x + x;
                key <<= 32;
                key |= depth_bits;
                p.keys[off] = key;
                p.vals[off] = g;
            }
            ++off;
        }
    }
}

// --------------------------------------------------------------------------- ranges + gather
// One thread per sorted instance: tile range boundaries (rasterizer_impl.cu:116-138) and the
// one-time gather into the tile-major attribute stream.
__global__ void __launch_bounds__(256) ranges_gather_kernel(GatherParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (p.dev_count != nullptr) p.R = (uint32_t)p.dev_count[0];  // sync-free mode
    if (i >= p.R) return;
    const uint64_t key = p.keys_sorted[i];
    const uint32_t tile = (uint32_t)(key >> 32);
    if (i == 0) {
        p.ranges[tile].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(p.keys_sorted[i - 1] >> 32);
        if (tile != prev) {
            p.ranges[prev].y = i;
            p.ranges[tile].x = i;
        }
    }
    if (i == p.R - 1) p.ranges[tile].y = p.R;

    const uint32_t g = p.vals_sorted[i];
    const uint32_t v = tile / p.tiles_per_view;
    const size_t vg = (size_t)v * p.P + g;
    const GeomRec rec = p.ws_rec[vg];
    const float* col = p.colors + (size_t)v * p.colors_view_stride + (size_t)g * 3;
    InstRec out;
    out.q0 = rec.a;
    out.q1 = make_float4(rec.b.x, rec.b.y, col[0], col[1]);
    out.q2 = make_float4(col[2], rec.b.z, __uint_as_float(g),
                         __uint_as_float(footprint_half_extent(rec.a.z, rec.a.w, rec.b.x, rec.b.y)));
    p.stream[i] = out;
}

// --------------------------------------------------------------------------- blend (forward)
// One CTA per 16x16 tile: 8 CONSUMER warps, each owning an 8x4 pixel block, plus 1 PRODUCER warp.  The producer streams
// the tile's depth-sorted attribute records through a ring of AGR_STAGES shared-memory buffers with cp.async.bulk
// (TMA bulk copy) and full/empty mbarriers; consumers never meet at a block-wide barrier, so a warp whose pixels saturate
// early or whose block is missed by most Gaussians (sub-tile culling below) runs ahead instead of waiting for its siblings.
// Per-pixel arithmetic and its order are exactly forward.cu:329-368.
template <int BATCH, int STAGES>
__global__ void __launch_bounds__(AGR_TILE_PIX + 32) blend_fwd_kernel(BlendFwdParams p) {
    __shared__ __align__(128) InstRec s_rec[STAGES][BATCH];
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES];
    __shared__ uint32_t s_warp_last[AGR_TILE_PIX / 32];
    __shared__ int s_warps_done;

    const uint32_t tile_lin = blockIdx.x;  // view*tiles + tile
    const uint32_t v = tile_lin / p.tiles_per_view;
    const uint32_t t = tile_lin - v * p.tiles_per_view;
    const uint32_t tile_y = t / p.grid_x, tile_x = t - tile_y * p.grid_x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t NCONS = AGR_TILE_PIX / 32;

    const uint2 range = p.ranges[tile_lin];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + BATCH - 1) / BATCH;
    const InstRec* src = p.stream + range.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], NCONS); }
        s_warps_done = 0;
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCONS) {
        // ================= producer warp (one elected lane) =================
        if (lane == 0) {
            int issued = 0;
            for (int i = 0; i < rounds; ++i) {
                const int s = i % STAGES;
                bool all_done = false;
                if (i >= STAGES) {
                    // wait for the stage to be released — or for the whole tile to finish, in which case the consumers
                    // have stopped arriving and nothing more must be fetched
                    while (!mbar_try_wait(&s_empty[s], ((i / STAGES) & 1) ^ 1)) {
                        if (*(volatile int*)&s_warps_done == (int)NCONS) { all_done = true; break; }
                    }
                }
                if (all_done || *(volatile int*)&s_warps_done == (int)NCONS) break;   // every pixel of the tile is finished
                const int n = min(BATCH, total - i * BATCH);
                bulk_load(&s_rec[s][0], src + (size_t)i * BATCH, (uint32_t)n * sizeof(InstRec), &s_full[s]);
                issued = i + 1;
            }
            // no copy may be in flight when the CTA retires: wait for the (at most STAGES) youngest ones
            for (int i = max(0, issued - STAGES); i < issued; ++i) mbar_wait(&s_full[i % STAGES], (i / STAGES) & 1);
        }
        return;
    }

    // ================= consumer warps =================
    const uint32_t px = tile_x * AGR_TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = tile_y * AGR_TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)p.W && py < (uint32_t)p.H;
    const float2 pixf = make_float2((float)px, (float)py);
    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    const float bx0 = (float)(tile_x * AGR_TILE_X + (warp & 1) * 8), bx1 = bx0 + 7.f;
    const float by0 = (float)(tile_y * AGR_TILE_Y + (warp >> 1) * 4), by1 = by0 + 3.f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, weight = 0.f, D = 0.f;
    bool warp_done = false;

    for (int i = 0; i < rounds; ++i) {
        if (!warp_done && __all_sync(0xffffffffu, done)) {
            warp_done = true;
            if (lane == 0) atomicAdd(&s_warps_done, 1);
        }
        const int s = i % STAGES;
        // Wait for batch i — or for the tile to be finished (all 8 warps saturated), after which the producer stops
        // fetching and batch i may never arrive.  Lane 0 polls, the verdict is broadcast so the warp leaves together.
        int stop = 0;
        if (lane == 0) {
            while (!mbar_try_wait(&s_full[s], (i / STAGES) & 1)) {
                if (*(volatile int*)&s_warps_done == (int)NCONS) { stop = 1; break; }
            }
        }
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) break;
        mbar_wait(&s_full[s], (i / STAGES) & 1);   // already complete: per-thread acquire of the TMA-written data
        const int n = min(BATCH, total - i * BATCH);
        // Sub-tile culling: lane l tests record c+l against this warp's 8x4 pixel block (bounding-box overlap with
        // the Gaussian's alpha >= 1/255 footprint); the warp then walks only the surviving records, in list order.
        for (int c = 0; c < n && !warp_done; c += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            const int jl = c + (int)lane;
            bool hit = false;
            if (jl < n) {
                const float4 t0 = s_rec[s][jl].q0;
                const float2 ext = unpack_extent(s_rec[s][jl].q2.w);
                hit = (t0.x + ext.x >= bx0) && (t0.x - ext.x <= bx1) && (t0.y + ext.y >= by0) && (t0.y - ext.y <= by1);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                if (done) continue;
                const int j = c + b;
                const float4 q0 = s_rec[s][j].q0;
                const float4 q1 = s_rec[s][j].q1;
                const float2 d = make_float2(q0.x - pixf.x, q0.y - pixf.y);
                const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
                if (power > 0.0f) continue;
                const float alpha = min(0.99f, q1.y * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) { done = true; continue; }
                const float4 q2 = s_rec[s][j].q2;
                C0 += q1.z * alpha * T;
                C1 += q1.w * alpha * T;
                C2 += q2.x * alpha * T;
                weight += alpha * T;
                D += q2.y * alpha * T;
                T = test_T;
                last_contributor = (uint32_t)(i * BATCH + j + 1);   // == the reference's running `contributor`
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
    }
    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix_id = (size_t)p.W * py + px;
        const float* bg = p.background + (size_t)v * p.bg_view_stride;
        p.n_contrib[v * HW + pix_id] = last_contributor;
        float* oc = p.out_color + (size_t)v * 3 * HW;
        oc[0 * HW + pix_id] = C0 + T * bg[0];
        oc[1 * HW + pix_id] = C1 + T * bg[1];
        oc[2 * HW + pix_id] = C2 + T * bg[2];
        p.out_alpha[v * HW + pix_id] = weight;
        p.out_depth[v * HW + pix_id] = D;
    }
    // tile-wide max of last_contributor: the backward never has to look past it
    uint32_t m = inside ? last_contributor : 0u;
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) s_warp_last[warp] = m;
    asm volatile("bar.sync 1, %0;" ::"r"(AGR_TILE_PIX) : "memory");   // the 8 consumer warps only
    if (threadIdx.x == 0) {
        uint32_t mm = 0;
#pragma unroll
        for (int w = 0; w < AGR_TILE_PIX / 32; ++w) mm = max(mm, s_warp_last[w]);
        p.tile_last[tile_lin] = mm;
    }
}

// Sync-free mode: publish the instance count on the device, flag overflow, and pad the unused tail of the key
// buffer with all-ones keys so that sorting `capacity` items leaves the R real instances in front.
__global__ void __launch_bounds__(256) finalize_count_kernel(const uint64_t* __restrict__ offsets_last, uint64_t capacity,
                                                            uint64_t* __restrict__ keys, int64_t* __restrict__ status) {
    const uint64_t total = *offsets_last;
    const uint64_t R = total < capacity ? total : capacity;
    if (blockIdx.x == 0 && threadIdx.x == 0) { status[0] = (int64_t)R; status[1] = total > capacity ? 1 : 0; }
    for (uint64_t i = R + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (uint64_t)gridDim.x * blockDim.x)
        keys[i] = ~0ull;
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    const float3 pt = make_float3(means3D[3 * g], means3D[3 * g + 1], means3D[3 * g + 2]);
    const float3 pv = xform_point_4x3(pt, view);
    present[g] = pv.z > 0.2f ? 1 : 0;
}

// --------------------------------------------------------------------------- launchers
void launch_preprocess_fwd(const PreprocessFwdParams& p, const ViewScalars& vs, cudaStream_t s) {
    preprocess_fwd_kernel<<<(p.P + 255) / 256, 256, 0, s>>>(p, vs);
}
void launch_duplicate(const DuplicateParams& p, cudaStream_t s) {
    const size_t n = (size_t)p.V * p.P;
    duplicate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
}
void launch_finalize_count(const uint64_t* offsets_last, uint64_t capacity, uint64_t* keys, int64_t* status, cudaStream_t s) {
    finalize_count_kernel<<<148 * 4, 256, 0, s>>>(offsets_last, capacity, keys, status);
}
void launch_ranges_gather(const GatherParams& p, cudaStream_t s) {
    if (p.R == 0) return;
    ranges_gather_kernel<<<(p.R + 255) / 256, 256, 0, s>>>(p);
}
void launch_blend_fwd(const BlendFwdParams& p, cudaStream_t s) {
    blend_fwd_kernel<AGR_BATCH, AGR_STAGES><<<p.num_tiles_total, AGR_TILE_PIX + 32, 0, s>>>(p);
}
void launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
}

}  // namespace agr
