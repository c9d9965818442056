/*
 * TEST INFRASTRUCTURE — CPU restatement (plain C, fp32) of the reference rasterizer.
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * reference legs may load it.  The product (animatablegaussians_b200) never does.
 *
 * Parity pin: validated against the UNMODIFIED reference CUDA kernels (oracle/_ref/
 * libref_rasterizer.so, built by oracle/build_ref.py) in tests/test_raster_gpu.py and against
 * the golden vectors in tests/golden/ that were produced by those kernels on a B200
 * (tests/golden/make_raster_golden.py).  The reference tree has no tests / golden vectors of
 * its own.
 *
 * Each function cites the reference code it restates; RAST =
 * gaussians/diff_gaussian_rasterization_depth_alpha in the reference tree.
 * Compile:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC raster_oracle.c -o _build/libraster_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16
#define BLOCK_Y 16

typedef struct {
    int P, W, H, gx, gy, M;
    int64_t R;
    /* per Gaussian (reference GeometryState, rasterizer_impl.h:29-44) */
    float *depths, *means2D, *cov3D, *conic_opacity, *rgb;
    int *radii;
    uint8_t *clamped;
    uint32_t *tiles_touched, *point_offsets;
    /* per instance (BinningState) */
    uint64_t *keys;
    uint32_t *point_list;
    /* per tile / pixel (ImageState) */
    uint32_t *ranges; /* 2 per tile */
    uint32_t *n_contrib;
} OracleState;

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                              -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                              -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* ---- column-major 3x3 helpers standing in for glm::mat3 (m[col][row]) ---- */
typedef struct { float c[3][3]; } m3;
static m3 m3_mul(m3 A, m3 B) { /* glm type_mat3x3.inl:486-518 */
    m3 R;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            R.c[j][i] = A.c[0][i] * B.c[j][0] + A.c[1][i] * B.c[j][1] + A.c[2][i] * B.c[j][2];
    return R;
}
static m3 m3_t(m3 A) { m3 R; for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) R.c[j][i] = A.c[i][j]; return R; }
static m3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
    m3 r = {{{a0, a1, a2}, {b0, b1, b2}, {c0, c1, c2}}};
    return r;
}

/* auxiliary.h:58-97 */
static void xf43(const float* p, const float* m, float* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void xf44(const float* p, const float* m, float* o) {
    xf43(p, m, o);
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}
/* auxiliary.h:41-44 (double on purpose) */
static float ndc2pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }
static uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
/* auxiliary.h:46-56 */
static void get_rect(float px, float py, int r, int gx, int gy, uint32_t* mn, uint32_t* mx) {
    mn[0] = umin((uint32_t)gx, (uint32_t)imax(0, (int)((px - r) / BLOCK_X)));
    mn[1] = umin((uint32_t)gy, (uint32_t)imax(0, (int)((py - r) / BLOCK_Y)));
    mx[0] = umin((uint32_t)gx, (uint32_t)imax(0, (int)((px + r + BLOCK_X - 1) / BLOCK_X)));
    mx[1] = umin((uint32_t)gy, (uint32_t)imax(0, (int)((py + r + BLOCK_Y - 1) / BLOCK_Y)));
}

/* forward.cu:118-152 — NOTE: quaternion deliberately NOT normalised (forward.cu:127) */
static void build_R(const float* q, m3* R) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    *R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                 2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                 2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}
static void compute_cov3D(const float* scale, float mod, const float* q, float* cov3D) {
    m3 S = m3_cols(mod * scale[0], 0, 0, 0, mod * scale[1], 0, 0, 0, mod * scale[2]);
    m3 R; build_R(q, &R);
    m3 M = m3_mul(S, R);
    m3 Sg = m3_mul(m3_t(M), M);
    cov3D[0] = Sg.c[0][0]; cov3D[1] = Sg.c[0][1]; cov3D[2] = Sg.c[0][2];
    cov3D[3] = Sg.c[1][1]; cov3D[4] = Sg.c[1][2]; cov3D[5] = Sg.c[2][2];
}

/* forward.cu:74-113 (and its re-computation in backward.cu:166-199) */
static void compute_cov2D(const float* mean, float fx, float fy, float tanx, float tany, const float* cov3D,
                          const float* view, float* cov, m3* Tout, float* tout, float* xmul, float* ymul) {
    float t[3]; xf43(mean, view, t);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    if (xmul) *xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    if (ymul) *ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    m3 J = m3_cols(fx / t[2], 0.0f, -(fx * t[0]) / (t[2] * t[2]), 0.0f, fy / t[2], -(fy * t[1]) / (t[2] * t[2]), 0, 0, 0);
    m3 Wm = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    m3 T = m3_mul(Wm, J);
    m3 Vrk = m3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    m3 c = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
    cov[0] = c.c[0][0] + 0.3f; cov[1] = c.c[0][1]; cov[2] = c.c[1][1] + 0.3f;
    if (Tout) *Tout = T;
    if (tout) { tout[0] = t[0]; tout[1] = t[1]; tout[2] = t[2]; }
}

/* forward.cu:20-71 */
static void sh_to_rgb(int idx, int deg, int M, const float* means, const float* campos, const float* shs,
                      uint8_t* clamped, float* out) {
    const float* pos = means + 3 * idx;
    float dir[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
    const float* sh = shs + (size_t)idx * M * 3;
    for (int c = 0; c < 3; ++c) {
        float result = SH_C0 * sh[c];
        if (deg > 0) {
            result = result - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                result = result + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] +
                         SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + SH_C2[3] * xz * sh[21 + c] +
                         SH_C2[4] * (xx - yy) * sh[24 + c];
                if (deg > 2) {
                    result = result + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + SH_C3[1] * xy * z * sh[30 + c] +
                             SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                             SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                             SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + SH_C3[5] * z * (xx - yy) * sh[42 + c] +
                             SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
                }
            }
        }
        result += 0.5f;
        clamped[3 * idx + c] = (result < 0);
        out[c] = fmaxf(result, 0.0f);
    }
}

typedef struct { uint64_t key; uint32_t val; uint32_t order; } kv_t;
static int kv_cmp(const void* a, const void* b) {
    const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0); /* stable, like CUB radix sort */
}

static void free_state(OracleState* s) {
    free(s->depths); free(s->means2D); free(s->cov3D); free(s->conic_opacity); free(s->rgb); free(s->radii);
    free(s->clamped); free(s->tiles_touched); free(s->point_offsets); free(s->keys); free(s->point_list);
    free(s->ranges); free(s->n_contrib);
    memset(s, 0, sizeof(*s));
}

void* oracle_raster_create(void) { return calloc(1, sizeof(OracleState)); }
void oracle_raster_destroy(void* h) { if (h) { free_state((OracleState*)h); free(h); } }

/* Rasterizer::forward, rasterizer_impl.cu:197-339. Returns num_rendered. */
int64_t oracle_raster_forward(void* h, int P, int D, int M, const float* background, int W, int H,
                              const float* means3D, const float* shs, const float* colors_precomp,
                              const float* opacities, const float* scales, float scale_modifier,
                              const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                              const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                              float* out_color, float* out_depth, float* out_alpha, int* radii_out) {
    OracleState* s = (OracleState*)h;
    free_state(s);
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    s->P = P; s->W = W; s->H = H; s->gx = gx; s->gy = gy; s->M = M;
    const size_t Pn = P > 0 ? P : 1;
    s->depths = calloc(Pn, 4); s->means2D = calloc(Pn * 2, 4); s->cov3D = calloc(Pn * 6, 4);
    s->conic_opacity = calloc(Pn * 4, 4); s->rgb = calloc(Pn * 3, 4); s->radii = calloc(Pn, 4);
    s->clamped = calloc(Pn * 3, 1); s->tiles_touched = calloc(Pn, 4); s->point_offsets = calloc(Pn, 4);
    s->ranges = calloc((size_t)gx * gy * 2, 4); s->n_contrib = calloc((size_t)W * H, 4);
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);

    /* preprocessCUDA, forward.cu:155-256 */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        s->radii[idx] = 0; s->tiles_touched[idx] = 0;
        const float* p_orig = means3D + 3 * idx;
        float p_view[3]; xf43(p_orig, viewmatrix, p_view);
        if (p_view[2] <= 0.2f) continue; /* in_frustum, auxiliary.h:139-164 */
        float p_hom[4]; xf44(p_orig, projmatrix, p_hom);
        float p_w = 1.0f / (p_hom[3] + 0.0000001f);
        float p_proj[3] = {p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w};
        const float* cov3D;
        if (cov3D_precomp) cov3D = cov3D_precomp + 6 * idx;
        else { compute_cov3D(scales + 3 * idx, scale_modifier, rotations + 4 * idx, s->cov3D + 6 * idx); cov3D = s->cov3D + 6 * idx; }
        float cov[3];
        compute_cov2D(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, cov, NULL, NULL, NULL, NULL);
        float det = (cov[0] * cov[2] - cov[1] * cov[1]);
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float px = ndc2pix(p_proj[0], W), py = ndc2pix(p_proj[1], H);
        uint32_t mn[2], mx[2];
        get_rect(px, py, (int)my_radius, gx, gy, mn, mx);
        if ((mx[0] - mn[0]) * (mx[1] - mn[1]) == 0) continue;
        if (!colors_precomp) sh_to_rgb(idx, D, M, means3D, campos, shs, s->clamped, s->rgb + 3 * idx);
        s->depths[idx] = p_view[2];
        s->radii[idx] = (int)my_radius;
        s->means2D[2 * idx] = px; s->means2D[2 * idx + 1] = py;
        s->conic_opacity[4 * idx] = conic[0]; s->conic_opacity[4 * idx + 1] = conic[1];
        s->conic_opacity[4 * idx + 2] = conic[2]; s->conic_opacity[4 * idx + 3] = opacities[idx];
        s->tiles_touched[idx] = (mx[1] - mn[1]) * (mx[0] - mn[0]);
    }
    if (radii_out) memcpy(radii_out, s->radii, (size_t)P * 4);

    /* InclusiveSum, rasterizer_impl.cu:278 */
    uint32_t acc = 0;
    for (int i = 0; i < P; ++i) { acc += s->tiles_touched[i]; s->point_offsets[i] = acc; }
    const int64_t R = acc;
    s->R = R;

    /* duplicateWithKeys, rasterizer_impl.cu:70-111 + stable sort (rasterizer_impl.cu:304-309) */
    kv_t* kv = malloc(sizeof(kv_t) * (R > 0 ? R : 1));
    for (int idx = 0; idx < P; ++idx) {
        if (s->radii[idx] <= 0) continue;
        uint32_t off = idx == 0 ? 0 : s->point_offsets[idx - 1];
        uint32_t mn[2], mx[2];
        get_rect(s->means2D[2 * idx], s->means2D[2 * idx + 1], s->radii[idx], gx, gy, mn, mx);
        uint32_t dbits; memcpy(&dbits, &s->depths[idx], 4);
        for (uint32_t y = mn[1]; y < mx[1]; ++y)
            for (uint32_t x = mn[0]; x < mx[0]; ++x) {
                uint64_t key = (uint64_t)(y * gx + x);
                key <<= 32; key |= dbits;
                kv[off].key = key; kv[off].val = idx; kv[off].order = off; off++;
            }
    }
    qsort(kv, R, sizeof(kv_t), kv_cmp);
    s->keys = malloc(8 * (R > 0 ? R : 1)); s->point_list = malloc(4 * (R > 0 ? R : 1));
    for (int64_t i = 0; i < R; ++i) { s->keys[i] = kv[i].key; s->point_list[i] = kv[i].val; }
    free(kv);

    /* identifyTileRanges, rasterizer_impl.cu:116-138 */
    for (int64_t i = 0; i < R; ++i) {
        uint32_t cur = (uint32_t)(s->keys[i] >> 32);
        if (i == 0) s->ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(s->keys[i - 1] >> 32);
            if (cur != prev) { s->ranges[2 * prev + 1] = (uint32_t)i; s->ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) s->ranges[2 * cur + 1] = (uint32_t)R;
    }

    /* renderCUDA, forward.cu:261-381 — one pixel at a time */
    const float* features = colors_precomp ? colors_precomp : s->rgb;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; ++tile) {
        const int ty = tile / gx, tx = tile % gx;
        const uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const size_t pix_id = (size_t)W * pyi + pxi;
                const float pixx = (float)pxi, pixy = (float)pyi;
                float T = 1.0f, C[3] = {0, 0, 0}, weight = 0, Dp = 0;
                uint32_t contributor = 0, last_contributor = 0;
                for (uint32_t k = r0; k < r1; ++k) {
                    contributor++;
                    const uint32_t id = s->point_list[k];
                    const float dx = s->means2D[2 * id] - pixx, dy = s->means2D[2 * id + 1] - pixy;
                    const float* co = s->conic_opacity + 4 * id;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break; /* done = true */
                    for (int ch = 0; ch < 3; ++ch) C[ch] += features[id * 3 + ch] * alpha * T;
                    weight += alpha * T;
                    Dp += s->depths[id] * alpha * T;
                    T = test_T;
                    last_contributor = contributor;
                }
                s->n_contrib[pix_id] = last_contributor;
                for (int ch = 0; ch < 3; ++ch) out_color[ch * HW + pix_id] = C[ch] + T * background[ch];
                out_alpha[pix_id] = weight;
                out_depth[pix_id] = Dp;
            }
    }
    return R;
}

/* backward.cu:20-139 */
static void sh_backward(int idx, int deg, int M, const float* means, const float* campos, const float* shs,
                        const uint8_t* clamped, const float* dL_dcolor, float* dL_dmeans, float* dL_dshs) {
    const float* pos = means + 3 * idx;
    float dir_orig[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
    float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
    const float* sh = shs + (size_t)idx * M * 3;
    float dRGB[3];
    for (int c = 0; c < 3; ++c) dRGB[c] = dL_dcolor[3 * idx + c] * (clamped[3 * idx + c] ? 0 : 1);
    float* dsh = dL_dshs + (size_t)idx * M * 3;
    float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
    float coef[16]; int n = 1;
    coef[0] = SH_C0;
    if (deg > 0) {
        coef[1] = -SH_C1 * y; coef[2] = SH_C1 * z; coef[3] = -SH_C1 * x; n = 4;
        for (int c = 0; c < 3; ++c) { dx[c] = -SH_C1 * sh[9 + c]; dy[c] = -SH_C1 * sh[3 + c]; dz[c] = SH_C1 * sh[6 + c]; }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            coef[4] = SH_C2[0] * xy; coef[5] = SH_C2[1] * yz; coef[6] = SH_C2[2] * (2.f * zz - xx - yy);
            coef[7] = SH_C2[3] * xz; coef[8] = SH_C2[4] * (xx - yy); n = 9;
            for (int c = 0; c < 3; ++c) {
                dx[c] += SH_C2[0] * y * sh[12 + c] + SH_C2[2] * 2.f * -x * sh[18 + c] + SH_C2[3] * z * sh[21 + c] + SH_C2[4] * 2.f * x * sh[24 + c];
                dy[c] += SH_C2[0] * x * sh[12 + c] + SH_C2[1] * z * sh[15 + c] + SH_C2[2] * 2.f * -y * sh[18 + c] + SH_C2[4] * 2.f * -y * sh[24 + c];
                dz[c] += SH_C2[1] * y * sh[15 + c] + SH_C2[2] * 2.f * 2.f * z * sh[18 + c] + SH_C2[3] * x * sh[21 + c];
            }
            if (deg > 2) {
                coef[9] = SH_C3[0] * y * (3.f * xx - yy); coef[10] = SH_C3[1] * xy * z;
                coef[11] = SH_C3[2] * y * (4.f * zz - xx - yy); coef[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                coef[13] = SH_C3[4] * x * (4.f * zz - xx - yy); coef[14] = SH_C3[5] * z * (xx - yy);
                coef[15] = SH_C3[6] * x * (xx - 3.f * yy); n = 16;
                for (int c = 0; c < 3; ++c) {
                    dx[c] += (SH_C3[0] * sh[27 + c] * 3.f * 2.f * xy + SH_C3[1] * sh[30 + c] * yz + SH_C3[2] * sh[33 + c] * -2.f * xy +
                              SH_C3[3] * sh[36 + c] * -3.f * 2.f * xz + SH_C3[4] * sh[39 + c] * (-3.f * xx + 4.f * zz - yy) +
                              SH_C3[5] * sh[42 + c] * 2.f * xz + SH_C3[6] * sh[45 + c] * 3.f * (xx - yy));
                    dy[c] += (SH_C3[0] * sh[27 + c] * 3.f * (xx - yy) + SH_C3[1] * sh[30 + c] * xz +
                              SH_C3[2] * sh[33 + c] * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * sh[36 + c] * -3.f * 2.f * yz +
                              SH_C3[4] * sh[39 + c] * -2.f * xy + SH_C3[5] * sh[42 + c] * -2.f * yz + SH_C3[6] * sh[45 + c] * -3.f * 2.f * xy);
                    dz[c] += (SH_C3[1] * sh[30 + c] * xy + SH_C3[2] * sh[33 + c] * 4.f * 2.f * yz +
                              SH_C3[3] * sh[36 + c] * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * sh[39 + c] * 4.f * 2.f * xz +
                              SH_C3[5] * sh[42 + c] * (xx - yy));
                }
            }
        }
    }
    for (int k = 0; k < n && k < M; ++k) for (int c = 0; c < 3; ++c) dsh[3 * k + c] = coef[k] * dRGB[c];
    float dd[3] = {dx[0] * dRGB[0] + dx[1] * dRGB[1] + dx[2] * dRGB[2], dy[0] * dRGB[0] + dy[1] * dRGB[1] + dy[2] * dRGB[2],
                   dz[0] * dRGB[0] + dz[1] * dRGB[1] + dz[2] * dRGB[2]};
    /* dnormvdv, auxiliary.h:107-117 */
    const float* v = dir_orig;
    float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dL_dmeans[3 * idx + 0] += ((+sum2 - v[0] * v[0]) * dd[0] - v[1] * v[0] * dd[1] - v[2] * v[0] * dd[2]) * inv;
    dL_dmeans[3 * idx + 1] += (-v[0] * v[1] * dd[0] + (sum2 - v[1] * v[1]) * dd[1] - v[2] * v[1] * dd[2]) * inv;
    dL_dmeans[3 * idx + 2] += (-v[0] * v[2] * dd[0] - v[1] * v[2] * dd[1] + (sum2 - v[2] * v[2]) * dd[2]) * inv;
}

/* Rasterizer::backward, rasterizer_impl.cu:343-447. All outputs must be zero-filled by the caller
 * (rasterize_points.cu:158-167). dL_dconic is (P,4) used as x,y,-,w (backward.cu:165,593-595). */
void oracle_raster_backward(void* h, int P, int D, int M, const float* background, int W, int H,
                            const float* means3D, const float* shs, const float* colors_precomp, const float* alphas,
                            const float* scales, float scale_modifier, const float* rotations,
                            const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                            const float* campos, float tan_fovx, float tan_fovy,
                            const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas,
                            float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth,
                            float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot) {
    OracleState* s = (OracleState*)h;
    const int gx = s->gx, gy = s->gy;
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    const float* colors = colors_precomp ? colors_precomp : s->rgb;
    const size_t HW = (size_t)H * W;
    const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H);

    /* renderCUDA (bwd), backward.cu:415-601 */
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; ++tile) {
        const int ty = tile / gx, tx = tile % gx;
        const uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const size_t pix_id = (size_t)W * pyi + pxi;
                const float pixx = (float)pxi, pixy = (float)pyi;
                const float T_final = 1 - alphas[pix_id];
                float T = T_final;
                uint32_t contributor = r1 - r0;
                const uint32_t last_contributor = s->n_contrib[pix_id];
                float accum_rec[3] = {0, 0, 0}, dL_dpixel[3], accum_depth_rec = 0, accum_alpha_rec = 0;
                for (int i = 0; i < 3; ++i) dL_dpixel[i] = dL_dpix[i * HW + pix_id];
                const float dL_dpixel_depth = dL_dpix_depth[pix_id], dL_dalpha = dL_dalphas[pix_id];
                float last_alpha = 0, last_color[3] = {0, 0, 0}, last_depth = 0;
                for (uint32_t kk = r1; kk > r0; --kk) {
                    const uint32_t k = kk - 1;
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t id = s->point_list[k];
                    const float dx = s->means2D[2 * id] - pixx, dy = s->means2D[2 * id + 1] - pixy;
                    const float* co = s->conic_opacity + 4 * id;
                    const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dopa = 0.0f;
                    for (int ch = 0; ch < 3; ++ch) {
                        const float c = colors[id * 3 + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dopa += (c - accum_rec[ch]) * dL_dpixel[ch];
#pragma omp atomic
                        dL_dcolor[id * 3 + ch] += dchannel_dcolor * dL_dpixel[ch];
                    }
                    const float c_d = s->depths[id];
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dopa += (c_d - accum_depth_rec) * dL_dpixel_depth;
#pragma omp atomic
                    dL_ddepth[id] += dchannel_dcolor * dL_dpixel_depth;
                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dopa += (1 - accum_alpha_rec) * dL_dalpha;
                    dL_dopa *= T;
                    last_alpha = alpha;
                    float bg_dot_dpixel = 0;
                    for (int i = 0; i < 3; ++i) bg_dot_dpixel += background[i] * dL_dpixel[i];
                    dL_dopa += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                    const float dL_dG = co[3] * dL_dopa;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
#pragma omp atomic
                    dL_dmean2D[3 * id + 0] += dL_dG * dG_ddelx * ddelx_dx;
#pragma omp atomic
                    dL_dmean2D[3 * id + 1] += dL_dG * dG_ddely * ddely_dy;
#pragma omp atomic
                    dL_dconic[4 * id + 0] += -0.5f * gdx * dx * dL_dG;
#pragma omp atomic
                    dL_dconic[4 * id + 1] += -0.5f * gdx * dy * dL_dG;
#pragma omp atomic
                    dL_dconic[4 * id + 3] += -0.5f * gdy * dy * dL_dG;
#pragma omp atomic
                    dL_dopacity[id] += G * dL_dopa;
                }
            }
    }

    /* computeCov2DCUDA (backward.cu:144-274) + preprocessCUDA bwd (backward.cu:346-412) */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        if (!(s->radii[idx] > 0)) continue;
        const float* cov3D = cov3D_precomp ? cov3D_precomp + 6 * idx : s->cov3D + 6 * idx;
        const float* mean = means3D + 3 * idx;
        const float dcx = dL_dconic[4 * idx], dcy = dL_dconic[4 * idx + 1], dcz = dL_dconic[4 * idx + 3];
        float cov[3], t[3], xm, ym; m3 T;
        compute_cov2D(mean, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, cov, &T, t, &xm, &ym);
        const float a = cov[0], b = cov[1], c = cov[2];
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcov = dL_dcov3D + 6 * idx;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
            dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
            dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
            dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
            dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
            dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
        } else {
            for (int i = 0; i < 6; ++i) dcov[i] = 0;
        }
        const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        float dT0[3], dT1[3];
        for (int j = 0; j < 3; ++j) {
            dT0[j] = 2 * (T.c[0][0] * V[j][0] + T.c[0][1] * V[j][1] + T.c[0][2] * V[j][2]) * dL_da +
                     (T.c[1][0] * V[j][0] + T.c[1][1] * V[j][1] + T.c[1][2] * V[j][2]) * dL_db;
            dT1[j] = 2 * (T.c[1][0] * V[j][0] + T.c[1][1] * V[j][1] + T.c[1][2] * V[j][2]) * dL_dc +
                     (T.c[0][0] * V[j][0] + T.c[0][1] * V[j][1] + T.c[0][2] * V[j][2]) * dL_db;
        }
        const float* vm = viewmatrix;
        const float Wm[3][3] = {{vm[0], vm[4], vm[8]}, {vm[1], vm[5], vm[9]}, {vm[2], vm[6], vm[10]}};
        const float dL_dJ00 = Wm[0][0] * dT0[0] + Wm[0][1] * dT0[1] + Wm[0][2] * dT0[2];
        const float dL_dJ02 = Wm[2][0] * dT0[0] + Wm[2][1] * dT0[1] + Wm[2][2] * dT0[2];
        const float dL_dJ11 = Wm[1][0] * dT1[0] + Wm[1][1] * dT1[1] + Wm[1][2] * dT1[2];
        const float dL_dJ12 = Wm[2][0] * dT1[0] + Wm[2][1] * dT1[1] + Wm[2][2] * dT1[2];
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = xm * -focal_x * tz2 * dL_dJ02;
        const float dL_dty = ym * -focal_y * tz2 * dL_dJ12;
        const float dL_dtz = -focal_x * tz2 * dL_dJ00 - focal_y * tz2 * dL_dJ11 + (2 * focal_x * t[0]) * tz3 * dL_dJ02 + (2 * focal_y * t[1]) * tz3 * dL_dJ12;
        float* dm = dL_dmean3D + 3 * idx;
        dm[0] = vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz;
        dm[1] = vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz;
        dm[2] = vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz;

        const float* pj = projmatrix;
        float m_hom[4]; xf44(mean, pj, m_hom);
        const float m_w = 1.0f / (m_hom[3] + 0.0000001f);
        const float mul1 = (pj[0] * mean[0] + pj[4] * mean[1] + pj[8] * mean[2] + pj[12]) * m_w * m_w;
        const float mul2 = (pj[1] * mean[0] + pj[5] * mean[1] + pj[9] * mean[2] + pj[13]) * m_w * m_w;
        const float g2x = dL_dmean2D[3 * idx], g2y = dL_dmean2D[3 * idx + 1];
        dm[0] += (pj[0] * m_w - pj[3] * mul1) * g2x + (pj[1] * m_w - pj[3] * mul2) * g2y;
        dm[1] += (pj[4] * m_w - pj[7] * mul1) * g2x + (pj[5] * m_w - pj[7] * mul2) * g2y;
        dm[2] += (pj[8] * m_w - pj[11] * mul1) * g2x + (pj[9] * m_w - pj[11] * mul2) * g2y;
        const float mul3 = vm[2] * mean[0] + vm[6] * mean[1] + vm[10] * mean[2] + vm[14];
        dm[0] += (vm[2] - vm[3] * mul3) * dL_ddepth[idx];
        dm[1] += (vm[6] - vm[7] * mul3) * dL_ddepth[idx];
        dm[2] += (vm[10] - vm[11] * mul3) * dL_ddepth[idx];

        if (shs) sh_backward(idx, D, M, means3D, campos, shs, s->clamped, dL_dcolor, dL_dmean3D, dL_dsh);

        if (scales) { /* computeCov3D bwd, backward.cu:278-341 */
            const float* q = rotations + 4 * idx;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            m3 R; build_R(q, &R);
            const float sx = scale_modifier * scales[3 * idx], sy = scale_modifier * scales[3 * idx + 1], sz = scale_modifier * scales[3 * idx + 2];
            m3 S = m3_cols(sx, 0, 0, 0, sy, 0, 0, 0, sz);
            m3 Mm = m3_mul(S, R);
            m3 dSig = m3_cols(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4], 0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
            m3 M2 = Mm;
            for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) M2.c[j][i] = Mm.c[j][i] * 2.0f;
            m3 dM = m3_mul(M2, dSig);
            m3 Rt = m3_t(R), dMt = m3_t(dM);
            float* ds = dL_dscale + 3 * idx;
            ds[0] = Rt.c[0][0] * dMt.c[0][0] + Rt.c[0][1] * dMt.c[0][1] + Rt.c[0][2] * dMt.c[0][2];
            ds[1] = Rt.c[1][0] * dMt.c[1][0] + Rt.c[1][1] * dMt.c[1][1] + Rt.c[1][2] * dMt.c[1][2];
            ds[2] = Rt.c[2][0] * dMt.c[2][0] + Rt.c[2][1] * dMt.c[2][1] + Rt.c[2][2] * dMt.c[2][2];
            for (int i = 0; i < 3; ++i) { dMt.c[0][i] *= sx; dMt.c[1][i] *= sy; dMt.c[2][i] *= sz; }
            float* dq = dL_drot + 4 * idx;
            dq[0] = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) + 2 * x * (dMt.c[1][2] - dMt.c[2][1]);
            dq[1] = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) + 2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
            dq[2] = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) + 2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
            dq[3] = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) + 2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
        }
    }
}

/* checkFrustum, rasterizer_impl.cu:54-66 */
void oracle_raster_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present) {
    (void)projmatrix;
    for (int i = 0; i < P; ++i) {
        float pv[3]; xf43(means3D + 3 * i, viewmatrix, pv);
        present[i] = pv[2] > 0.2f;
    }
}
