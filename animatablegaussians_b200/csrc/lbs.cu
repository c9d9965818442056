// Fused per-Gaussian linear blend skinning (forward + backward), include/agr_lbs.h.
//
// Reference being replaced: network/avatar.py:84-91 (transform_cano2live) and, for the points-only
// variant, avatar.py:128-131,150-151.  pytorch3d 0.7.4 quaternion_to_matrix / matrix_to_quaternion are
// restated below (real-first quaternion; quaternion_to_matrix divides by |q|^2; matrix_to_quaternion
// picks the best-conditioned of four candidates, floor 0.1 on the divisor, no sign standardisation).
//
// Layout / roofline: the kernel is HBM-bound — 4*J bytes of weights per Gaussian dominate (220 B of the
// 276 B/Gaussian forward traffic at J = 55).  Each CTA streams its contiguous [128 x J] weight slab into
// shared memory with one cp.async.bulk (TMA bulk copy) while the joint matrices (J x 12 floats) are
// loaded cooperatively; each thread then owns one Gaussian: its weight row is read from shared memory
// (stride J, J odd -> conflict-free) and the joint matrices are broadcast reads.
#include "../../include/agr_lbs.h"
#include "../../include/agr_rasterizer.h"
#include "raster_kernels.cuh"

namespace agr {

constexpr int LBS_THREADS = 128;

struct Q2M { float m[9]; };

// pytorch3d.transforms.quaternion_to_matrix
__device__ __forceinline__ Q2M quat_to_mat(float r, float i, float j, float k) {
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    Q2M o;
    o.m[0] = 1 - two_s * (j * j + k * k); o.m[1] = two_s * (i * j - k * r);     o.m[2] = two_s * (i * k + j * r);
    o.m[3] = two_s * (i * j + k * r);     o.m[4] = 1 - two_s * (i * i + k * k); o.m[5] = two_s * (j * k - i * r);
    o.m[6] = two_s * (i * k - j * r);     o.m[7] = two_s * (j * k + i * r);     o.m[8] = 1 - two_s * (i * i + j * j);
    return o;
}

__device__ __forceinline__ float sqrt_pos(float x) { return x > 0.f ? sqrtf(x) : 0.f; }

// pytorch3d.transforms.matrix_to_quaternion (0.7.4). m row-major 3x3. Returns the selected candidate row
// and (for the backward) which one was selected.
__device__ __forceinline__ float4 mat_to_quat(const float* m, int* sel_out) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    const float qa0 = sqrt_pos(1.0f + m00 + m11 + m22);
    const float qa1 = sqrt_pos(1.0f + m00 - m11 - m22);
    const float qa2 = sqrt_pos(1.0f - m00 + m11 - m22);
    const float qa3 = sqrt_pos(1.0f - m00 - m11 + m22);
    int sel = 0; float best = qa0;   // torch.argmax: first maximal index
    if (qa1 > best) { best = qa1; sel = 1; }
    if (qa2 > best) { best = qa2; sel = 2; }
    if (qa3 > best) { best = qa3; sel = 3; }
    const float den = 2.0f * fmaxf(best, 0.1f);
    float4 q;
    if (sel == 0)      q = make_float4(qa0 * qa0, m21 - m12, m02 - m20, m10 - m01);
    else if (sel == 1) q = make_float4(m21 - m12, qa1 * qa1, m10 + m01, m02 + m20);
    else if (sel == 2) q = make_float4(m02 - m20, m10 + m01, qa2 * qa2, m12 + m21);
    else               q = make_float4(m10 - m01, m20 + m02, m21 + m12, qa3 * qa3);
    q.x = q.x / den; q.y = q.y / den; q.z = q.z / den; q.w = q.w / den;
    if (sel_out) *sel_out = sel;
    return q;
}

// Blend J joint matrices with this thread's weight row (shared memory) -> 12 floats (rows 0..2).
__device__ __forceinline__ void blend(const float* __restrict__ w_row, const float* __restrict__ s_mats, int J, float* M) {
#pragma unroll
    for (int e = 0; e < 12; ++e) M[e] = 0.f;
    for (int j = 0; j < J; ++j) {
        const float w = w_row[j];
        const float4* a = reinterpret_cast<const float4*>(s_mats + 12 * j);
        const float4 a0 = a[0], a1 = a[1], a2 = a[2];
        M[0] += w * a0.x; M[1] += w * a0.y; M[2]  += w * a0.z; M[3]  += w * a0.w;
        M[4] += w * a1.x; M[5] += w * a1.y; M[6]  += w * a1.z; M[7]  += w * a1.w;
        M[8] += w * a2.x; M[9] += w * a2.y; M[10] += w * a2.z; M[11] += w * a2.w;
    }
}

template <bool WITH_ROT>
__global__ void __launch_bounds__(LBS_THREADS) lbs_fwd_kernel(int N, int J, const float* __restrict__ weights,
                                                             const float* __restrict__ jnt, const float* __restrict__ xyz_in,
                                                             const float* __restrict__ aux_in,  // rot (N,4) or vec (N,3) or null
                                                             float* __restrict__ xyz_out, float* __restrict__ aux_out,
                                                             float* __restrict__ pt_mats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_w = reinterpret_cast<float*>(smem_raw);                    // LBS_THREADS * J
    float* s_mats = s_w + ((LBS_THREADS * J + 3) / 4) * 4;              // J * 12, 16B aligned
    __shared__ __align__(8) uint64_t s_bar;

    const int n0 = blockIdx.x * LBS_THREADS;
    const int rows = min(LBS_THREADS, N - n0);
    if (threadIdx.x == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    __syncthreads();
    const size_t slab_bytes = (size_t)rows * J * sizeof(float);
    const float* slab = weights + (size_t)n0 * J;
    // bulk copy needs 16-byte aligned source and size; the slab start is aligned when (n0*J*4) % 16 == 0
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(slab) & 15) == 0) && ((slab_bytes & 15) == 0);
    if (bulk_ok) {
        if (threadIdx.x == 0) bulk_load(s_w, slab, (uint32_t)slab_bytes, &s_bar);
    } else {
        for (int i = threadIdx.x; i < rows * J; i += LBS_THREADS) s_w[i] = slab[i];
    }
    for (int i = threadIdx.x; i < J * 12; i += LBS_THREADS) {
        const int j = i / 12, e = i - 12 * j;
        s_mats[i] = jnt[16 * j + e];   // rows 0..2 of the 4x4
    }
    if (bulk_ok) mbar_wait(&s_bar, 0);
    __syncthreads();

    const int n = n0 + threadIdx.x;
    if (n >= N) return;
    float M[12];
    blend(s_w + threadIdx.x * J, s_mats, J, M);

    const float x = xyz_in[3 * n], y = xyz_in[3 * n + 1], z = xyz_in[3 * n + 2];
    xyz_out[3 * n + 0] = M[0] * x + M[1] * y + M[2] * z + M[3];
    xyz_out[3 * n + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
    xyz_out[3 * n + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];

    if (WITH_ROT) {
        const float4 q = reinterpret_cast<const float4*>(aux_in)[n];
        const Q2M R = quat_to_mat(q.x, q.y, q.z, q.w);
        float P[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                P[3 * r + c] = M[4 * r + 0] * R.m[c] + M[4 * r + 1] * R.m[3 + c] + M[4 * r + 2] * R.m[6 + c];
        reinterpret_cast<float4*>(aux_out)[n] = mat_to_quat(P, nullptr);
        if (pt_mats) {
            float4* o = reinterpret_cast<float4*>(pt_mats + (size_t)12 * n);
            o[0] = make_float4(M[0], M[1], M[2], M[3]);
            o[1] = make_float4(M[4], M[5], M[6], M[7]);
            o[2] = make_float4(M[8], M[9], M[10], M[11]);
        }
    } else if (aux_in != nullptr) {
        const float vx = aux_in[3 * n], vy = aux_in[3 * n + 1], vz = aux_in[3 * n + 2];
        aux_out[3 * n + 0] = M[0] * vx + M[1] * vy + M[2] * vz;
        aux_out[3 * n + 1] = M[4] * vx + M[5] * vy + M[6] * vz;
        aux_out[3 * n + 2] = M[8] * vx + M[9] * vy + M[10] * vz;
    }
}

__global__ void __launch_bounds__(256) lbs_bwd_kernel(int N, const float* __restrict__ pt_mats, const float* __restrict__ rot_in,
                                                     const float* __restrict__ d_xyz_out, const float* __restrict__ d_rot_out,
                                                     float* __restrict__ d_xyz_in, float* __restrict__ d_rot_in) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float4* mp = reinterpret_cast<const float4*>(pt_mats + (size_t)12 * n);
    const float4 r0 = mp[0], r1 = mp[1], r2 = mp[2];
    const float M[9] = {r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};

    // positions: x' = M x + t  ->  dx = M^T dx'
    const float gx = d_xyz_out[3 * n], gy = d_xyz_out[3 * n + 1], gz = d_xyz_out[3 * n + 2];
    d_xyz_in[3 * n + 0] = M[0] * gx + M[3] * gy + M[6] * gz;
    d_xyz_in[3 * n + 1] = M[1] * gx + M[4] * gy + M[7] * gz;
    d_xyz_in[3 * n + 2] = M[2] * gx + M[5] * gy + M[8] * gz;

    // rotations: q' = m2q(P), P = M R(q)
    const float4 q = reinterpret_cast<const float4*>(rot_in)[n];
    const Q2M R = quat_to_mat(q.x, q.y, q.z, q.w);
    float P[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            P[3 * r + c] = M[3 * r + 0] * R.m[c] + M[3 * r + 1] * R.m[3 + c] + M[3 * r + 2] * R.m[6 + c];
    int sel;
    const float4 qo = mat_to_quat(P, &sel);
    const float4 g = reinterpret_cast<const float4*>(d_rot_out)[n];

    // ---- backward of matrix_to_quaternion: dP from g
    // selected candidate row: num[4] / den, den = 2*max(qa, 0.1), qa = sqrt_pos(tr_sel)
    const float sgn[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
    const float tr = 1.0f + sgn[sel][0] * P[0] + sgn[sel][1] * P[4] + sgn[sel][2] * P[8];
    const float qa = sqrt_pos(tr);
    const float den = 2.0f * fmaxf(qa, 0.1f);
    const float gq[4] = {g.x, g.y, g.z, g.w};
    const float qv[4] = {qo.x, qo.y, qo.z, qo.w};
    // d/d(num_k) = g_k / den ; d/d(den) = -sum_k g_k * q_k / den
    float dnum[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) dnum[k] = gq[k] / den;
    float dden = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) dden -= gq[k] * qv[k] / den;
    // num_sel = qa^2 (= tr when tr > 0, else 0 with zero gradient);  den depends on qa only when qa > 0.1
    float dqa = 2.0f * qa * dnum[sel];
    if (qa > 0.1f) dqa += 2.0f * dden;
    const float dtr = (tr > 0.f) ? dqa * 0.5f / qa : 0.f;   // d sqrt(tr) = 1/(2 sqrt(tr))
    float dP[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) dP[e] = 0.f;
    dP[0] = sgn[sel][0] * dtr; dP[4] = sgn[sel][1] * dtr; dP[8] = sgn[sel][2] * dtr;
    // off-diagonal numerators (P index = 3*row + col): m21 = P[7], m12 = P[5], m02 = P[2], m20 = P[6], m10 = P[3], m01 = P[1]
    if (sel == 0) {       // (., m21-m12, m02-m20, m10-m01)
        dP[7] += dnum[1]; dP[5] -= dnum[1]; dP[2] += dnum[2]; dP[6] -= dnum[2]; dP[3] += dnum[3]; dP[1] -= dnum[3];
    } else if (sel == 1) { // (m21-m12, ., m10+m01, m02+m20)
        dP[7] += dnum[0]; dP[5] -= dnum[0]; dP[3] += dnum[2]; dP[1] += dnum[2]; dP[2] += dnum[3]; dP[6] += dnum[3];
    } else if (sel == 2) { // (m02-m20, m10+m01, ., m12+m21)
        dP[2] += dnum[0]; dP[6] -= dnum[0]; dP[3] += dnum[1]; dP[1] += dnum[1]; dP[5] += dnum[3]; dP[7] += dnum[3];
    } else {               // (m10-m01, m20+m02, m21+m12, .)
        dP[3] += dnum[0]; dP[1] -= dnum[0]; dP[6] += dnum[1]; dP[2] += dnum[1]; dP[7] += dnum[2]; dP[5] += dnum[2];
    }
    // ---- P = M R  ->  dR = M^T dP
    float dR[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            dR[3 * r + c] = M[0 + r] * dP[c] + M[3 + r] * dP[3 + c] + M[6 + r] * dP[6 + c];
    // ---- backward of quaternion_to_matrix: R = I + two_s * B(q), two_s = 2/|q|^2
    const float qr = q.x, qi = q.y, qj = q.z, qk = q.w;
    const float n2 = qr * qr + qi * qi + qj * qj + qk * qk;
    const float two_s = 2.0f / n2;
    // B entries (R = delta - two_s*(..) on the diagonal, two_s*(..) off-diagonal)
    const float B[9] = {-(qj * qj + qk * qk), qi * qj - qk * qr, qi * qk + qj * qr,
                        qi * qj + qk * qr, -(qi * qi + qk * qk), qj * qk - qi * qr,
                        qi * qk - qj * qr, qj * qk + qi * qr, -(qi * qi + qj * qj)};
    float dtwo_s = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) dtwo_s += dR[e] * B[e];
    float dB[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) dB[e] = dR[e] * two_s;
    float dqr = 0.f, dqi = 0.f, dqj = 0.f, dqk = 0.f;
    // B00 = -(j^2+k^2)
    dqj += -2 * qj * dB[0]; dqk += -2 * qk * dB[0];
    // B01 = ij - kr
    dqi += qj * dB[1]; dqj += qi * dB[1]; dqk += -qr * dB[1]; dqr += -qk * dB[1];
    // B02 = ik + jr
    dqi += qk * dB[2]; dqk += qi * dB[2]; dqj += qr * dB[2]; dqr += qj * dB[2];
    // B10 = ij + kr
    dqi += qj * dB[3]; dqj += qi * dB[3]; dqk += qr * dB[3]; dqr += qk * dB[3];
    // B11 = -(i^2+k^2)
    dqi += -2 * qi * dB[4]; dqk += -2 * qk * dB[4];
    // B12 = jk - ir
    dqj += qk * dB[5]; dqk += qj * dB[5]; dqi += -qr * dB[5]; dqr += -qi * dB[5];
    // B20 = ik - jr
    dqi += qk * dB[6]; dqk += qi * dB[6]; dqj += -qr * dB[6]; dqr += -qj * dB[6];
    // B21 = jk + ir
    dqj += qk * dB[7]; dqk += qj * dB[7]; dqi += qr * dB[7]; dqr += qi * dB[7];
    // B22 = -(i^2+j^2)
    dqi += -2 * qi * dB[8]; dqj += -2 * qj * dB[8];
    // two_s = 2/n2 -> d two_s / dq = -4 q / n2^2
    const float c2 = -4.0f / (n2 * n2) * dtwo_s;
    dqr += c2 * qr; dqi += c2 * qi; dqj += c2 * qj; dqk += c2 * qk;
    reinterpret_cast<float4*>(d_rot_in)[n] = make_float4(dqr, dqi, dqj, dqk);
}

static size_t lbs_smem_bytes(int J) {
    return (size_t)(((LBS_THREADS * J + 3) / 4) * 4 + J * 12) * sizeof(float);
}

}  // namespace agr

extern "C" {

int agr_lbs_forward(int32_t N, int32_t J, const float* weights, const float* jnt_mats, const float* xyz_in,
                    const float* rot_in, float* xyz_out, float* rot_out, float* pt_mats, void* cuda_stream) {
    using namespace agr;
    if (N < 0 || J < 1 || J > 256) return AGR_ERR_INVALID_ARGUMENT;
    if (N == 0) return AGR_OK;
    if (!weights || !jnt_mats || !xyz_in || !rot_in || !xyz_out || !rot_out) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t smem = lbs_smem_bytes(J);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(lbs_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lbs_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    lbs_fwd_kernel<true><<<(N + LBS_THREADS - 1) / LBS_THREADS, LBS_THREADS, smem, s>>>(N, J, weights, jnt_mats, xyz_in, rot_in,
                                                                                      xyz_out, rot_out, pt_mats);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_lbs_backward(int32_t N, const float* pt_mats, const float* rot_in, const float* d_xyz_out, const float* d_rot_out,
                     float* d_xyz_in, float* d_rot_in, void* cuda_stream) {
    using namespace agr;
    if (N < 0) return AGR_ERR_INVALID_ARGUMENT;
    if (N == 0) return AGR_OK;
    if (!pt_mats || !rot_in || !d_xyz_out || !d_rot_out || !d_xyz_in || !d_rot_in) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    lbs_bwd_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, pt_mats, rot_in, d_xyz_out, d_rot_out, d_xyz_in, d_rot_in);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

int agr_lbs_points(int32_t N, int32_t J, const float* weights, const float* jnt_mats, const float* xyz_in,
                   const float* vec_in, float* xyz_out, float* vec_out, void* cuda_stream) {
    using namespace agr;
    if (N < 0 || J < 1 || J > 256) return AGR_ERR_INVALID_ARGUMENT;
    if (N == 0) return AGR_OK;
    if (!weights || !jnt_mats || !xyz_in || !xyz_out || (vec_in && !vec_out)) return AGR_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t smem = lbs_smem_bytes(J);
    cudaFuncSetAttribute(lbs_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    lbs_fwd_kernel<false><<<(N + LBS_THREADS - 1) / LBS_THREADS, LBS_THREADS, smem, s>>>(N, J, weights, jnt_mats, xyz_in, vec_in,
                                                                                       xyz_out, vec_out, nullptr);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}

}  // extern "C"
