// SMPL-X joint chain (BASELINE config 0 plumbing): axis-angle -> rotation matrices (Rodrigues) and the kinematic-tree
// composition that yields the per-joint affine matrices A consumed as `cano2live_jnt_mats`.
// Reference: smplx/lbs.py:300-336 (batch_rodrigues) and :349-405 (batch_rigid_transform), which walks the 55 joints in a
// Python loop of 4x4 matmuls (~110 ATen launches).  Here: one CTA, one thread per joint for Rodrigues and the local
// transforms, then a single thread composes the chain (55 x 4x4 products) out of shared memory.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/agr_rasterizer.h"
#include "../../include/agr_lbs.h"

namespace agr {

__global__ void smpl_chain_kernel(int J, const float* __restrict__ pose, int pose_is_rotmat, const float* __restrict__ joints,
                                  const int32_t* __restrict__ parents, float* __restrict__ rot_out, float* __restrict__ posed_joints,
                                  float* __restrict__ A) {
    extern __shared__ float sm[];          // local[J][16], chain[J][16]
    float* local = sm;
    float* chain = sm + 16 * J;
    const int j = threadIdx.x;
    if (j < J) {
        float R[9];
        if (pose_is_rotmat) {
            for (int e = 0; e < 9; ++e) R[e] = pose[9 * j + e];
        } else {
            // batch_rodrigues: angle = |r + 1e-8| (the epsilon is added to every component), dir = r / angle
            const float rx0 = pose[3 * j], ry0 = pose[3 * j + 1], rz0 = pose[3 * j + 2];
            const float ax = rx0 + 1e-8f, ay = ry0 + 1e-8f, az = rz0 + 1e-8f;
            const float angle = sqrtf(ax * ax + ay * ay + az * az);
            const float rx = rx0 / angle, ry = ry0 / angle, rz = rz0 / angle;
            const float c = cosf(angle), s = sinf(angle);
            const float K[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
            float KK[9];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) KK[3 * a + b] = K[3 * a] * K[b] + K[3 * a + 1] * K[3 + b] + K[3 * a + 2] * K[6 + b];
            for (int e = 0; e < 9; ++e) R[e] = ((e % 4 == 0) ? 1.f : 0.f) + s * K[e] + (1.f - c) * KK[e];
        }
        if (rot_out) for (int e = 0; e < 9; ++e) rot_out[9 * j + e] = R[e];
        // local transform [R | rel_joint; 0 0 0 1], rel_joint = joint - joint[parent] (root: joint itself)
        float t[3];
        const int par = parents[j];
        for (int a = 0; a < 3; ++a) t[a] = joints[3 * j + a] - ((j > 0) ? joints[3 * par + a] : 0.f);
        float* L = local + 16 * j;
        for (int a = 0; a < 3; ++a) { L[4 * a] = R[3 * a]; L[4 * a + 1] = R[3 * a + 1]; L[4 * a + 2] = R[3 * a + 2]; L[4 * a + 3] = t[a]; }
        L[12] = 0.f; L[13] = 0.f; L[14] = 0.f; L[15] = 1.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int e = 0; e < 16; ++e) chain[e] = local[e];
        for (int i = 1; i < J; ++i) {
            const float* P = chain + 16 * parents[i];
            const float* L = local + 16 * i;
            float* C = chain + 16 * i;
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) C[4 * a + b] = P[4 * a] * L[b] + P[4 * a + 1] * L[4 + b] + P[4 * a + 2] * L[8 + b] + P[4 * a + 3] * L[12 + b];
        }
    }
    __syncthreads();
    if (j < J) {
        const float* C = chain + 16 * j;
        for (int a = 0; a < 3; ++a) posed_joints[3 * j + a] = C[4 * a + 3];
        // rel_transforms = T - pad(T @ [joint; 0]) into the last column
        const float jx = joints[3 * j], jy = joints[3 * j + 1], jz = joints[3 * j + 2];
        for (int a = 0; a < 4; ++a) {
            const float corr = C[4 * a] * jx + C[4 * a + 1] * jy + C[4 * a + 2] * jz;
            A[16 * j + 4 * a + 0] = C[4 * a + 0];
            A[16 * j + 4 * a + 1] = C[4 * a + 1];
            A[16 * j + 4 * a + 2] = C[4 * a + 2];
            A[16 * j + 4 * a + 3] = C[4 * a + 3] - corr;
        }
    }
}

}  // namespace agr

extern "C" int agr_smpl_joint_chain(int32_t J, const float* pose, int32_t pose_is_rotmat, const float* joints, const int32_t* parents,
                                    float* rot_mats_out, float* posed_joints, float* A, void* cuda_stream) {
    if (J < 1 || J > 1024 || !pose || !joints || !parents || !posed_joints || !A) return AGR_ERR_INVALID_ARGUMENT;
    const int threads = ((J + 31) / 32) * 32;
    agr::smpl_chain_kernel<<<1, threads, (size_t)32 * J * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
        J, pose, pose_is_rotmat, joints, parents, rot_mats_out, posed_joints, A);
    return cudaGetLastError() == cudaSuccess ? AGR_OK : AGR_ERR_CUDA;
}
