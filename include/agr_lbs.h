/*
 * agr_lbs.h — C ABI of the fused per-Gaussian linear-blend-skinning kernel.
 *
 * Replaces AvatarNet.transform_cano2live (network/avatar.py:84-91), which the reference runs as
 * three torch.einsum + pytorch3d.transforms.quaternion_to_matrix / matrix_to_quaternion
 * (~12 ATen launches and an (N,4,4) intermediate):
 *     pt_mats   = einsum('nj,jxy->nxy', lbs, cano2live_jnt_mats)
 *     positions = pt_mats[:, :3, :3] @ positions + pt_mats[:, :3, 3]
 *     rotations = matrix_to_quaternion(pt_mats[:, :3, :3] @ quaternion_to_matrix(rotations))
 * pytorch3d == 0.7.4 is pinned by the reference's requirements.txt:9 but is not vendored; its two
 * functions are restated from the published source (real-first quaternions, no sign standardisation).
 *
 * One kernel does blend-weight reduce + rotation build + compose + back-conversion, forward and
 * backward.  All pointers are device pointers owned by the caller; stream passed as void*.
 * Returns 0 on success, else AgrStatus (agr_rasterizer.h).
 */
#ifndef AGR_LBS_H_
#define AGR_LBS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* forward.
 *   weights   (N,J) fp32 dense LBS weights      (AvatarNet.lbs, avatar.py:31)
 *   jnt_mats  (J,4,4) fp32 row-major cano->live  (items['cano2live_jnt_mats'])
 *   xyz_in    (N,3), rot_in (N,4) real-first quaternion (need not be unit)
 *   xyz_out   (N,3), rot_out (N,4)
 *   pt_mats   (N,12) OUT: rows 0..2 of the blended matrix (saved for backward; may be NULL at inference) */
int agr_lbs_forward(int32_t N, int32_t J, const float* weights, const float* jnt_mats,
                    const float* xyz_in, const float* rot_in,
                    float* xyz_out, float* rot_out, float* pt_mats, void* cuda_stream);

/* backward w.r.t. xyz_in and rot_in (weights / joint matrices are data, not parameters).
 *   pt_mats (N,12) from the forward, rot_in (N,4), d_xyz_out (N,3), d_rot_out (N,4)
 *   d_xyz_in (N,3), d_rot_in (N,4) are overwritten. */
int agr_lbs_backward(int32_t N, const float* pt_mats, const float* rot_in,
                     const float* d_xyz_out, const float* d_rot_out,
                     float* d_xyz_in, float* d_rot_in, void* cuda_stream);

/* positions / normals only (no rotations): get_viewdir_feat and get_pose_map
 * (avatar.py:126-159) skin init points and normals with the same blended matrices.
 *   vec_in (N,3) or NULL: transformed with the 3x3 part only (normals). */
int agr_lbs_points(int32_t N, int32_t J, const float* weights, const float* jnt_mats,
                   const float* xyz_in, const float* vec_in, float* xyz_out, float* vec_out,
                   void* cuda_stream);

/* SMPL-X joint chain (smplx/lbs.py:300-336 batch_rodrigues + :349-405 batch_rigid_transform), batch 1.
 *   pose: (J,3) axis-angle, or (J,9) rotation matrices when pose_is_rotmat != 0; joints (J,3) rest joints;
 *   parents (J) int32 kinematic tree (parents[0] ignored).  Outputs: rot_mats_out (J,9) or NULL,
 *   posed_joints (J,3), A (J,4,4) relative transforms (what dataset_mv_rgb.py:172 turns into cano2live_jnt_mats). */
int agr_smpl_joint_chain(int32_t J, const float* pose, int32_t pose_is_rotmat, const float* joints, const int32_t* parents,
                         float* rot_mats_out, float* posed_joints, float* A, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
