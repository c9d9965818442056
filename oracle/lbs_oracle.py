"""TEST INFRASTRUCTURE — CPU restatement (PyTorch tensor ops, any float dtype) of
AvatarNet.transform_cano2live (network/avatar.py:84-91) including the two pytorch3d functions it calls.

PARITY UNPINNED for the quaternion functions: pytorch3d == 0.7.4 (reference requirements.txt:9) is a
third-party dependency that is neither vendored in /root/reference nor installed here, and the reference
has no tests or golden vectors for this path.  quaternion_to_matrix / matrix_to_quaternion below restate the
published pytorch3d 0.7.x algorithm (pytorch3d/transforms/rotation_conversions.py): real-first quaternions,
q2m divides by |q|^2, m2q selects the best-conditioned of four candidates with a 0.1 floor on the divisor
and does NOT standardise the sign (added only in later releases).  Gradients come from autograd, exactly as
in the reference."""
import torch
import torch.nn.functional as F


def quaternion_to_matrix(quaternions):
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    positive_mask = x > 0
    ret[positive_mask] = torch.sqrt(x[positive_mask])
    return ret


def matrix_to_quaternion(matrix):
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                             1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    return quat_candidates[F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch_dim + (4,))


def transform_cano2live(lbs, jnt_mats, positions, rotations):
    """network/avatar.py:84-91"""
    pt_mats = torch.einsum('nj,jxy->nxy', lbs, jnt_mats)
    positions = torch.einsum('nxy,ny->nx', pt_mats[..., :3, :3], positions) + pt_mats[..., :3, 3]
    rot_mats = quaternion_to_matrix(rotations)
    rot_mats = torch.einsum('nxy,nyz->nxz', pt_mats[..., :3, :3], rot_mats)
    return positions, matrix_to_quaternion(rot_mats)


def skin_points(lbs, jnt_mats, points, normals=None):
    """network/avatar.py:128-130 / 150-151"""
    pt_mats = torch.einsum('nj,jxy->nxy', lbs, jnt_mats)
    live = torch.einsum('nxy,ny->nx', pt_mats[..., :3, :3], points) + pt_mats[..., :3, 3]
    if normals is None:
        return live
    return live, torch.einsum('nxy,ny->nx', pt_mats[..., :3, :3], normals)
