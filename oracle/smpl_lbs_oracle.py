"""TEST INFRASTRUCTURE — restatement of smplx/lbs.py:152-405 (lbs, blend_shapes, vertices2joints, batch_rodrigues,
batch_rigid_transform) in plain PyTorch, any float dtype, batch 1.  PINNED by tests/golden/smpl_lbs.npz, produced by
importing the UNMODIFIED reference module (tests/golden/make_smpl_lbs_golden.py)."""
import torch
import torch.nn.functional as F


def batch_rodrigues(rot_vecs):
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(d, 1, dim=1)
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    J = joints.shape[0]
    rel = joints.clone()
    rel[1:] -= joints[parents[1:]]
    T = torch.zeros(J, 4, 4, dtype=joints.dtype)
    T[:, :3, :3] = rot_mats
    T[:, :3, 3] = rel
    T[:, 3, 3] = 1
    chain = [T[0]]
    for i in range(1, J):
        chain.append(chain[int(parents[i])] @ T[i])
    tr = torch.stack(chain, 0)
    posed = tr[:, :3, 3]
    jh = F.pad(joints, [0, 1])[..., None]
    rel_tr = tr - F.pad(tr @ jh, [3, 0])
    return posed, rel_tr


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    v_shaped = v_template + torch.einsum('l,mkl->mk', betas[0], shapedirs)
    J = J_regressor @ v_shaped
    rot = batch_rodrigues(pose.view(-1, 3))
    pose_feature = (rot[1:] - torch.eye(3, dtype=rot.dtype)).reshape(1, -1)
    v_posed = v_shaped + (pose_feature @ posedirs).view(-1, 3)
    posed, A = batch_rigid_transform(rot, J, parents)
    Tm = (lbs_weights @ A.view(-1, 16)).view(-1, 4, 4)
    vh = F.pad(v_posed, [0, 1], value=1.0)
    verts = torch.einsum('vab,vb->va', Tm, vh)[:, :3]
    return verts, posed, A
