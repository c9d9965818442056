"""SMPL-X linear blend skinning — host-side mirror of the reference's smplx/lbs.py:152-252 (`lbs`), the CPU plumbing of
BASELINE config 0 that supplies the joint affine matrices `A` (-> cano2live_jnt_mats).

Same signature and return values as the reference function.  The shape/pose blend-shape contractions are two small
GEMVs (cuBLAS through torch); the kinematic chain is ONE kernel (agr_smpl_joint_chain, the reference loops over the 55
joints in Python) and the vertex skinning reuses the fused LBS kernel (agr_lbs_points).  Batch size 1, like the avatar."""
import ctypes as C

import torch

from . import _lib, lbs as _lbs, stats

_p = C.c_void_p
_lib.register_symbols({
    "agr_smpl_joint_chain": (C.c_int, [C.c_int32, _p, C.c_int32, _p, _p, _p, _p, _p, _p]),
})


def joint_chain(pose, joints, parents, pose2rot=True):
    """pose (J,3) axis-angle (or (J,3,3) rotation matrices), joints (J,3), parents (J) -> rot_mats (J,3,3),
    posed_joints (J,3), A (J,4,4): batch_rodrigues + batch_rigid_transform (smplx/lbs.py:300-405)."""
    lib = _lib.load()
    dev = joints.device
    J = joints.shape[0]
    po = pose.detach().float().contiguous()
    jt = joints.detach().float().contiguous()
    pa = parents.to(device=dev, dtype=torch.int32).contiguous()
    rot = torch.empty((J, 3, 3), dtype=torch.float32, device=dev)
    pj = torch.empty((J, 3), dtype=torch.float32, device=dev)
    A = torch.empty((J, 4, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), stats.stage("lbs", launches=1):
        st = lib.agr_smpl_joint_chain(J, C.c_void_p(po.data_ptr()), int(not pose2rot), C.c_void_p(jt.data_ptr()),
                                      C.c_void_p(pa.data_ptr()), C.c_void_p(rot.data_ptr()), C.c_void_p(pj.data_ptr()),
                                      C.c_void_p(A.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if st != _lib.AGR_OK:
        raise RuntimeError("agr_smpl_joint_chain failed: %d" % st)
    return rot, pj, A


@torch.no_grad()
def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True,
        return_affine_mat=False):
    """smplx/lbs.py:152-252. betas (1,NB), pose (1,(J)*3) [or (1,J,9|3,3) with pose2rot=False], v_template (1,V,3) or (V,3),
    shapedirs (V,3,NB), posedirs (P,V*3), J_regressor (J,V), parents (J), lbs_weights (V,J).
    Returns verts (1,V,3), joints (1,J,3) [, A (1,J,4,4)]."""
    if max(betas.shape[0], pose.shape[0]) != 1:
        raise RuntimeError("animatablegaussians_b200.smpl_lbs.lbs handles batch size 1 (one pose per step)")
    if not betas.is_cuda:
        raise RuntimeError("CUDA tensors required (no CPU fallback)")
    vt = v_template.reshape(-1, 3)
    v_shaped = vt + torch.einsum('l,mkl->mk', betas[0], shapedirs)              # blend_shapes, lbs.py:279-297
    J = torch.einsum('jv,vk->jk', J_regressor, v_shaped)                        # vertices2joints, lbs.py:255-276
    nj = J_regressor.shape[0]
    rot, posed_joints, A = joint_chain(pose.reshape(nj, -1) if pose2rot else pose.reshape(nj, 9), J, parents, pose2rot)
    ident = torch.eye(3, dtype=rot.dtype, device=rot.device)
    pose_feature = (rot[1:] - ident).reshape(1, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(-1, 3)
    verts = _lbs.skin_points(lbs_weights, A, v_posed)
    if return_affine_mat:
        return verts[None], posed_joints[None], A[None]
    return verts[None], posed_joints[None]
