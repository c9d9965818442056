"""Host-side mirror of AvatarNet.transform_cano2live (network/avatar.py:84-91): one fused sm_100a kernel
(include/agr_lbs.h) instead of 3 einsums + pytorch3d quaternion round trip."""
import ctypes as C

import torch

from . import _lib, stats

_p = C.c_void_p
_lib.register_symbols({
    "agr_lbs_forward": (C.c_int, [C.c_int32, C.c_int32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "agr_lbs_backward": (C.c_int, [C.c_int32, _p, _p, _p, _p, _p, _p, _p]),
    "agr_lbs_points": (C.c_int, [C.c_int32, C.c_int32, _p, _p, _p, _p, _p, _p, _p]),
})


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check(st, what):
    if st != _lib.AGR_OK:
        raise RuntimeError("%s failed (status %d): %s" % (what, st, _lib.cuda_error_string() if st == _lib.AGR_ERR_CUDA else ""))


class _LBS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, jnt_mats, positions, rotations):
        lib = _lib.load()
        dev = positions.device
        N, J = weights.shape
        w = weights.detach().float().contiguous()
        A = jnt_mats.detach().float().contiguous()
        x = positions.detach().float().contiguous()
        q = rotations.detach().float().contiguous()
        xo = torch.empty_like(x)
        qo = torch.empty_like(q)
        need_bwd = positions.requires_grad or rotations.requires_grad
        pt = torch.empty((N, 12), dtype=torch.float32, device=dev) if need_bwd else None
        with torch.cuda.device(dev), stats.stage("lbs", launches=1):
            _check(lib.agr_lbs_forward(N, J, _ptr(w), _ptr(A), _ptr(x), _ptr(q), _ptr(xo), _ptr(qo), _ptr(pt), _stream(dev)),
                   "agr_lbs_forward")
        if need_bwd:
            ctx.save_for_backward(pt, q)
        return xo, qo

    @staticmethod
    def backward(ctx, gx, gq):
        lib = _lib.load()
        pt, q = ctx.saved_tensors
        dev = q.device
        N = q.shape[0]
        gx = torch.zeros((N, 3), dtype=torch.float32, device=dev) if gx is None else gx.float().contiguous()
        gq = torch.zeros((N, 4), dtype=torch.float32, device=dev) if gq is None else gq.float().contiguous()
        dx = torch.empty((N, 3), dtype=torch.float32, device=dev)
        dq = torch.empty((N, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), stats.stage("lbs", launches=1):
            _check(lib.agr_lbs_backward(N, _ptr(pt), _ptr(q), _ptr(gx), _ptr(gq), _ptr(dx), _ptr(dq), _stream(dev)),
                   "agr_lbs_backward")
        return None, None, dx, dq


def transform_cano2live(lbs_weights, cano2live_jnt_mats, positions, rotations):
    """positions (N,3), rotations (N,4) real-first -> posed positions, posed rotations."""
    return _LBS.apply(lbs_weights, cano2live_jnt_mats, positions, rotations)


def skin_points(lbs_weights, jnt_mats, points, normals=None):
    """No-grad LBS of points (and optionally normals with the 3x3 part): avatar.py:128-130,150-151."""
    lib = _lib.load()
    dev = points.device
    N, J = lbs_weights.shape
    w = lbs_weights.detach().float().contiguous()
    A = jnt_mats.detach().float().contiguous()
    x = points.detach().float().contiguous()
    v = normals.detach().float().contiguous() if normals is not None else None
    xo = torch.empty_like(x)
    vo = torch.empty_like(v) if v is not None else None
    with torch.cuda.device(dev), stats.stage("lbs", launches=1):
        _check(lib.agr_lbs_points(N, J, _ptr(w), _ptr(A), _ptr(x), _ptr(v), _ptr(xo), _ptr(vo), _stream(dev)), "agr_lbs_points")
    return xo if normals is None else (xo, vo)
