"""Camera setup of the render path — host-side mirror of the reference's
gaussians/gaussian_renderer.py:44-66 (render3) and utils/graphics_utils.py:51-85
(getProjectionMatrix, focal2fov).

The reference pulls intr[0,0].item() / intr[1,1].item() back to the host per call (two syncs,
gaussian_renderer.py:45-46); here camera parameters are taken from host copies (numpy / python floats)
and every device matrix of a whole view batch is produced with ONE pinned host->device copy.
"""
import math

import numpy as np
import torch

from .rasterizer import BatchedRasterizationSettings, GaussianRasterizationSettings


def focal2fov(focal, pixels):
    """utils/graphics_utils.py:84-85"""
    return 2 * math.atan(pixels / (2 * focal))


def get_projection_matrix(znear, zfar, K, img_w, img_h):
    """K-aware off-centre projection, utils/graphics_utils.py:51-79 (K is not None branch), float32 like
    the reference's torch.zeros(4,4)."""
    K = np.asarray(K, dtype=np.float32)
    near_fx = np.float32(znear) / K[0, 0]
    near_fy = np.float32(znear) / K[1, 1]
    left = -(np.float32(img_w) - K[0, 2]) * near_fx
    right = K[0, 2] * near_fx
    bottom = (K[1, 2] - np.float32(img_h)) * near_fy
    top = K[1, 2] * near_fy
    P = np.zeros((4, 4), np.float32)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_block(extr, intr, img_w, img_h, znear=0.1, zfar=100.0):
    """Host (numpy) computation of everything render3 derives from one (extr, intr):
    returns dict(tanfovx, tanfovy, viewmatrix(4,4), projmatrix(4,4), campos(3))
    in the reference's transposed layout (gaussian_renderer.py:49-52)."""
    extr = np.asarray(extr, dtype=np.float32)
    intr = np.asarray(intr, dtype=np.float32)
    fovx = focal2fov(float(intr[0, 0]), img_w)
    fovy = focal2fov(float(intr[1, 1]), img_h)
    world_view = extr.T.copy()
    proj = get_projection_matrix(znear, zfar, intr, img_w, img_h).T
    full_proj = (world_view @ proj).astype(np.float32)
    campos = np.linalg.inv(extr)[:3, 3].astype(np.float32)
    return dict(tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5), viewmatrix=world_view,
                projmatrix=full_proj, campos=campos)


def make_raster_settings(extr, intr, img_w, img_h, bg_color, device, scaling_modifier=1.0, sh_degree=0):
    cb = camera_block(_np(extr), _np(intr), img_w, img_h)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return GaussianRasterizationSettings(
        image_height=int(img_h), image_width=int(img_w), tanfovx=cb["tanfovx"], tanfovy=cb["tanfovy"],
        bg=bg_color, scale_modifier=scaling_modifier, viewmatrix=t(cb["viewmatrix"]), projmatrix=t(cb["projmatrix"]),
        sh_degree=sh_degree, campos=t(cb["campos"]), prefiltered=False, debug=False)


def make_batched_settings(extrs, intrs, img_w, img_h, bg_color, device, scaling_modifier=1.0, sh_degree=0):
    """V cameras -> one BatchedRasterizationSettings; matrices travel in one pinned H2D copy."""
    blocks = [camera_block(_np(e), _np(k), img_w, img_h) for e, k in zip(extrs, intrs)]
    V = len(blocks)
    host = torch.empty((V, 35), dtype=torch.float32, pin_memory=torch.cuda.is_available())
    for v, cb in enumerate(blocks):
        host[v, :16] = torch.from_numpy(cb["viewmatrix"].reshape(-1))
        host[v, 16:32] = torch.from_numpy(cb["projmatrix"].reshape(-1))
        host[v, 32:35] = torch.from_numpy(cb["campos"])
    dev = host.to(device, non_blocking=True)
    return BatchedRasterizationSettings(
        image_height=int(img_h), image_width=int(img_w), tanfovx=[b["tanfovx"] for b in blocks],
        tanfovy=[b["tanfovy"] for b in blocks], bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=dev[:, :16].reshape(V, 4, 4).contiguous(), projmatrix=dev[:, 16:32].reshape(V, 4, 4).contiguous(),
        sh_degree=sh_degree, campos=dev[:, 32:35].contiguous(), prefiltered=False, debug=False)


def _np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)
