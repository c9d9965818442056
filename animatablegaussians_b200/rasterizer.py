"""Host-side mirror of the reference package `diff_gaussian_rasterization_depth_alpha`
(RAST/diff_gaussian_rasterization_depth_alpha/__init__.py:21-223; RAST =
gaussians/diff_gaussian_rasterization_depth_alpha in the reference tree).

Same public names, argument meaning and error behaviour:
    rasterize_gaussians, GaussianRasterizationSettings, GaussianRasterizer(.forward/.markVisible)
plus the view-batched entry the reference lacks:
    rasterize_gaussians_batched / BatchedRasterizationSettings
All arithmetic happens in hand-written sm_100a CUDA behind the C ABI of include/agr_rasterizer.h;
PyTorch only owns memory, streams and the autograd wiring.  No CPU / PyTorch fallback exists.
"""
import ctypes as C
from typing import NamedTuple, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, stats

_EMPTY = None


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _f32c(t: Optional[torch.Tensor], device):
    """contiguous float32 on `device`; empty / None -> None (the reference's 'absent' empty tensors)."""
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class BatchedRasterizationSettings(NamedTuple):
    """V views of identical image size. viewmatrix/projmatrix: (V,4,4) in the reference's
    transposed layout; campos (V,3); bg (3,) or (V,3); tanfovx/tanfovy: sequences of V floats."""
    image_height: int
    image_width: int
    tanfovx: Sequence[float]
    tanfovy: Sequence[float]
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    capacity: Optional[int] = None  # fixed instance capacity -> sync-free (CUDA-graph capturable) forward


# device int64[2] = (instances emitted, overflow flag) of the most recent sync-free forward; check it after the
# graph replay / at the end of the step:  if int(last_device_status[1]): capacity was too small.
last_device_status = None


# capacity (in tile instances) that worked last time for a given problem shape
_capacity_hint = {}


def _raise_status(st, what):
    if st == _lib.AGR_ERR_CUDA:
        raise RuntimeError("%s: CUDA error: %s" % (what, _lib.cuda_error_string()))
    names = {1: "invalid argument", 2: "binning capacity", 4: "workspace too small"}
    raise RuntimeError("%s failed: %s" % (what, names.get(st, str(st))))


class _Ctx:
    pass


def _forward_impl(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, V):
    """Shared by the per-view drop-in (V = 1) and the batched entry. Returns outputs + saved state."""
    lib = _lib.load()
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (this build has no CPU path)")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)

    means3D = _f32c(means3D, dev) if P > 0 else means3D
    sh = _f32c(sh, dev)
    colors_precomp = _f32c(colors_precomp, dev)
    opacities = _f32c(opacities, dev)
    scales = _f32c(scales, dev)
    rotations = _f32c(rotations, dev)
    cov3Ds_precomp = _f32c(cov3Ds_precomp, dev)
    bg = _f32c(rs.bg, dev)
    view = _f32c(rs.viewmatrix, dev)
    proj = _f32c(rs.projmatrix, dev)
    campos = _f32c(rs.campos, dev)

    M = sh.shape[1] if sh is not None else 0
    colors_stride = 0
    if colors_precomp is not None and colors_precomp.ndim == 3:
        if colors_precomp.shape[0] != V:
            raise RuntimeError("per-view colors_precomp must have shape (V, P, 3)")
        colors_stride = P * 3
    bg_stride = 3 if (bg is not None and bg.ndim == 2) else 0

    if isinstance(rs.tanfovx, (float, int)):
        tfx, tfy = [float(rs.tanfovx)], [float(rs.tanfovy)]
    else:
        tfx, tfy = [float(x) for x in rs.tanfovx], [float(y) for y in rs.tanfovy]
    if len(tfx) != V or len(tfy) != V:
        raise RuntimeError("tanfovx/tanfovy must have one entry per view")
    tfx_c = (C.c_float * V)(*tfx)
    tfy_c = (C.c_float * V)(*tfy)

    out_color = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev)
    out_depth = torch.empty((V, 1, H, W), dtype=torch.float32, device=dev)
    out_alpha = torch.empty((V, 1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((V, max(P, 0)), dtype=torch.int32, device=dev)

    key = (P, V, H, W)
    fixed_capacity = getattr(rs, "capacity", None)
    sync_free = fixed_capacity is not None
    capacity = int(fixed_capacity) if sync_free else _capacity_hint.get(key, max(4 * P * V, 1 << 16))
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    num_rendered = C.c_int64(0)
    status = torch.empty(2, dtype=torch.int64, device=dev) if sync_free else None
    ws = _lib.AgrRasterWorkspace()
    geom = image = binning = None
    with torch.cuda.device(dev):
        for _attempt in range(3):
            st = lib.agr_raster_workspace(P, V, W, H, M, capacity, C.byref(ws))
            if st != _lib.AGR_OK:
                _raise_status(st, "agr_raster_workspace")
            if geom is None:
                geom = torch.empty(ws.geom_bytes, dtype=torch.uint8, device=dev)
                image = torch.empty(ws.image_bytes, dtype=torch.uint8, device=dev)
            binning = torch.empty(ws.binning_bytes, dtype=torch.uint8, device=dev)
            a = _lib.AgrRasterForwardArgs()
            a.P, a.V, a.width, a.height = P, V, W, H
            a.sh_degree, a.sh_coeffs = int(rs.sh_degree), M
            a.scale_modifier = float(rs.scale_modifier)
            a.prefiltered, a.debug = int(bool(rs.prefiltered)), int(bool(rs.debug))
            a.background, a.bg_view_stride = _ptr(bg), bg_stride
            a.means3D, a.shs = _ptr(means3D), _ptr(sh)
            a.colors_precomp, a.colors_view_stride = _ptr(colors_precomp), colors_stride
            a.opacities, a.scales, a.rotations = _ptr(opacities), _ptr(scales), _ptr(rotations)
            a.cov3D_precomp = _ptr(cov3Ds_precomp)
            a.viewmatrix, a.projmatrix, a.campos = _ptr(view), _ptr(proj), _ptr(campos)
            a.tan_fovx = C.cast(tfx_c, C.POINTER(C.c_float))
            a.tan_fovy = C.cast(tfy_c, C.POINTER(C.c_float))
            a.out_color, a.out_depth, a.out_alpha, a.radii = _ptr(out_color), _ptr(out_depth), _ptr(out_alpha), _ptr(radii)
            a.geom_ws, a.geom_bytes = _ptr(geom), ws.geom_bytes
            a.image_ws, a.image_bytes = _ptr(image), ws.image_bytes
            a.binning_ws, a.binning_bytes = _ptr(binning), ws.binning_bytes
            a.capacity = capacity
            a.num_rendered = None if sync_free else C.pointer(num_rendered)
            a.device_status = _ptr(status)
            with stats.stage("raster_fwd", launches=4):
                st = lib.agr_raster_forward(C.byref(a), stream)
            if st == _lib.AGR_ERR_BINNING_CAPACITY:
                if num_rendered.value >= (1 << 31) or num_rendered.value > 256 * max(P, 1) * V:
                    raise RuntimeError("rasterizer: %d (tile, Gaussian) instances — degenerate input (screen-filling "
                                       "Gaussians); refusing to allocate the binning workspace" % num_rendered.value)
                capacity = int(num_rendered.value * 1.25) + 1024
                continue
            break
        if st != _lib.AGR_OK:
            _raise_status(st, "agr_raster_forward")
    if sync_free:
        global last_device_status
        last_device_status = status
        R = -1
    else:
        R = int(num_rendered.value)
        _capacity_hint[key] = max(int(R * 1.15) + 1024, 1 << 12)

    saved = _Ctx()
    saved.tensors = (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, image,
                     out_alpha, bg, view, proj, campos)
    saved.meta = dict(P=P, V=V, H=H, W=W, M=M, capacity=capacity, R=R, tfx=tfx, tfy=tfy,
                      colors_stride=colors_stride, bg_stride=bg_stride, sh_degree=int(rs.sh_degree),
                      scale_modifier=float(rs.scale_modifier), debug=bool(rs.debug), ws=(ws.geom_bytes, ws.image_bytes,
                                                                                        ws.binning_bytes, ws.backward_bytes))
    return out_color, radii, out_depth, out_alpha, saved


def _backward_impl(saved_tensors, meta, grad_color, grad_depth, grad_alpha):
    lib = _lib.load()
    (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, image, out_alpha, bg, view,
     proj, campos) = saved_tensors
    P, V, H, W, M = meta["P"], meta["V"], meta["H"], meta["W"], meta["M"]
    dev = means3D.device

    def _g(t, shape):
        if t is None:
            return torch.zeros(shape, dtype=torch.float32, device=dev)
        return t.to(torch.float32).contiguous()

    grad_color = _g(grad_color, (V, 3, H, W))
    grad_depth = _g(grad_depth, (V, 1, H, W))
    grad_alpha = _g(grad_alpha, (V, 1, H, W))

    per_view_colors = meta["colors_stride"] != 0
    dmeans3D = torch.empty((P, 3), dtype=torch.float32, device=dev)
    dmeans2D = torch.empty((V, P, 3), dtype=torch.float32, device=dev)
    dcolors = torch.empty(((V, P, 3) if per_view_colors else (P, 3)), dtype=torch.float32, device=dev) \
        if colors_precomp is not None else None
    dopacity = torch.empty((P, 1), dtype=torch.float32, device=dev)
    dcov3D = torch.empty((P, 6), dtype=torch.float32, device=dev) if cov3Ds_precomp is not None else None
    dsh = torch.empty((P, M, 3), dtype=torch.float32, device=dev) if sh is not None else None
    dscales = torch.empty((P, 3), dtype=torch.float32, device=dev) if scales is not None else None
    drot = torch.empty((P, 4), dtype=torch.float32, device=dev) if rotations is not None else None
    if P == 0:
        return dmeans3D, dmeans2D, dsh, dcolors, dopacity, dscales, drot, dcov3D

    gb, ib, bb, bwb = meta["ws"]
    bwd_ws = torch.empty(bwb, dtype=torch.uint8, device=dev)
    tfx_c = (C.c_float * V)(*meta["tfx"])
    tfy_c = (C.c_float * V)(*meta["tfy"])
    a = _lib.AgrRasterBackwardArgs()
    a.P, a.V, a.width, a.height = P, V, W, H
    a.sh_degree, a.sh_coeffs = meta["sh_degree"], M
    a.scale_modifier, a.debug = meta["scale_modifier"], int(meta["debug"])
    a.background, a.bg_view_stride = _ptr(bg), meta["bg_stride"]
    a.means3D, a.shs = _ptr(means3D), _ptr(sh)
    a.colors_precomp, a.colors_view_stride = _ptr(colors_precomp), meta["colors_stride"]
    a.scales, a.rotations, a.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp)
    a.viewmatrix, a.projmatrix, a.campos = _ptr(view), _ptr(proj), _ptr(campos)
    a.tan_fovx = C.cast(tfx_c, C.POINTER(C.c_float))
    a.tan_fovy = C.cast(tfy_c, C.POINTER(C.c_float))
    a.radii, a.out_alpha = _ptr(radii), _ptr(out_alpha)
    a.dL_dout_color, a.dL_dout_depth, a.dL_dout_alpha = _ptr(grad_color), _ptr(grad_depth), _ptr(grad_alpha)
    a.dL_dmeans3D, a.dL_dmeans2D, a.dL_dcolors = _ptr(dmeans3D), _ptr(dmeans2D), _ptr(dcolors)
    a.dL_dopacity, a.dL_dcov3D, a.dL_dsh = _ptr(dopacity), _ptr(dcov3D), _ptr(dsh)
    a.dL_dscales, a.dL_drotations = _ptr(dscales), _ptr(drot)
    a.geom_ws, a.geom_bytes = _ptr(geom), gb
    a.image_ws, a.image_bytes = _ptr(image), ib
    a.binning_ws, a.binning_bytes = _ptr(binning), bb
    a.backward_ws, a.backward_bytes = _ptr(bwd_ws), bwb
    a.capacity, a.num_rendered = meta["capacity"], meta["R"]
    with torch.cuda.device(dev), stats.stage("raster_bwd", launches=2):
        st = lib.agr_raster_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if st != _lib.AGR_OK:
        _raise_status(st, "agr_raster_backward")
    return dmeans3D, dmeans2D, dsh, dcolors, dopacity, dscales, drot, dcov3D


# ------------------------------------------------------------------------------------------------
# Reference surface (one view per call) — __init__.py:21-158
# ------------------------------------------------------------------------------------------------
def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args[:-1])  # copy before they can be corrupted (__init__.py:83-90)
            try:
                color, radii, depth, alpha, saved = _forward_impl(*args, 1)
            except Exception:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            color, radii, depth, alpha, saved = _forward_impl(*args, 1)
        ctx.meta = saved.meta
        ctx.none_mask = [t is None for t in saved.tensors]
        ctx.save_for_backward(*[t if t is not None else torch.empty(0) for t in saved.tensors])
        return color[0], radii[0], depth[0], alpha[0]

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        tensors = tuple(None if m else t for t, m in zip(ctx.saved_tensors, ctx.none_mask))
        meta = ctx.meta
        gc = grad_color[None] if grad_color is not None else None
        gd = grad_depth[None] if grad_depth is not None else None
        ga = grad_alpha[None] if grad_alpha is not None else None
        if meta["debug"]:
            try:
                out = _backward_impl(tensors, meta, gc, gd, ga)
            except Exception:
                torch.save(cpu_deep_copy_tuple(tensors), "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            out = _backward_impl(tensors, meta, gc, gd, ga)
        dmeans3D, dmeans2D, dsh, dcolors, dopacity, dscales, drot, dcov3D = out
        return (dmeans3D, dmeans2D[0], dsh, dcolors, dopacity, dscales, drot, dcov3D, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean frustum mask (z_view > 0.2), __init__.py:179-188."""
        lib = _lib.load()
        with torch.no_grad():
            rs = self.raster_settings
            pos = positions.detach().to(torch.float32).contiguous()
            P = pos.shape[0]
            present = torch.zeros((P,), dtype=torch.uint8, device=pos.device)
            view = rs.viewmatrix.to(pos.device, torch.float32).contiguous()
            proj = rs.projmatrix.to(pos.device, torch.float32).contiguous()
            with torch.cuda.device(pos.device):
                st = lib.agr_raster_mark_visible(P, _ptr(pos), _ptr(view), _ptr(proj), _ptr(present),
                                                 C.c_void_p(torch.cuda.current_stream(pos.device).cuda_stream))
            if st != _lib.AGR_OK:
                _raise_status(st, "agr_raster_mark_visible")
        return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)


# ------------------------------------------------------------------------------------------------
# View-batched entry: V cameras of one Gaussian set in ONE call
# ------------------------------------------------------------------------------------------------
class _RasterizeGaussiansBatched(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        V = int(rs.viewmatrix.shape[0])
        color, radii, depth, alpha, saved = _forward_impl(means3D, sh, colors_precomp, opacities, scales, rotations,
                                                          cov3Ds_precomp, rs, V)
        ctx.meta = saved.meta
        ctx.none_mask = [t is None for t in saved.tensors]
        ctx.save_for_backward(*[t if t is not None else torch.empty(0) for t in saved.tensors])
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        tensors = tuple(None if m else t for t, m in zip(ctx.saved_tensors, ctx.none_mask))
        dmeans3D, dmeans2D, dsh, dcolors, dopacity, dscales, drot, dcov3D = _backward_impl(
            tensors, ctx.meta, grad_color, grad_depth, grad_alpha)
        return (dmeans3D, dmeans2D, dsh, dcolors, dopacity, dscales, drot, dcov3D, None)


def rasterize_gaussians_batched(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                raster_settings: BatchedRasterizationSettings):
    """means2D: (V,P,3) dummy that receives screen-space gradients (or None).
    colors_precomp: (P,3) shared by all views or (V,P,3). Returns color (V,3,H,W), radii (V,P),
    depth (V,1,H,W), alpha (V,1,H,W)."""
    V = int(raster_settings.viewmatrix.shape[0])
    if V > _lib.AGR_MAX_VIEWS:
        raise RuntimeError("at most %d views per call" % _lib.AGR_MAX_VIEWS)
    if means2D is None:
        means2D = torch.zeros((V,) + tuple(means3D.shape), dtype=means3D.dtype, device=means3D.device)
    e = torch.Tensor([])
    return _RasterizeGaussiansBatched.apply(
        means3D, means2D, sh if sh is not None else e, colors_precomp if colors_precomp is not None else e,
        opacities, scales if scales is not None else e, rotations if rotations is not None else e,
        cov3Ds_precomp if cov3Ds_precomp is not None else e, raster_settings)
