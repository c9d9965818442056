"""Photometric loss head after the rasterizer (first part): `photometric_loss` mirrors the
L1 image + L1 mask terms of the trainer (main_avatar.py:193-222) for a batch of V views in ONE CUDA pass that also
produces the gradients (include/agr_loss.h).  LPIPS and the patch crop are not part of it."""
import ctypes as C

import torch

from . import _lib, stats

_p = C.c_void_p
_lib.register_symbols({
    "agr_photometric_loss": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int64, C.c_float, C.c_float, _p, _p, _p, _p]),
    "agr_photometric_loss_u8": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int64, C.c_float, C.c_float, _p, _p, _p, _p]),
})


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, alpha, gt_rgb, mask, boundary, bg, w_l1, w_mask):
        lib = _lib.load()
        rgb_c = rgb.detach().float().contiguous()
        alpha_c = alpha.detach().float().contiguous() if alpha is not None else None
        u8 = gt_rgb.dtype == torch.uint8           # camera bytes: value / 255 on the device
        gt_c = gt_rgb.detach().contiguous() if u8 else gt_rgb.detach().float().contiguous()
        m8, b8 = mask.to(torch.uint8).contiguous(), boundary.to(torch.uint8).contiguous()
        bg_c = bg.detach().float().contiguous()
        pixels = rgb_c.numel() // 3
        if gt_c.numel() != 3 * pixels or m8.numel() != pixels or b8.numel() != pixels or (alpha_c is not None and alpha_c.numel() != pixels):
            raise ValueError("photometric_loss: inconsistent shapes")
        sums = torch.zeros(2, dtype=torch.float32, device=rgb_c.device)
        d_rgb = torch.empty_like(rgb_c)
        d_alpha = torch.empty_like(alpha_c) if alpha_c is not None else None
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(rgb_c.device), stats.stage("loss_head", launches=1):
            fn = lib.agr_photometric_loss_u8 if u8 else lib.agr_photometric_loss
            st = fn(ptr(rgb_c), ptr(alpha_c), ptr(gt_c), ptr(m8), ptr(b8), ptr(bg_c), pixels, float(w_l1),
                                          float(w_mask), ptr(sums), ptr(d_rgb), ptr(d_alpha),
                                          C.c_void_p(torch.cuda.current_stream(rgb_c.device).cuda_stream))
        if st != _lib.AGR_OK:
            raise RuntimeError("agr_photometric_loss failed: %d" % st)
        l1 = sums[0] / (3.0 * pixels)
        mk = sums[1] / float(pixels)
        ctx.save_for_backward(d_rgb, d_alpha)
        ctx.shapes = (rgb.shape, None if alpha is None else alpha.shape)
        ctx.mark_non_differentiable(l1, mk)
        return w_l1 * l1 + w_mask * mk, l1, mk

    @staticmethod
    def backward(ctx, g_total, _g_l1, _g_mk):
        d_rgb, d_alpha = ctx.saved_tensors
        rs, as_ = ctx.shapes
        gr = (d_rgb * g_total).view(rs) if ctx.needs_input_grad[0] else None
        ga = (d_alpha * g_total).view(as_) if (d_alpha is not None and ctx.needs_input_grad[1]) else None
        return gr, ga, None, None, None, None, None, None


def photometric_loss(rgb_maps, mask_maps, color_imgs, mask_imgs, boundary_mask_imgs, bg_color, w_l1=1.0, w_mask=0.1):
    """rgb_maps (V,H,W,3) / mask_maps (V,H,W,1) from `AvatarNet.render_views`, ground truth colour (V,H,W,3), boolean
    masks (V,H,W), bg_color (3,) tensor; `color_imgs` may be uint8 (the camera bytes; /255 happens in the kernel).  Returns (w_l1 * l1 + w_mask * mask_loss, l1, mask_loss); the last two are
    the per-term values the trainer logs (not differentiable)."""
    if not rgb_maps.is_cuda:
        raise RuntimeError("photometric_loss runs on the GPU only (no CPU fallback)")
    return _PhotometricLoss.apply(rgb_maps, mask_maps, color_imgs, mask_imgs, boundary_mask_imgs, bg_color, w_l1, w_mask)


def crop_image(gt_mask, patch_size, randomly, *images, bg_color=None, generator=None):
    """Patch crop in front of the perceptual loss, as the trainer does it (main_avatar.py:75-115): the bounding box of
    `gt_mask > 0` ((H, W)) is cut from every (3, H, W) image, centred on a square `bg_color` canvas of the box's longer
    side, and either resized to `patch_size` (bilinear) or — `randomly` and the square larger than the patch — a random
    patch_size window of it is taken (the same window for every image).  Returns a list (one tensor if one image).

    Like the reference, the box comes back to the host once per call (its corner indices size the canvas).  The box is read
    from row / column occupancy instead of `argwhere` (same four numbers, no (n, 2) index list)."""
    import torch.nn.functional as F
    if not images:
        raise ValueError("crop_image: no images")
    occ = gt_mask > 0.
    rows, cols = occ.any(1), occ.any(0)
    if not bool(rows.any()):
        raise ValueError("crop_image: empty mask")     # the reference fails here too (min of an empty tensor)
    r, c = torch.nonzero(rows).flatten(), torch.nonzero(cols).flatten()
    min_v, max_v, min_u, max_u = int(r[0]), int(r[-1]), int(c[0]), int(c[-1])
    len_v, len_u = max_v - min_v, max_u - min_u       # exclusive of the last row / column, as the reference slices
    max_size = max(len_v, len_u)
    window = randomly and max_size > patch_size
    if window:
        rv = int(torch.randint(0, max_size - patch_size + 1, (1,), generator=generator))
        ru = int(torch.randint(0, max_size - patch_size + 1, (1,), generator=generator))
    out = []
    for image in images:
        bg = torch.zeros(3, dtype=image.dtype, device=image.device) if bg_color is None else bg_color.to(image)
        canvas = bg[:, None, None] * torch.ones((3, max_size, max_size), dtype=image.dtype, device=image.device)
        if len_v > len_u:
            s = (max_size - len_u) // 2
            canvas[:, :, s:s + len_u] = image[:, min_v:max_v, min_u:max_u]
        else:
            s = (max_size - len_v) // 2
            canvas[:, s:s + len_v, :] = image[:, min_v:max_v, min_u:max_u]
        if window:
            canvas = canvas[:, rv:rv + patch_size, ru:ru + patch_size]
        else:
            canvas = F.interpolate(canvas[None], size=(patch_size, patch_size), mode="bilinear")[0]
        out.append(canvas)
    return out if len(out) > 1 else out[0]
