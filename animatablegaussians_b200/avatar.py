"""AvatarNet for B200 — host-side mirror of the reference's network/avatar.py (`AvatarNet`, avatar.py:16-239),
gaussians/gaussian_model.py (the parts AvatarNet uses: create_from_pcd + activations, :46-62,115-183) and
gaussians/gaussian_renderer.py (`render3`, :19-106).

Same constructor options (`random_style`, `with_viewdirs`, `weight_viewdirs`), same attributes the trainer
touches (SURVEY.md §8b "model boundary"), same `render(items, bg_color, use_pca, use_vae)` signature and return
keys, same state_dict prefixes (color_net./position_net./other_net./viewdir_net.) — main_avatar.py runs unchanged.

B200-first additions (results identical):
  * the constant boolean-mask gather `map[cano_smpl_mask]` (avatar.py:97,110,122) uses indices computed once;
  * transform_cano2live is one fused kernel (lbs.py);
  * render_views(): V cameras of ONE pose in a single pass — position/other nets and the view-independent
    prefix of the colour net run once, only the view-dependent tail of the colour net, the view-direction
    features and the rasterizer run per view, and the rasterizer takes all V views in one batched call.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import contextlib
import ctypes as C
import os

from . import _lib, camera, lbs, parallel, stats
from . import styleunet_ops as ops
from .rasterizer import GaussianRasterizer, rasterize_gaussians_batched
from .styleunet import DualStyleUNet


_lib.register_symbols({
    "agr_gather_maps_forward": (C.c_int, [C.c_int32] + [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
    "agr_gather_maps_backward": (C.c_int, [C.c_int32] + [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
})


class _GatherMaps(torch.autograd.Function):
    """(front, back) NHWC decoder outputs (V,C,S,S) -> (V,N,C) fp32 at the canonical-mask pixels (agr_avatar.h)."""

    @staticmethod
    def forward(ctx, front, back, half, pix):
        lib = _lib.load()
        front, back = ops._nhwc(front), ops._nhwc(back)
        V, Cc, S, _ = front.shape
        N = half.shape[0]
        out = torch.empty((V, N, Cc), dtype=torch.float32, device=front.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(front.device), stats.stage("avatar_gather", launches=1):
            st = lib.agr_gather_maps_forward(ops._code(front), p(front), p(back), p(half), p(pix), p(out), V, S, Cc, N,
                                             C.c_void_p(torch.cuda.current_stream(front.device).cuda_stream))
        if st != _lib.AGR_OK:
            raise RuntimeError("agr_gather_maps_forward failed: %d" % st)
        ctx.save_for_backward(half, pix)
        ctx.meta = (V, Cc, S, N, front.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        half, pix = ctx.saved_tensors
        V, Cc, S, N, dt = ctx.meta
        g = g.float().contiguous()
        df = torch.empty((V, Cc, S, S), dtype=dt, device=g.device, memory_format=torch.channels_last)
        db = torch.empty_like(df)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(g.device), stats.stage("avatar_gather", launches=1):
            st = lib.agr_gather_maps_backward(ops._code(df), p(g), p(half), p(pix), p(df), p(db), V, S, Cc, N,
                                              C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream))
        if st != _lib.AGR_OK:
            raise RuntimeError("agr_gather_maps_backward failed: %d" % st)
        return df, db, None, None


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def knn_mean_dist2(points, K=3, chunk=2048):
    """Mean squared distance to the K nearest neighbours (excluding self) — what the reference gets from
    pytorch3d.ops.knn_points(K=4)[0][0, :, 1:].mean(-1) (gaussian_model.py:170).  Chunked brute force.  The expansion
    |a|^2 + |b|^2 - 2ab only RANKS candidates (in float64 about the centroid: for metre-scale coordinates with mm spacing
    it cancels to a few percent in fp32); the returned distances are exact differences ((p_i - p_j)^2).sum(), as knn_points
    computes them."""
    N = points.shape[0]
    out = torch.empty(N, dtype=points.dtype, device=points.device)
    pc = (points - points.mean(0, keepdim=True)).double()
    sq = (pc * pc).sum(-1)
    kk = min(2 * (K + 1), N)
    for s in range(0, N, chunk):
        p = pc[s:s + chunk]
        d2 = sq[s:s + chunk, None] + sq[None, :] - 2.0 * (p @ pc.T)
        idx = torch.topk(d2, kk, dim=1, largest=False).indices                    # candidates (self included)
        diff = points[s:s + chunk, None, :] - points[idx]                          # exact, in the input dtype
        exact = (diff * diff).sum(-1)
        vals = torch.topk(exact, min(K + 1, kk), dim=1, largest=False).values
        out[s:s + chunk] = vals[:, 1:].mean(-1)
    return out


class GaussianModel:
    """Canonical Gaussian attributes (constants of the avatar; not an nn.Module in the reference either,
    gaussian_model.py:44 — SURVEY.md Appendix B.1)."""

    def __init__(self, sh_degree=0):
        self.max_sh_degree = sh_degree
        self.scaling_activation = torch.exp
        self.scaling_inverse_activation = torch.log
        self.opacity_activation = torch.sigmoid
        self.inverse_opacity_activation = inverse_sigmoid
        self.rotation_activation = F.normalize

    def create_from_pcd(self, points, colors=None, spatial_lr_scale=1.0, dist2=None):
        pts = points.float()
        if dist2 is None:
            dist2 = knn_mean_dist2(pts)
        dist2 = torch.clamp_min(dist2, 0.0000001)
        self._xyz = pts
        self._scaling = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
        self._rotation = torch.zeros((pts.shape[0], 4), device=pts.device)
        self._rotation[:, 0] = 1
        self._opacity = inverse_sigmoid(0.1 * torch.ones((pts.shape[0], 1), dtype=torch.float, device=pts.device))

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: s.scaling_activation(s._scaling))
    get_scaling_raw = property(lambda s: s._scaling)
    get_rotation = property(lambda s: s.rotation_activation(s._rotation))
    get_rotation_raw = property(lambda s: s._rotation)
    get_opacity = property(lambda s: s.opacity_activation(s._opacity))
    get_opacity_raw = property(lambda s: s._opacity)


def render3(gaussian_vals, bg_color, extr, intr, img_w, img_h, scaling_modifier=1.0):
    """gaussians/gaussian_renderer.py:19-106 for the colours-precomputed case used by AvatarNet."""
    means3D = gaussian_vals['positions']
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    rs = camera.make_raster_settings(extr, intr, int(img_w), int(img_h), bg_color, means3D.device, scaling_modifier,
                                     gaussian_vals.get('max_sh_degree', 0))
    assert not ('colors' in gaussian_vals and 'shs' in gaussian_vals), "Cannot use both color and SH!"
    rendered_image, radii, rendered_depth, rendered_alpha = GaussianRasterizer(raster_settings=rs)(
        means3D=means3D, means2D=screenspace_points, shs=gaussian_vals.get('shs'), colors_precomp=gaussian_vals.get('colors'),
        opacities=gaussian_vals['opacity'], scales=gaussian_vals['scales'], rotations=gaussian_vals['rotations'],
        cov3D_precomp=None)
    return {"render": rendered_image, "depth": rendered_depth, "mask": rendered_alpha,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}


class AvatarNet(nn.Module):
    def __init__(self, opt, canonical=None, device=None):
        """canonical: dict(cano_smpl_map (H,2H,3), cano_nml_map (H,2H,3), lbs (N,J)[, dist2 (N,)]) of numpy arrays.
        When None the three files are read from config.opt['train']['data']['data_dir'] exactly like the
        reference (avatar.py:27,31,43)."""
        super().__init__()
        self.opt = opt
        self.random_style = opt.get('random_style', False)
        self.with_viewdirs = opt.get('with_viewdirs', True)
        self.concurrent_nets = os.environ.get('AGR_SERIAL_NETS', '0') != '1'   # render_views: nets on parallel streams
        # render_views under torch.distributed: each view-independent network runs on ONE owner rank (parallel.py)
        # AGR_NET_PARALLEL: 0 = never, 1 = whenever world_size > 1, unset = from 3 ranks on (at 2 ranks one of them would own
        # two of the three networks: measured 34.3 ms against 30.5 ms for the replicated scheme, profiles/SUMMARY_r02.md)
        self.net_parallel_min_world = {'0': 1 << 30, '1': 2}.get(os.environ.get('AGR_NET_PARALLEL', ''), 3)
        self._np_meta, self._np_dummy = {}, None
        if device is None:
            try:
                import config  # the reference's global config module, when running under main_avatar.py
                device = config.device
            except Exception:
                device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        self.device_ = torch.device(device)
        if canonical is None:
            canonical = self._load_canonical_from_config()
        dev = self.device_
        self.max_sh_degree = 0
        self.cano_gaussian_model = GaussianModel(sh_degree=self.max_sh_degree)
        self.cano_smpl_map = torch.from_numpy(np.asarray(canonical['cano_smpl_map'])).to(torch.float32).to(dev)
        self.cano_smpl_mask = torch.linalg.norm(self.cano_smpl_map, dim=-1) > 0.
        self.init_points = self.cano_smpl_map[self.cano_smpl_mask]
        self.lbs = torch.from_numpy(np.asarray(canonical['lbs'])).to(torch.float32).to(dev).contiguous()
        d2 = canonical.get('dist2')
        self.cano_gaussian_model.create_from_pcd(self.init_points, None, spatial_lr_scale=2.5,
                                                 dist2=None if d2 is None else torch.from_numpy(np.asarray(d2)).float().to(dev))
        size = self.cano_smpl_map.shape[0]
        self.map_size = size
        self.color_net = DualStyleUNet(inp_size=size // 2, inp_ch=3, out_ch=3, out_size=size, style_dim=512, n_mlp=2)
        self.position_net = DualStyleUNet(inp_size=size // 2, inp_ch=3, out_ch=3, out_size=size, style_dim=512, n_mlp=2)
        self.other_net = DualStyleUNet(inp_size=size // 2, inp_ch=3, out_ch=8, out_size=size, style_dim=512, n_mlp=2)
        mk = lambda n: torch.ones([1, n.style_dim], dtype=torch.float32, device=dev) / np.sqrt(n.style_dim)
        self.color_style, self.position_style, self.other_style = mk(self.color_net), mk(self.position_net), mk(self.other_net)
        if self.with_viewdirs:
            self.cano_nml_map = torch.from_numpy(np.asarray(canonical['cano_nml_map'])).to(torch.float32).to(dev)
            self.cano_nmls = self.cano_nml_map[self.cano_smpl_mask]
            self.viewdir_net = nn.Sequential(nn.Conv2d(1, 64, 4, 2, 1), nn.LeakyReLU(0.2, inplace=True), nn.Conv2d(64, 128, 4, 2, 1))
        # constant gather indices of the mask (row-major order == boolean-mask order of the reference)
        rc = torch.nonzero(self.cano_smpl_mask)
        self._half = (rc[:, 1] >= size).long()
        self._pix = rc[:, 0] * size + (rc[:, 1] % size)
        self._half32, self._pix32 = self._half.int().contiguous(), self._pix.int().contiguous()
        self._flat = rc[:, 0] * (2 * size) + rc[:, 1]

    @staticmethod
    def _load_canonical_from_config():
        import cv2 as cv
        import config
        d = config.opt['train']['data']['data_dir'] + '/smpl_pos_map/'
        return dict(cano_smpl_map=cv.imread(d + 'cano_smpl_pos_map.exr', cv.IMREAD_UNCHANGED),
                    cano_nml_map=cv.imread(d + 'cano_smpl_nml_map.exr', cv.IMREAD_UNCHANGED),
                    lbs=np.load(d + 'init_pts_lbs.npy'))

    # ------------------------------------------------------------------ map -> per-Gaussian gather
    def _gather(self, maps):
        """(1, 2C, S, S) front|back maps -> (N, C) in cano_smpl_mask order; == cat on W, permute, [mask].
        A batch (V, 2C, S, S) gives (V, N, C)."""
        C = maps.shape[1] // 2
        if maps.shape[0] == 1:
            m = maps[0].reshape(2, C, -1)
            return m[self._half, :, self._pix]
        m = maps.reshape(maps.shape[0], 2, C, -1)
        return m[:, self._half, :, self._pix].permute(1, 0, 2).contiguous()   # advanced dims come first: (N,V,C)

    def _gather_pair(self, front, back):
        """Decoder outputs (V,C,S,S) x2 (compute dtype, NHWC) -> (N,C) for V == 1 else (V,N,C); one fused gather."""
        out = _GatherMaps.apply(front, back, self._half32, self._pix32)
        return out[0] if out.shape[0] == 1 else out

    def _as_map(self, maps):
        front, back = torch.split(maps, [maps.shape[1] // 2] * 2, 1)
        return torch.cat([front, back], 3)[0].permute(1, 2, 0)

    # ------------------------------------------------------------------ reference methods
    def generate_mean_hands(self, pose_map=None):
        """network/avatar.py:52-82: Gaussian attributes of ONE reference pose, later blended over the hand regions
        (test mode, `fix_hand`).  `pose_map`: (>=3, S/2, S/2) posed position map; None reads the map the reference reads,
        <data_dir>/smpl_pos_map/<fix_hand_id>.exr through the trainer's `config` module (main_avatar.py:583-584)."""
        lbs_argmax = self.lbs.argmax(1)
        self.hand_mask = torch.logical_or(torch.logical_or(lbs_argmax == 20, lbs_argmax == 21), lbs_argmax >= 25)
        if pose_map is None:
            import glob
            import cv2 as cv
            import config
            paths = sorted(glob.glob(config.opt['train']['data']['data_dir'] + '/smpl_pos_map/%08d.exr' % config.opt['test']['fix_hand_id']))
            m = cv.imread(paths[0], cv.IMREAD_UNCHANGED)
            half = m.shape[1] // 2
            m = np.concatenate([m[:, :half], m[:, half:]], 2).transpose((2, 0, 1))
            pose_map = torch.from_numpy(m).to(torch.float32).to(self.device_)
        pose_map = pose_map[:3]
        self.hand_positions = self.get_positions(pose_map)
        self.hand_opacity, self.hand_scales, self.hand_rotations = self.get_others(pose_map)
        self.hand_colors, _ = self.get_colors(pose_map)

    def _fix_hand_active(self):
        """render()'s `fix_hand` switch: the trainer's config (network/avatar.py:187) or an explicit opt / attribute."""
        if self.training:
            return False
        if self.opt.get('fix_hand', False) or getattr(self, 'fix_hand', False):
            return True
        cfg = __import__('sys').modules.get('config')
        o = getattr(cfg, 'opt', None) if cfg is not None else None
        return bool(o and o.get('mode') == 'test' and o.get('test', {}).get('fix_hand', False))

    def _blend_hands(self, items, cano_pts, opacity, scales, rotations):
        """network/avatar.py:187-205 (utils/geo_util.py:104-114 normalize_vert_bbox, per-axis): soft box weights around
        the canonical MANO vertices, zeroed below the body centre, blending the mean-hand attributes in."""
        def nbox(verts):
            lo, hi = verts.min(0, keepdim=True)[0], verts.max(0, keepdim=True)[0]
            return 2 * (self.init_points - 0.5 * (hi + lo)) / (hi - lo)
        wl = torch.sigmoid(2.5 * (nbox(items['left_cano_mano_v'])[..., 0:1] + 2.0))
        wr = torch.sigmoid(-2.5 * (nbox(items['right_cano_mano_v'])[..., 0:1] - 2.0))
        below = self.init_points[..., 1] < items['cano_smpl_center'][1]
        wl[below] = 0.
        wr[below] = 0.
        s = torch.maximum(wl + wr, torch.ones_like(wl))
        w = wl / s + wr / s
        return (w * self.hand_positions + (1.0 - w) * cano_pts, w * self.hand_opacity + (1.0 - w) * opacity,
                w * self.hand_scales + (1.0 - w) * scales, w * self.hand_rotations + (1.0 - w) * rotations)

    def transform_cano2live(self, gaussian_vals, items):
        pos, rot = lbs.transform_cano2live(self.lbs, items['cano2live_jnt_mats'], gaussian_vals['positions'],
                                           gaussian_vals['rotations'])
        gaussian_vals['positions'], gaussian_vals['rotations'] = pos, rot
        return gaussian_vals

    def get_positions(self, pose_map, return_map=False):
        position_map, _ = self.position_net([self.position_style], pose_map[None], randomize_noise=False)
        delta_position = 0.05 * self._gather(position_map)
        positions = delta_position + self.cano_gaussian_model.get_xyz
        if return_map:
            return positions, self._as_map(position_map)
        return positions

    def _activate_others(self, others):
        g = self.cano_gaussian_model
        opacity, scales, rotations = torch.split(others, [1, 3, 4], 1)
        opacity = g.opacity_activation(opacity + g.get_opacity_raw)
        scales = g.scaling_activation(scales + g.get_scaling_raw)
        rotations = g.rotation_activation(rotations + g.get_rotation_raw)
        return opacity, scales, rotations

    def get_others(self, pose_map):
        other_map, _ = self.other_net([self.other_style], pose_map[None], randomize_noise=False)
        return self._activate_others(self._gather(other_map))

    def _color_style(self):
        return torch.rand_like(self.color_style) if self.random_style and self.training else self.color_style

    def get_colors(self, pose_map, front_viewdirs=None, back_viewdirs=None):
        color_map, _ = self.color_net([self._color_style()], pose_map[None], randomize_noise=False,
                                      view_feature1=front_viewdirs, view_feature2=back_viewdirs)
        return self._gather(color_map), self._as_map(color_map)

    def _viewdir_maps(self, items, live_pts, live_nmls, cam_pos=None):
        if cam_pos is None:
            cam_pos = -torch.matmul(torch.linalg.inv(items['extr'][:3, :3]), items['extr'][:3, 3])
        viewdirs = F.normalize(cam_pos[None] - live_pts, dim=-1, eps=1e-3)
        if self.training:
            viewdirs += torch.randn(viewdirs.shape, dtype=viewdirs.dtype, device=viewdirs.device) * 0.1
        viewdirs = F.normalize(viewdirs, dim=-1, eps=1e-3)
        viewdirs = (live_nmls * viewdirs).sum(-1)
        viewdirs_map = torch.zeros(self.cano_nml_map.shape[0] * self.cano_nml_map.shape[1], dtype=viewdirs.dtype, device=viewdirs.device)
        viewdirs_map[self._flat] = viewdirs
        viewdirs_map = viewdirs_map.view(1, 1, *self.cano_nml_map.shape[:2])
        viewdirs_map = F.interpolate(viewdirs_map, None, 0.5, 'nearest')
        half = viewdirs_map.shape[-1] // 2
        return torch.split(viewdirs_map, [half, half], -1)

    def _viewdir_features(self, front, back):
        """weight_viewdirs * viewdir_net(view map) for the front and the back maps (avatar.py:46-50,146-147):
        Conv2d(1,64,4,2,1) -> LeakyReLU(0.2) -> Conv2d(64,128,4,2,1), both halves in one batch, on the kernels of
        include/agr_conv.h (bias + activation in the epilogue) instead of two cuDNN convolutions each."""
        c1, c2 = self.viewdir_net[0], self.viewdir_net[2]
        dt = ops.compute_dtype()
        Vn = front.shape[0]
        x = torch.cat([front, back], 0).to(dt).contiguous(memory_format=torch.channels_last)   # (2V,1,h,w): C = 1
        w1, h1 = ops.mod_weight(c1.weight, None, 1.0, False, dt)
        w2, h2 = ops.mod_weight(c2.weight, None, 1.0, False, dt)
        y = ops.conv2d(x, w1, h1, bias=c1.bias, activate=2, stride=2, pad=1)
        y = ops.conv2d(y, w2, h2, bias=c2.bias, activate=0, stride=2, pad=1)
        w = self.opt.get('weight_viewdirs', 1.)
        if w != 1.:
            y = y * w
        return y[:Vn], y[Vn:]

    def get_viewdir_feat(self, items, live=None, cam_pos=None):
        with torch.no_grad():
            if live is None:
                live = lbs.skin_points(self.lbs, items['cano2live_jnt_mats'], self.init_points, self.cano_nmls)
            front_viewdirs, back_viewdirs = self._viewdir_maps(items, *live, cam_pos=cam_pos)
        return self._viewdir_features(front_viewdirs, back_viewdirs)

    def get_viewdir_feat_batched(self, live, cam_pos):
        """get_viewdir_feat (avatar.py:126-147) for V camera centres at once -> two (V,128,S/8,S/8) features."""
        live_pts, live_nmls = live
        with torch.no_grad():
            viewdirs = F.normalize(cam_pos[:, None, :] - live_pts[None], dim=-1, eps=1e-3)           # (V,N,3)
            if self.training:
                viewdirs = viewdirs + torch.randn(viewdirs.shape, dtype=viewdirs.dtype, device=viewdirs.device) * 0.1
            viewdirs = F.normalize(viewdirs, dim=-1, eps=1e-3)
            viewdirs = (live_nmls[None] * viewdirs).sum(-1)                                            # (V,N)
            Vn = cam_pos.shape[0]
            Hm, Wm = self.cano_nml_map.shape[:2]
            vmap = torch.zeros(Vn, Hm * Wm, dtype=viewdirs.dtype, device=viewdirs.device)
            vmap[:, self._flat] = viewdirs
            vmap = F.interpolate(vmap.view(Vn, 1, Hm, Wm), None, 0.5, 'nearest')
            half = vmap.shape[-1] // 2
            front, back = torch.split(vmap, [half, half], -1)
        return self._viewdir_features(front, back)

    def get_pose_map(self, items):
        live_pts = lbs.skin_points(self.lbs, items['cano2live_jnt_mats_woRoot'], self.init_points)
        live_pos_map = torch.zeros_like(self.cano_smpl_map)
        live_pos_map.view(-1, 3)[self._flat] = live_pts
        live_pos_map = F.interpolate(live_pos_map.permute(2, 0, 1)[None], None, [0.5, 0.5], mode='nearest')[0]
        half = live_pos_map.shape[2] // 2
        live_pos_map = torch.cat(torch.split(live_pos_map, [half, half], 2), 0)
        items.update({'smpl_pos_map': live_pos_map})
        return live_pos_map

    def _bg(self, bg_color):
        return torch.from_numpy(np.asarray(bg_color)).to(torch.float32).to(self.device_)

    def render(self, items, bg_color=(0., 0., 0.), use_pca=False, use_vae=False):
        """Note that no batch index in items. (avatar.py:161-239)"""
        bg_color = self._bg(bg_color)
        pose_map = items['smpl_pos_map'][:3]
        assert not (use_pca and use_vae), "Cannot use both PCA and VAE!"
        if use_pca:
            pose_map = items['smpl_pos_map_pca'][:3]
        if use_vae:
            pose_map = items['smpl_pos_map_vae'][:3]
        cano_pts, pos_map = self.get_positions(pose_map, return_map=True)
        opacity, scales, rotations = self.get_others(pose_map)
        if self.with_viewdirs:
            front_viewdirs, back_viewdirs = self.get_viewdir_feat(items)
        else:
            front_viewdirs, back_viewdirs = None, None
        colors, color_map = self.get_colors(pose_map, front_viewdirs, back_viewdirs)
        if self._fix_hand_active():
            cano_pts, opacity, scales, rotations = self._blend_hands(items, cano_pts, opacity, scales, rotations)
        gaussian_vals = {'positions': cano_pts, 'opacity': opacity, 'scales': scales, 'rotations': rotations,
                         'colors': colors, 'max_sh_degree': self.max_sh_degree}
        nonrigid_offset = gaussian_vals['positions'] - self.init_points
        gaussian_vals = self.transform_cano2live(gaussian_vals, items)
        render_ret = render3(gaussian_vals, bg_color, items['extr'], items['intr'], items['img_w'], items['img_h'])
        ret = {'rgb_map': render_ret['render'].permute(1, 2, 0), 'mask_map': render_ret['mask'].permute(1, 2, 0),
               'offset': nonrigid_offset, 'pos_map': pos_map}
        if not self.training:
            ret.update({'cano_tex_map': color_map, 'posed_gaussians': gaussian_vals})
        return ret

    # ------------------------------------------------------------------ view batch (one pose, V cameras)
    def _side_streams(self):
        if getattr(self, "_streams", None) is None:
            self._streams = (torch.cuda.Stream(self.device_), torch.cuda.Stream(self.device_))
        return self._streams

    def prepare_views(self, extrs, intrs, img_w, img_h, bg_color=(0., 0., 0.), capacity=None):
        """Host-side camera setup for a view batch (everything render3 derives from extr/intr,
        gaussian_renderer.py:44-52, plus the camera centres of get_viewdir_feat, avatar.py:131), uploaded with one
        pinned copy.  `capacity` (tile instances) selects the sync-free rasterizer so that render_views() issues no
        host synchronisation and can be captured in a CUDA graph."""
        dev = self.device_
        bs = camera.make_batched_settings(extrs, intrs, int(img_w), int(img_h), self._bg(bg_color), dev)
        if capacity is not None:
            bs = bs._replace(capacity=int(capacity))
        # cam_pos = -inv(R) t  == camera centre == campos of the raster settings
        return {"settings": bs, "cam_pos": bs.campos, "V": len(extrs)}

    @property
    def net_parallel(self):
        """True when render_views() runs the per-pose networks owner-computes style under the current process group."""
        if not parallel.active():
            return False
        import torch.distributed as dist
        return dist.get_world_size() >= self.net_parallel_min_world

    def _replicated_nets(self, items, pose_map, views, V):
        """Every rank (or the only one) runs the three networks -> (position attributes (N,3), other attributes (N,8), colours)."""
        main = torch.cuda.current_stream()
        side = self._side_streams() if (self.concurrent_nets and pose_map.is_cuda) else None
        if side is not None:
            for st in side:
                st.wait_stream(main)
        with torch.cuda.stream(side[0]) if side is not None else contextlib.nullcontext():
            pf, pb = self.position_net.forward_maps([self.position_style], pose_map[None])
            pgather = self._gather_pair(pf, pb)
        with torch.cuda.stream(side[1]) if side is not None else contextlib.nullcontext():
            of, ob = self.other_net.forward_maps([self.other_style], pose_map[None])
            ogather = self._gather_pair(of, ob)
        if self.with_viewdirs:
            with torch.no_grad():
                live = lbs.skin_points(self.lbs, items['cano2live_jnt_mats'], self.init_points, self.cano_nmls)
            prefix = self.color_net.forward_prefix([self._color_style()], pose_map[None])
            fv, bv = self.get_viewdir_feat_batched(live, views["cam_pos"])
            cf, cb = self.color_net.forward_view_tail(prefix, fv, bv, as_pair=True)
            colors = self._gather_pair(cf, cb)
            if V == 1:
                colors = colors[None]
        else:
            cf, cb = self.color_net.forward_maps([self._color_style()], pose_map[None])
            colors = self._gather_pair(cf, cb)
        if side is not None:
            for st, t in zip(side, (pgather, ogather)):
                main.wait_stream(st)
                t.record_stream(main)
        return pgather, ogather, colors

    def _owner_computes(self, items, pose_map, views, V):
        """Same three results with each view-independent network on ONE owner rank (animatablegaussians_b200/parallel.py):
        position net on rank 0, "other" net on rank 1, colour prefix on rank 2 (mod world size); owners broadcast what the
        rendering ranks need, the backward pass reduces the ranks' gradients onto the owner.  The view-dependent colour tail
        and the view-direction net run on every rank for its own views."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        own_p, own_o, own_c = 0, 1 % world, 2 % world
        if self._np_dummy is None:
            self._np_dummy = torch.zeros(1, device=pose_map.device, requires_grad=True)

        def exchange(key, owner, tensors):
            if key not in self._np_meta:          # first (eager) call: the receivers learn shapes and dtype
                self._np_meta[key] = parallel.share_meta(owner, tensors)
            metas, dtype = self._np_meta[key]
            return parallel.owner_broadcast(owner, metas, dtype, self._np_dummy, tensors)

        # networks this rank owns run next to each other: position / other nets on the side streams, the colour net on the
        # main stream (as in the replicated scheme); the exchanges follow in one fixed order on the main stream
        main = torch.cuda.current_stream()
        side = self._side_streams() if self.concurrent_nets else None
        pt, ot = [], []
        if rank == own_p:
            if side is not None:
                side[0].wait_stream(main)
            with torch.cuda.stream(side[0]) if side is not None else contextlib.nullcontext():
                pt = [self._gather_pair(*self.position_net.forward_maps([self.position_style], pose_map[None]))]
        if rank == own_o:
            if side is not None:
                side[1].wait_stream(main)
            with torch.cuda.stream(side[1]) if side is not None else contextlib.nullcontext():
                ot = [self._gather_pair(*self.other_net.forward_maps([self.other_style], pose_map[None]))]
        prefix, ct = None, []
        if self.with_viewdirs:
            with torch.no_grad():
                live = lbs.skin_points(self.lbs, items['cano2live_jnt_mats'], self.init_points, self.cano_nmls)
            if rank == own_c:
                prefix = self.color_net.forward_prefix([self._color_style()], pose_map[None])
                ct = self.color_net.tail_state(prefix)
            else:
                prefix = self.color_net.tail_prefix([self._color_style()])
        elif rank == own_c:
            ct = [self._gather_pair(*self.color_net.forward_maps([self._color_style()], pose_map[None]))]
        if side is not None:
            for st, ts in ((side[0], pt), (side[1], ot)):
                if ts:
                    main.wait_stream(st)
                    ts[0].record_stream(main)
        (pgather,) = exchange("position", own_p, pt)
        (ogather,) = exchange("other", own_o, ot)
        if not self.with_viewdirs:
            (colors,) = exchange("color", own_c, ct)
            return pgather, ogather, colors
        state = exchange("color_prefix", own_c, ct)
        prefix = self.color_net.with_tail_state(prefix, list(state))
        fv, bv = self.get_viewdir_feat_batched(live, views["cam_pos"])
        cf, cb = self.color_net.forward_view_tail(prefix, fv, bv, as_pair=True)
        colors = self._gather_pair(cf, cb)
        if V == 1:
            colors = colors[None]
        return pgather, ogather, colors

    def render_views(self, items, extrs=None, intrs=None, img_w=None, img_h=None, bg_color=(0., 0., 0.), return_depth=False,
                     views=None):
        """items: pose-level entries ('smpl_pos_map', 'cano2live_jnt_mats'); extrs/intrs: V host (numpy) camera
        matrices, or `views` = prepare_views(...).  Returns rgb_maps (V,H,W,3), mask_maps (V,H,W,1)
        [, depth_maps (V,H,W,1)], offset, pos_map — per view identical to render() with that camera."""
        if views is None:
            views = self.prepare_views(extrs, intrs, img_w, img_h, bg_color)
        V = views["V"]
        pose_map = items['smpl_pos_map'][:3]
        # The three U-Nets are independent until the rasterizer.  Position and "other" nets (batch 1: 2 - 128 CTAs per
        # convolution, less than the 148 SMs) run on two side streams next to the colour net, forward AND backward
        # (autograd replays each node on its forward stream); inside a captured step they become parallel graph branches.
        if pose_map.is_cuda and self.net_parallel:
            pgather, ogather, colors = self._owner_computes(items, pose_map, views, V)
        else:
            pgather, ogather, colors = self._replicated_nets(items, pose_map, views, V)
        pos_map = None   # (the (S,2S,3) map view is only produced by render(); the trainer reads it for visualisation)
        cano_pts = 0.05 * pgather + self.cano_gaussian_model.get_xyz
        opacity, scales, rotations = self._activate_others(ogather)
        nonrigid_offset = cano_pts - self.init_points
        pos, rot = lbs.transform_cano2live(self.lbs, items['cano2live_jnt_mats'], cano_pts, rotations)
        color, radii, depth, alpha = rasterize_gaussians_batched(pos, None, None, colors, opacity, scales, rot, None,
                                                                 views["settings"])
        ret = {'rgb_maps': color.permute(0, 2, 3, 1), 'mask_maps': alpha.permute(0, 2, 3, 1), 'offset': nonrigid_offset,
               'pos_map': pos_map, 'radii': radii}
        if return_depth:
            ret['depth_maps'] = depth.permute(0, 2, 3, 1)
        return ret


@torch.no_grad()
def emulate_pretrained_heads(net, pose_map, pos_std=0.1, other_std=0.3, color_std=0.3, opacity_shift=3.2):
    """Synthetic stand-in for the reference's pretraining stage (main_avatar.py:126-164,266-326), which fits the
    three nets to the canonical attributes (zero offsets, canonical scale/rotation/opacity) BEFORE any rendering.
    Random-init StyleGAN heads emit O(1..10) maps -> metre-sized, screen-filling Gaussians that no real training
    step ever rasterises.  The output maps are linear in the ToRGB weights/biases, so scaling those scales the maps
    exactly: position maps to std `pos_std` (x0.05 => ~5 mm offsets), other maps to `other_std`, colour maps to
    `color_std`, and the opacity logit is shifted by `opacity_shift` (logit(0.1)+3.2 ~ 1.0)."""
    for name, style, std in (("position_net", net.position_style, pos_std), ("other_net", net.other_style, other_std),
                             ("color_net", net.color_style, color_std)):
        n = getattr(net, name)
        m, _ = n([style], pose_map[None], randomize_noise=False)
        f = float(std / m.float().std().clamp_min(1e-12))
        for rgbs in (n.to_rgbs1, n.to_rgbs2):
            for t in rgbs:
                t.conv.weight.mul_(f)
                t.bias.mul_(f)
    for rgbs in (net.other_net.to_rgbs1, net.other_net.to_rgbs2):
        rgbs[-1].bias[0, 0] += 2.0 * opacity_shift  # LL sub-band of channel 0: the Haar synthesis halves it


def synthetic_canonical(P, size=1024, J=55, seed=31359):
    """Canonical inputs of AvatarNet for the synthetic capsule avatar: position / normal maps of
    shape (size, 2*size, 3) with exactly P valid pixels (front half | back half) and dense (P,J) LBS weights."""
    from . import synthetic as S
    half = (P + 1) // 2
    # valid pixels: a centred rectangle-ish silhouette per half, row-major, first `half` (front) / `P-half` (back)
    h_rows = int(math.ceil(math.sqrt(half * 1.7 / 0.45)))
    h_cols = int(math.ceil(half / h_rows))
    assert h_rows <= size and h_cols <= size, "P too large for the map"
    r0, c0 = (size - h_rows) // 2, (size - h_cols) // 2
    pos = np.zeros((size, 2 * size, 3), np.float32)
    nml = np.zeros((size, 2 * size, 3), np.float32)
    rng = np.random.default_rng(seed)
    pts, spacing = S.capsule_points(P, rng)
    # place points: front half then back half, each filling its silhouette row-major
    rr, cc = np.meshgrid(np.arange(h_rows), np.arange(h_cols), indexing="ij")
    rr, cc = rr.reshape(-1), cc.reshape(-1)
    order_f = np.arange(half)
    order_b = np.arange(P - half)
    # guard against exact-zero positions (mask = norm > 0)
    pts = pts + (np.abs(pts).sum(1, keepdims=True) == 0) * 1e-6
    pos[r0 + rr[order_f], c0 + cc[order_f]] = pts[:half]
    pos[r0 + rr[order_b], size + c0 + cc[order_b]] = pts[half:]
    n = pts.copy()
    n[:, 1] = 0
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-6)
    nml[r0 + rr[order_f], c0 + cc[order_f]] = n[:half]
    nml[r0 + rr[order_b], size + c0 + cc[order_b]] = n[half:]
    # boolean-mask order = row-major over the (size, 2*size) map
    mask = np.linalg.norm(pos, axis=-1) > 0
    ordered = pos[mask]
    w, mats = S.make_skinning(ordered, J=J, seed=seed)
    dist2 = np.full((ordered.shape[0],), float(spacing) ** 2, np.float32)
    return dict(cano_smpl_map=pos, cano_nml_map=nml, lbs=w, dist2=dist2), mats
