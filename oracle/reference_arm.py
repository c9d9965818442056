"""TEST / BENCH INFRASTRUCTURE — the reference arm of bench.py (`--impl reference`).

What runs: the reference's OWN algorithm and native kernels, at the reference's own granularity
(one pose x ONE view per iteration, batch_size 1, fp32, torch.optim.Adam — main_avatar.py:166-264):
  * rasterizer            : UNMODIFIED reference CUDA kernels (oracle/_ref/libref_rasterizer.so, forward.cu /
                            backward.cu / rasterizer_impl.cu compiled from /root/reference by oracle/build_ref.py)
  * fused lrelu, upfirdn2d: UNMODIFIED reference CUDA extensions (oracle/_ref/ref_fused.so, ref_upfirdn2d.so)
  * convolutions          : cuDNN through torch, fp32, TF32 at torch defaults — what conv2d_gradfix.py falls through to
  * module tree / LBS     : oracle restatements (oracle/styleunet_oracle.py, oracle/lbs_oracle.py) — the reference's
                            Python cannot travel to the GPU box, and pytorch3d is not installed anywhere.
A "step" = the same 16 views of one pose as the product arm = 16 reference iterations (each: 3 U-Nets fwd, LBS, raster,
loss, backward, Adam).  The reference has no CPU rasterizer (BASELINE.json), so this GPU run IS the reference baseline.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _RefRasterFn(torch.autograd.Function):
    """Autograd wiring of the reference kernels: restates _RasterizeGaussians (RAST/.../__init__.py:44-158)."""

    @staticmethod
    def forward(ctx, ref, means3D, colors, opacity, scales, rotations, cam, bg, H, W):
        color, radii, depth, alpha = ref.forward(bg, means3D, colors, opacity, scales, rotations, 1.0, None, cam["view"],
                                                 cam["proj"], cam["tanfovx"], cam["tanfovy"], H, W, campos=cam["campos"])
        ctx.ref = ref
        return color, depth, alpha

    @staticmethod
    def backward(ctx, gc, gd, ga):
        g = ctx.ref.backward(gc, gd, ga)
        return None, g["means3D"], g["colors"], g["opacity"], g["scales"], g["rotations"], None, None, None, None


class RefAvatar(nn.Module):
    """network/avatar.py:16-239 restated on top of the oracle modules (single view per call, like the reference)."""

    def __init__(self, canonical, device):
        super().__init__()
        from oracle import styleunet_oracle as so
        self.so = so
        dev = device
        self.cano_smpl_map = torch.from_numpy(canonical["cano_smpl_map"]).float().to(dev)
        self.cano_smpl_mask = torch.linalg.norm(self.cano_smpl_map, dim=-1) > 0.
        self.init_points = self.cano_smpl_map[self.cano_smpl_mask]
        self.lbs = torch.from_numpy(canonical["lbs"]).float().to(dev)
        N = self.init_points.shape[0]
        self.raw_scale = torch.log(torch.sqrt(torch.from_numpy(canonical["dist2"]).float().to(dev)))[:, None].repeat(1, 3)
        self.raw_rot = torch.zeros(N, 4, device=dev); self.raw_rot[:, 0] = 1
        self.raw_opacity = torch.full((N, 1), float(np.log(0.1 / 0.9)), device=dev)
        size = self.cano_smpl_map.shape[0]
        mk = lambda oc: so.DualStyleUNet(inp_size=size // 2, inp_ch=3, out_ch=oc, out_size=size, style_dim=512, n_mlp=2)
        self.color_net, self.position_net, self.other_net = mk(3), mk(3), mk(8)
        self.style = torch.ones(1, 512, device=dev) / np.sqrt(512)
        self.cano_nml_map = torch.from_numpy(canonical["cano_nml_map"]).float().to(dev)
        self.cano_nmls = self.cano_nml_map[self.cano_smpl_mask]
        self.viewdir_net = nn.Sequential(nn.Conv2d(1, 64, 4, 2, 1), nn.LeakyReLU(0.2, inplace=True), nn.Conv2d(64, 128, 4, 2, 1))

    def _gather(self, m, c):
        front, back = torch.split(m, [c, c], 1)
        return torch.cat([front, back], 3)[0].permute(1, 2, 0)[self.cano_smpl_mask]   # boolean mask, like the reference

    def render(self, pose_map, jnt_mats, extr, cam, bg, H, W, ref_raster):
        from oracle import lbs_oracle
        pos_map, _ = self.position_net([self.style], pose_map[None], randomize_noise=False)
        cano_pts = 0.05 * self._gather(pos_map, 3) + self.init_points
        other_map, _ = self.other_net([self.style], pose_map[None], randomize_noise=False)
        others = self._gather(other_map, 8)
        opacity = torch.sigmoid(others[:, :1] + self.raw_opacity)
        scales = torch.exp(others[:, 1:4] + self.raw_scale)
        rotations = F.normalize(others[:, 4:] + self.raw_rot)
        with torch.no_grad():  # get_viewdir_feat, avatar.py:126-147
            live_pts, live_nmls = lbs_oracle.skin_points(self.lbs, jnt_mats, self.init_points, self.cano_nmls)
            cam_pos = -torch.matmul(torch.linalg.inv(extr[:3, :3]), extr[:3, 3])
            vd = F.normalize(cam_pos[None] - live_pts, dim=-1, eps=1e-3)
            vd = F.normalize(vd + torch.randn_like(vd) * 0.1, dim=-1, eps=1e-3)
            vd = (live_nmls * vd).sum(-1)
            vmap = torch.zeros(*self.cano_nml_map.shape[:2], device=vd.device)
            vmap[self.cano_smpl_mask] = vd
            vmap = F.interpolate(vmap[None, None], None, 0.5, "nearest")
            half = vmap.shape[-1] // 2
            fv, bv = torch.split(vmap, [half, half], -1)
        fv, bv = self.viewdir_net(fv), self.viewdir_net(bv)
        color_map, _ = self.color_net([self.style], pose_map[None], randomize_noise=False, view_feature1=fv, view_feature2=bv)
        colors = self._gather(color_map, 3)
        offset = cano_pts - self.init_points
        pos, rot = lbs_oracle.transform_cano2live(self.lbs, jnt_mats, cano_pts, rotations)
        color, depth, alpha = _RefRasterFn.apply(ref_raster, pos, colors, opacity, scales, rot, cam, bg, H, W)
        return color, depth, alpha, offset


def main_stock(args, rank, world, local):
    """The reference's OWN Python through its stock code path (baseline/_ref, oracle/ref_stock.py): network/avatar.py
    AvatarNet.render -> gaussians/gaussian_renderer.py render3 -> the reference's _C / fused / upfirdn2d extensions,
    one pose x ONE view per iteration with torch.optim.Adam, as main_avatar.py:166-264 trains."""
    import bench
    from animatablegaussians_b200 import avatar as prod_avatar, synthetic as S  # synthetic data only (numpy)
    from oracle import ref_stock
    ref_stock.setup_paths("reference")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(31359)
    torch.backends.cudnn.benchmark = True
    P, IMG, NV, J = bench.P_GAUSS, bench.IMG, bench.N_VIEWS, bench.J
    canonical, mats = prod_avatar.synthetic_canonical(P, size=IMG, J=J)
    net = ref_stock.build_avatar(canonical, dev)
    net.train()
    extrs, Ks = S.ring_cameras(NV, img=IMG)
    jnt = torch.from_numpy(mats).to(dev)
    with torch.no_grad():
        pose = net.get_pose_map({"cano2live_jnt_mats_woRoot": jnt})
        # same emulated pre-trained state as the product arm (linear rescale of the ToRGB heads)
        for name, style, std in (("position_net", net.position_style, 0.1), ("other_net", net.other_style, 0.3), ("color_net", net.color_style, 0.3)):
            n = getattr(net, name)
            m, _ = n([style], pose[:3][None], randomize_noise=False)
            f = float(std / m.std().clamp_min(1e-12))
            for rgbs in (n.to_rgbs1, n.to_rgbs2):
                for t in rgbs:
                    t.conv.weight.mul_(f); t.bias.mul_(f)
        for rgbs in (net.other_net.to_rgbs1, net.other_net.to_rgbs2):
            rgbs[-1].bias[0, 0] += 6.4
    opt = torch.optim.Adam(net.parameters(), lr=1e-7)
    views = [dict(extr=torch.from_numpy(e).to(dev), intr=torch.from_numpy(k).to(dev)) for e, k in zip(extrs, Ks)]

    def step():
        for v in range(NV):  # one reference iteration per view (batch_size 1, configs/*/avatar.yaml)
            items = {"smpl_pos_map": pose, "cano2live_jnt_mats": jnt, "extr": views[v]["extr"], "intr": views[v]["intr"], "img_w": IMG, "img_h": IMG}
            out = net.render(items, bg_color=(0., 0., 0.))
            # AvatarNet.render drops the depth map (avatar.py:223-231): colour and mask only + the offset regulariser
            loss = (out["rgb_map"].sum() + out["mask_map"].sum()) * (1.0 / (IMG * IMG)) + 0.005 * torch.linalg.norm(out["offset"], dim=-1).mean()
            loss.backward()
            opt.step()
            opt.zero_grad()
        return loss

    for _ in range(max(args.warmup, 1)):
        step()
    sampler = bench.ClockSampler(local)
    sampler.start()
    ms = bench.device_time_ms(step, args.steps, 1)
    clocks = sampler.stop()
    ms_step = ms / args.steps
    value = NV / (ms_step * 1e-3)
    import diff_gaussian_rasterization_depth_alpha as D
    import network.avatar as NA
    out = {"impl": "reference", "metric": bench.METRIC, "value": value, "unit": "views/s", "n_gpus": 1, "steps": args.steps,
           "warmup": max(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "fp32 (TF32 convs at torch defaults), as the reference", "data": "synthetic",
           "config": {"workload": "same 16 views of one pose as the product arm, run by the reference's own code: 16 iterations of "
                                  "AvatarNet.render (3 DualStyleUNets, LBS, render3) + loss + backward + torch.optim.Adam, batch 1, fp32",
                      "gaussians": int(net.init_points.shape[0]), "views_per_step": NV, "image": [IMG, IMG],
                      "code_path": "stock: %s + %s (verbatim copies in baseline/_ref); _C / fused / upfirdn2d built unmodified for sm_100a; "
                                   "pytorch3d / plyfile stood in by oracle/shims" % (os.path.relpath(NA.__file__, ROOT), os.path.relpath(D.__file__, ROOT)),
                      "loss": "colour + mask sums (AvatarNet.render returns no depth) + offset regulariser; the product arm also "
                              "back-propagates a depth term"},
           "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "cpu_baseline": {"value": value, "unit": "views/s", "cores": 0, "kind": "reference",
                            "sample": "the reference has no CPU implementation of this path (BASELINE.json); this line is its "
                                      "own CUDA build on 1 B200, %d steps x 16 views" % args.steps},
           "clocks": clocks}
    print(json.dumps(out))
    return 0


def main(args, rank, world, local):
    if rank != 0:
        return 0
    from oracle import ref_stock
    if ref_stock.available() and os.environ.get("AGR_REF_ARM", "stock") == "stock":
        return main_stock(args, rank, world, local)
    import bench
    from animatablegaussians_b200 import avatar as prod_avatar, camera, synthetic as S  # synthetic data + camera math only
    from oracle import ref_rasterizer, styleunet_oracle as so
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(31359)
    torch.backends.cudnn.benchmark = True
    so.set_compute_dtype(torch.float32)
    so.use_reference_cuda_ops()
    P, IMG, NV, J = bench.P_GAUSS, bench.IMG, bench.N_VIEWS, bench.J
    canonical, mats = prod_avatar.synthetic_canonical(P, size=IMG, J=J)
    net = RefAvatar(canonical, dev).to(dev)
    net.train()
    extrs, Ks = S.ring_cameras(NV, img=IMG)
    cams = []
    for e, k in zip(extrs, Ks):
        cb = camera.camera_block(e, k, IMG, IMG)
        cams.append(dict(view=torch.from_numpy(cb["viewmatrix"]).to(dev), proj=torch.from_numpy(cb["projmatrix"]).to(dev),
                         campos=torch.from_numpy(cb["campos"]).to(dev), tanfovx=cb["tanfovx"], tanfovy=cb["tanfovy"],
                         extr=torch.from_numpy(e).to(dev)))
    jnt = torch.from_numpy(mats).to(dev)
    from oracle import lbs_oracle
    with torch.no_grad():  # get_pose_map, avatar.py:149-159
        live = lbs_oracle.skin_points(net.lbs, jnt, net.init_points)
        pm = torch.zeros_like(net.cano_smpl_map)
        pm[net.cano_smpl_mask] = live
        pm = F.interpolate(pm.permute(2, 0, 1)[None], None, [0.5, 0.5], mode="nearest")[0]
        half = pm.shape[2] // 2
        pose_map = torch.cat(torch.split(pm, [half, half], 2), 0)[:3].contiguous()
        # same emulated pre-trained state as the product arm (linear rescale of the ToRGB heads)
        for name, std in (("position_net", 0.1), ("other_net", 0.3), ("color_net", 0.3)):
            n = getattr(net, name)
            m, _ = n([net.style], pose_map[None], randomize_noise=False)
            f = float(std / m.std().clamp_min(1e-12))
            for rgbs in (n.to_rgbs1, n.to_rgbs2):
                for t in rgbs:
                    t.conv.weight.mul_(f); t.bias.mul_(f)
        for rgbs in (net.other_net.to_rgbs1, net.other_net.to_rgbs2):
            rgbs[-1].bias[0, 0] += 6.4
    opt = torch.optim.Adam(net.parameters(), lr=1e-7)
    ref = ref_rasterizer.RefRasterizer()
    bg = torch.zeros(3, device=dev)

    def step():
        for v in range(NV):  # one reference iteration per view
            color, depth, alpha, offset = net.render(pose_map, jnt, cams[v]["extr"], cams[v], bg, IMG, IMG, ref)
            loss = (color.sum() + depth.sum() + alpha.sum()) * (1.0 / (IMG * IMG)) + 0.005 * torch.linalg.norm(offset, dim=-1).mean()
            loss.backward()
            opt.step()
            opt.zero_grad()
        return loss

    for _ in range(max(args.warmup, 1)):
        step()
    sampler = bench.ClockSampler(local)
    sampler.start()
    ms = bench.device_time_ms(step, args.steps, 1)
    clocks = sampler.stop()
    ms_step = ms / args.steps
    value = NV / (ms_step * 1e-3)
    out = {"impl": "reference", "metric": bench.METRIC, "value": value, "unit": "views/s", "n_gpus": 1, "steps": args.steps,
           "warmup": max(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "fp32 (TF32 convs at torch defaults), as the reference", "data": "synthetic",
           "config": {"workload": "same 16 views of one pose as the product arm, run the reference's way: 16 iterations of "
                                  "(3 DualStyleUNets fwd, LBS, raster, loss, backward, torch Adam), batch 1, fp32",
                      "gaussians": int(net.init_points.shape[0]), "views_per_step": NV, "image": [IMG, IMG],
                      "kernels": "reference rasterizer + fused + upfirdn2d CUDA sources built unmodified (oracle/_ref); cuDNN convs"},
           "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "cpu_baseline": {"value": value, "unit": "views/s", "cores": 0, "kind": "reference",
                            "sample": "the reference has no CPU implementation of this path (BASELINE.json); this line is its "
                                      "own CUDA build on 1 B200, %d steps x 16 views" % args.steps},
           "clocks": clocks}
    print(json.dumps(out))
    return 0
