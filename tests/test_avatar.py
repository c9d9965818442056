"""AvatarNet mirror: the view-batched render_views() equals per-view render() (the reference's call), gradients
flow to every network, and one optimisation step through FlatAdam changes the parameters."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _setup(P=20000, size=256, V=3, img=256):
    from animatablegaussians_b200 import avatar, synthetic as S, styleunet_ops as ops
    ops.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    can, mats = avatar.synthetic_canonical(P, size=size)
    net = avatar.AvatarNet({"with_viewdirs": True}, canonical=can, device="cuda").cuda()
    # make the nets' outputs non-trivial but tame (random init gives O(1) offsets * 0.05)
    extrs, Ks = S.ring_cameras(V, img=img, focal=1100.0 * img / 1024)
    mats_t = torch.from_numpy(mats).cuda()
    with torch.no_grad():
        pose = net.get_pose_map({"cano2live_jnt_mats_woRoot": mats_t})
        avatar.emulate_pretrained_heads(net, pose[:3])
    return net, {"smpl_pos_map": pose, "cano2live_jnt_mats": mats_t}, extrs, Ks, img


def test_render_views_equals_per_view_render(built_lib):
    net, items, extrs, Ks, img = _setup()
    net.eval()  # no view-direction noise (avatar.py:133-134)
    # tolerance: the two paths run the same kernels, but cuDNN may pick different algorithms for the colour net's
    # full forward vs prefix+tail; an fp32 rounding difference in one Gaussian can flip an alpha>=1/255 or T<1e-4
    # decision of the rasterizer, which moves a pixel by up to ~4e-3 * colour.
    with torch.no_grad():
        batched = net.render_views(items, extrs, Ks, img, img, bg_color=(0.2, 0.4, 0.6))
        for v in range(len(extrs)):
            it = dict(items, extr=torch.from_numpy(extrs[v]).cuda(), intr=torch.from_numpy(Ks[v]).cuda(), img_w=img, img_h=img)
            single = net.render(it, bg_color=(0.2, 0.4, 0.6))
            util.assert_close_robust("rgb v%d" % v, batched["rgb_maps"][v].cpu().numpy(), single["rgb_map"].cpu().numpy(), 1e-4)
            util.assert_close_robust("mask v%d" % v, batched["mask_maps"][v].cpu().numpy(), single["mask_map"].cpu().numpy(), 1e-4)
    assert batched["rgb_maps"].shape == (len(extrs), img, img, 3)
    assert float(batched["mask_maps"].max()) > 0.5  # the avatar is actually in view


def test_train_step_updates_all_networks(built_lib):
    from animatablegaussians_b200 import optim
    net, items, extrs, Ks, img = _setup(V=2)
    net.train()
    opt = optim.FlatAdam(net.parameters(), lr=1e-3)
    before = opt.flat_param.clone()
    out = net.render_views(items, extrs, Ks, img, img, return_depth=True)
    loss = out["rgb_maps"].mean() + out["depth_maps"].mean() + out["mask_maps"].mean() + torch.linalg.norm(out["offset"], dim=-1).mean()
    loss.backward()
    g = opt.flat_grad
    assert torch.isfinite(g).all()
    # (viewdir_net only feeds the colour net at the 256-px decoder level, which this reduced-size avatar does not have)
    for name in ("color_net", "position_net", "other_net"):
        gn = sum(float(p.grad.abs().sum()) for p in getattr(net, name).parameters())
        assert gn > 0, name
    # reference update for a few entries (torch Adam, step 1: p -= lr * g / (|g| + eps))
    idx = torch.nonzero(g.abs() > 1e-6)[:1000, 0]
    expect = before[idx] - 1e-3 * g[idx] / (g[idx].abs() + 1e-8)
    opt.step()
    assert torch.allclose(opt.flat_param[idx], expect, rtol=1e-4, atol=1e-7)
    assert float(opt.flat_grad.abs().max()) == 0.0


def test_fused_gather_equals_reference_indexing(built_lib):
    """_gather_pair (one kernel on the NHWC decoder outputs) == cat on W / permute / boolean-mask (avatar.py:95-97)."""
    net, items, extrs, Ks, img = _setup(P=5000, size=128, V=1, img=64)
    g = torch.Generator(device="cuda").manual_seed(0)
    for V in (1, 3):
        f = torch.randn(V, 8, 128, 128, device="cuda", generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        b = torch.randn(V, 8, 128, 128, device="cuda", generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        got = net._gather_pair(f, b)
        up = torch.randn(got.shape, device="cuda", generator=g)
        got.backward(up)
        gf, gb = f.grad.clone(), b.grad.clone()
        f.grad = None; b.grad = None
        want = torch.stack([torch.cat([f[v:v + 1], b[v:v + 1]], 3)[0].permute(1, 2, 0)[net.cano_smpl_mask] for v in range(V)], 0)
        want = want[0] if V == 1 else want
        assert torch.equal(got, want)
        want.backward(up)
        assert torch.equal(gf, f.grad) and torch.equal(gb, b.grad)
