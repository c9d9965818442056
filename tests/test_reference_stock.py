"""The reference's OWN Python (baseline/_ref, installed verbatim by oracle/install_ref.py) driven through its stock code
path — gaussians/gaussian_renderer.py:19-106 `render3`, network/avatar.py:16-239 `AvatarNet` — three ways:
  (A) with the reference's own native extensions (_C, fused, upfirdn2d): the true baseline;
  (B) with THIS repo's drop-in surface behind the very same `import` statements
      (diff_gaussian_rasterization_depth_alpha/, dropin/fused.py, dropin/upfirdn2d.py): "the reference files run unchanged";
  (C) this package's AvatarNet (fp32 compute) loaded from the same state_dict.
(A) vs (B) pins the boundary, (A) vs (C) pins the whole composition at the reference's map size (1024: the view-feature
path, viewdir_net and the three full-size U-Nets all run), outputs and parameter gradients."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(native, case, inp, out, state=None):
    cmd = [sys.executable, "-m", "oracle.ref_stock", "--native", native, "--case", case, "--inp", inp, "--out", out]
    if state:
        cmd += ["--state", state]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "%s\n%s" % (r.stdout[-2000:], r.stderr[-4000:])
    return np.load(out, allow_pickle=False)


def _need_ref():
    from oracle import ref_stock
    if not ref_stock.available():
        pytest.skip("baseline/_ref not installed (python -m oracle.install_ref needs /root/reference)")


def _l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_reference_render3_runs_on_the_dropin(built_lib, tmp_path):
    """render3 (the reference file, unmodified) on the reference _C vs on this repo's drop-in package."""
    _need_ref()
    from animatablegaussians_b200 import synthetic as S
    P, img = 30000, 256
    g = S.make_gaussians(P, seed=7)
    extrs, Ks = S.ring_cameras(4, img=img, focal=275.0)
    rng = np.random.default_rng(0)
    inp = str(tmp_path / "in.npz")
    np.savez(inp, positions=g["xyz"], opacity=g["opacity"], scales=g["scales"] * 2.0, rotations=g["rotations"], colors=g["rgb"],
             bg=np.array([0.2, 0.4, 0.6], np.float32), extr=extrs[1], intr=Ks[1], H=img, W=img,
             g_color=rng.normal(size=(3, img, img)).astype(np.float32), g_depth=rng.normal(size=(1, img, img)).astype(np.float32),
             g_alpha=rng.normal(size=(1, img, img)).astype(np.float32))
    a = _run("reference", "render3", inp, str(tmp_path / "a.npz"))
    b = _run("dropin", "render3", inp, str(tmp_path / "b.npz"))
    assert "baseline/_ref/ext" in str(a["native_files"][0]) and "baseline" not in str(b["native_files"][0])
    assert np.array_equal(a["radii"], b["radii"]) and np.array_equal(a["visibility_filter"], b["visibility_filter"])
    assert float(a["mask"].max()) > 0.5
    for k in ("render", "depth", "mask", "d_positions", "d_opacity", "d_scales", "d_rotations", "d_colors", "d_viewspace"):
        util.assert_close(k, b[k], a[k], util.TOL)


def test_avatar_render_matches_stock_reference(built_lib, tmp_path):
    _need_ref()
    from animatablegaussians_b200 import avatar, synthetic as S, styleunet_ops as ops
    from oracle import ref_stock
    ops.set_compute_dtype(torch.float32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    P, size, img = 20000, 1024, 256
    can, mats = avatar.synthetic_canonical(P, size=size)
    can = {k: v for k, v in can.items() if k != "dist2"}           # both sides derive the scales from 3-NN distances
    net = avatar.AvatarNet({"with_viewdirs": True}, canonical=can, device="cuda").cuda()
    mats_t = torch.from_numpy(mats).cuda()
    with torch.no_grad():
        pose = net.get_pose_map({"cano2live_jnt_mats_woRoot": mats_t})
        avatar.emulate_pretrained_heads(net, pose[:3])
    state = str(tmp_path / "state.pt")
    torch.save(net.state_dict(), state)
    extrs, Ks = S.ring_cameras(8, img=img, focal=275.0)
    rng = np.random.default_rng(1)
    N = net.init_points.shape[0]
    inp = str(tmp_path / "in.npz")
    np.savez(inp, smpl_pos_map=pose.cpu().numpy(), jnt_mats=mats, extr=extrs[3], intr=Ks[3], H=img, W=img,
             g_rgb=rng.normal(size=(img, img, 3)).astype(np.float32), g_mask=rng.normal(size=(img, img, 1)).astype(np.float32),
             g_offset=(rng.normal(size=(N, 3)) * 1e-2).astype(np.float32), **can)
    a = _run("reference", "avatar", inp, str(tmp_path / "a.npz"), state)
    a2 = _run("reference", "avatar", inp, str(tmp_path / "a2.npz"), state)    # the reference against itself: its own noise floor
    b = _run("dropin", "avatar", inp, str(tmp_path / "b.npz"), state)
    assert int(a["missing"][0]) == 0 and int(a["n_state"][0]) == len(net.state_dict())      # state_dict round trip, strict

    z = np.load(inp)
    net.eval()
    items = {"smpl_pos_map": pose, "cano2live_jnt_mats": mats_t, "extr": torch.from_numpy(extrs[3]).cuda(),
             "intr": torch.from_numpy(Ks[3]).cuda(), "img_w": img, "img_h": img}
    out = net.render(items, bg_color=(0., 0., 0.))
    t = lambda k: torch.from_numpy(z[k]).cuda()
    loss = (out["rgb_map"] * t("g_rgb")).sum() + (out["mask_map"] * t("g_mask")).sum() + (out["offset"] * t("g_offset")).sum()
    loss.backward()
    named = dict(net.named_parameters())
    assert float(a["mask_map"].max()) > 0.5, "the avatar must be in view"
    util.assert_close("scaling (3-NN)", net.cano_gaussian_model.get_scaling.detach().cpu().numpy(), a["scaling"], 1e-5)
    for name, other in (("dropin", b), ("product", None)):
        get = (lambda k: other[k]) if other is not None else (lambda k: (out[k] if k in out else None).detach().cpu().numpy())
        # maps straight out of the U-Nets: tight
        util.assert_close(name + ":offset", get("offset"), a["offset"], 2e-4)
        util.assert_close(name + ":pos_map", get("pos_map"), a["pos_map"], 2e-4)
        util.assert_close(name + ":cano_tex_map", get("cano_tex_map"), a["cano_tex_map"], 2e-4)
        # images: the maps above differ by ~1e-5; through exp(scale), the 300-deep blend chains of this dense 256^2 scene and
        # the discrete alpha >= 1/255 / T < 1e-4 decisions that becomes ~1e-4 typical with a thin tail -> relative L2 plus a
        # bounded outlier fraction
        assert _l2(get("rgb_map"), a["rgb_map"]) <= 1e-3, name + ":rgb_map rel L2 %.3e" % _l2(get("rgb_map"), a["rgb_map"])
        assert _l2(get("mask_map"), a["mask_map"]) <= 1e-3, name + ":mask_map rel L2 %.3e" % _l2(get("mask_map"), a["mask_map"])
        util.assert_close_robust(name + ":rgb_map", get("rgb_map"), a["rgb_map"], 2e-3, outlier_frac=5e-3, outlier_tol=1e-1)
        util.assert_close_robust(name + ":mask_map", get("mask_map"), a["mask_map"], 2e-3, outlier_frac=5e-3, outlier_tol=1e-1)
        errs, floors = {}, {}
        for k in ref_stock.GRAD_KEYS:
            gk = other["grad:" + k] if other is not None else named[k].grad.detach().cpu().numpy()
            errs[k], floors[k] = _l2(gk, a["grad:" + k]), _l2(a2["grad:" + k], a["grad:" + k])
        print(name, "parameter-gradient rel L2 vs the stock reference (reference vs itself):",
              {k: "%.2e (%.2e)" % (errs[k], floors[k]) for k in errs})
        # Whole-network gradients cross ~40 leaky-ReLU kinks -> relative L2 (see tests/test_styleunet.py::_check).  The
        # geometry gradients of the rasterizer (means / scales / rotations -> position_net, other_net) are heavy-tailed and
        # the reference does not reproduce ITSELF on them from one run to the next (its U-Net forward differs by ~1e-7
        # between runs, which a handful of ill-conditioned Gaussians amplify; tools/diag_stock.py prints the per-tensor
        # numbers), so each key is held to the larger of 2e-2 and 4x the reference's own run-to-run distance.
        # Keys on which the reference differs from ITSELF by more than 10 % carry no information about anybody's correctness
        # (both runs are samples of a chaotic quantity): they are printed above and only required to be finite.
        unreliable = sorted(k for k in errs if floors[k] > 0.1)
        bad = {k: v for k, v in errs.items() if k not in unreliable and v > max(2e-2, 4.0 * floors[k])}
        assert not bad, "%s: %s" % (name, {k: "%.3e (floor %.3e)" % (v, floors[k]) for k, v in bad.items()})
        assert all(np.isfinite(errs[k]) for k in unreliable)
        # ... and the colour / view-direction path, which IS reproducible, must stay tight
        tight = [k for k in errs if k.startswith(("color_net", "viewdir_net"))]
        assert tight and all(errs[k] <= 2e-2 for k in tight)
    ops.set_compute_dtype(torch.float32)
