"""TEST / BENCH INFRASTRUCTURE — runs the UNMODIFIED reference Python (baseline/_ref/AnimatableGaussians, installed by
oracle/install_ref.py) through its own code path:
    gaussians/gaussian_renderer.py:19-106   render3
    network/avatar.py:16-239                AvatarNet.__init__ / .render
    network/styleunet/dual_styleunet.py     DualStyleUNet (+ fused_act.py, upfirdn2d.py, conv2d_gradfix.py)
with one of two sets of native modules behind the reference's `import` statements:
    native = "reference"   baseline/_ref/ext: the reference's own _C / fused / upfirdn2d extensions (the true baseline)
    native = "dropin"      this repo's drop-in surface: diff_gaussian_rasterization_depth_alpha/ (repo root) and
                           dropin/{fused,upfirdn2d}.py on libagr_b200.so — proves the reference files run unchanged on it.
Third-party packages missing from this image are stood in by oracle/shims (pytorch3d, plyfile); the EXR / npy data files
AvatarNet.__init__ reads are synthesised into a temp directory (cv2.imread is pointed at .npy twins of the .exr names:
OpenCV's EXR codec is disabled in this build).  Nothing of the reference's code is edited.

A process can hold only ONE native set (modules are imported by name), so comparisons run this file as a subprocess:
    python -m oracle.ref_stock --native reference|dropin --case render3|avatar --inp in.npz [--state state.pt] --out out.npz
Only tests/ and bench.py --impl reference use this module."""
import argparse
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_PY = os.path.join(ROOT, "baseline", "_ref", "AnimatableGaussians")
REF_EXT = os.path.join(ROOT, "baseline", "_ref", "ext")


def available():
    return os.path.isdir(os.path.join(REF_PY, "network")) and os.path.exists(os.path.join(REF_EXT, "fused.so"))


def setup_paths(native):
    assert native in ("reference", "dropin")
    if not available():
        raise RuntimeError("baseline/_ref is not installed: run `python -m oracle.install_ref` where /root/reference exists")
    paths = [REF_PY, os.path.join(HERE, "shims")]
    paths += [REF_EXT] if native == "reference" else [os.path.join(ROOT, "dropin"), ROOT]
    for p in reversed(paths):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if ROOT not in sys.path:
        sys.path.append(ROOT)       # `oracle.*` (the pytorch3d shim imports oracle.lbs_oracle)


def _patch_imread():
    import cv2
    if getattr(cv2, "_agr_patched", False):
        return
    orig = cv2.imread

    def imread(path, flags=None):
        if str(path).endswith(".exr") and os.path.exists(str(path) + ".npy"):
            return np.load(str(path) + ".npy")
        return orig(path, flags) if flags is not None else orig(path)
    cv2.imread, cv2._agr_patched = imread, True


def write_canonical(canonical, data_dir):
    """The three files AvatarNet.__init__ reads (network/avatar.py:27-31,43)."""
    d = os.path.join(data_dir, "smpl_pos_map")
    os.makedirs(d, exist_ok=True)
    np.save(os.path.join(d, "cano_smpl_pos_map.exr.npy"), np.asarray(canonical["cano_smpl_map"], np.float32))
    np.save(os.path.join(d, "cano_smpl_nml_map.exr.npy"), np.asarray(canonical["cano_nml_map"], np.float32))
    np.save(os.path.join(d, "init_pts_lbs.npy"), np.asarray(canonical["lbs"], np.float32))


def build_avatar(canonical, device="cuda:0", opt=None):
    """reference AvatarNet(opt) on synthetic canonical data; call setup_paths() first."""
    import torch
    _patch_imread()
    import config
    tmp = tempfile.mkdtemp(prefix="agr_ref_data_")
    write_canonical(canonical, tmp)
    config.opt = {"train": {"data": {"data_dir": tmp}}, "test": {}, "mode": "train"}
    config.device = torch.device(device)
    from network.avatar import AvatarNet
    return AvatarNet(opt if opt is not None else {"with_viewdirs": True}).to(device)


def _t(z, k, dev):
    import torch
    return torch.from_numpy(np.asarray(z[k])).to(dev)


def case_render3(z, dev):
    """gaussians/gaussian_renderer.py render3 on given Gaussians; upstream gradients for colour, depth, alpha."""
    import torch
    from gaussians.gaussian_renderer import render3
    names = ("positions", "opacity", "scales", "rotations", "colors")
    vals = {k: _t(z, k, dev).float().requires_grad_(True) for k in names}
    vals["max_sh_degree"] = 0
    H, W = int(z["H"]), int(z["W"])
    out = render3(vals, _t(z, "bg", dev).float(), _t(z, "extr", dev).float(), _t(z, "intr", dev).float(), W, H)
    loss = (out["render"] * _t(z, "g_color", dev)).sum() + (out["depth"] * _t(z, "g_depth", dev)).sum() + (out["mask"] * _t(z, "g_alpha", dev)).sum()
    loss.backward()
    res = {"render": out["render"], "depth": out["depth"], "mask": out["mask"], "radii": out["radii"],
           "visibility_filter": out["visibility_filter"], "d_viewspace": out["viewspace_points"].grad}
    res.update({"d_" + k: vals[k].grad for k in names})
    return {k: v.detach().cpu().numpy() for k, v in res.items()}


GRAD_KEYS = ("position_net.convs1.1.conv.weight", "position_net.to_rgbs2.5.conv.weight", "position_net.conv_in.1.weight",
             "other_net.convs2.10.conv.weight", "other_net.to_rgbs1.5.bias", "other_net.style.1.weight",
             "color_net.convs1.11.conv.weight", "color_net.convs1.10.conv.modulation.weight", "color_net.comb_convs.1.0.weight",
             "color_net.cond_convs.0.conv2.1.weight", "color_net.to_rgbs1.5.conv.weight", "color_net.convs2.11.noise.weight",
             "viewdir_net.0.weight", "viewdir_net.2.weight", "viewdir_net.2.bias")


PG_KEYS = ("positions", "opacity", "scales", "rotations", "colors")


def case_avatar(z, dev, state):
    """AvatarNet.render (network/avatar.py:161-239) in eval mode (no view-direction noise), loss on rgb / mask / offset."""
    import torch
    canonical = {k: z[k] for k in ("cano_smpl_map", "cano_nml_map", "lbs")}
    net = build_avatar(canonical, dev)
    missing = net.load_state_dict(torch.load(state, map_location=dev), strict=True)
    net.eval()
    items = {"smpl_pos_map": _t(z, "smpl_pos_map", dev).float(), "cano2live_jnt_mats": _t(z, "jnt_mats", dev).float(),
             "extr": _t(z, "extr", dev).float(), "intr": _t(z, "intr", dev).float(), "img_w": int(z["W"]), "img_h": int(z["H"])}
    out = net.render(items, bg_color=(0., 0., 0.))
    pg = out["posed_gaussians"]          # eval mode returns the posed Gaussians handed to render3 (network/avatar.py:233-237)
    for k in PG_KEYS:
        pg[k].retain_grad()
    loss = (out["rgb_map"] * _t(z, "g_rgb", dev)).sum() + (out["mask_map"] * _t(z, "g_mask", dev)).sum() + (out["offset"] * _t(z, "g_offset", dev)).sum()
    loss.backward()
    named = dict(net.named_parameters())
    res = {"rgb_map": out["rgb_map"], "mask_map": out["mask_map"], "offset": out["offset"], "pos_map": out["pos_map"],
           "cano_tex_map": out["cano_tex_map"], "loss": loss.reshape(1), "scaling": net.cano_gaussian_model.get_scaling,
           "n_state": torch.tensor([len(net.state_dict())]), "missing": torch.tensor([len(missing.missing_keys) + len(missing.unexpected_keys)])}
    for k in GRAD_KEYS:
        res["grad:" + k] = named[k].grad
    for k in PG_KEYS:                    # the rasterizer's inputs and the gradients it returned for them
        res["pg:" + k], res["dpg:" + k] = pg[k], pg[k].grad
    return {k: v.detach().float().cpu().numpy() for k, v in res.items()}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--native", required=True, choices=["reference", "dropin"])
    ap.add_argument("--case", required=True, choices=["render3", "avatar"])
    ap.add_argument("--inp", required=True)
    ap.add_argument("--state")
    ap.add_argument("--out", required=True)
    a = ap.parse_args(argv)
    setup_paths(a.native)
    import torch
    torch.backends.cudnn.allow_tf32 = False           # parity runs compare fp32 against fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    z = np.load(a.inp)
    res = case_render3(z, "cuda:0") if a.case == "render3" else case_avatar(z, "cuda:0", a.state)
    import diff_gaussian_rasterization_depth_alpha as D
    import fused
    import upfirdn2d
    res["native_files"] = np.array([D.__file__, fused.__file__, upfirdn2d.__file__])
    np.savez(a.out, **res)


if __name__ == "__main__":
    main()
