"""Diagnostic for tests/test_reference_stock.py::test_avatar_render_matches_stock_reference: where do the reference's own
extensions (A) and the drop-in (B) start to differ?  Prints, for the rasterizer inputs / returned gradients and the sampled
parameter gradients: relative L2, max-normalised error, and how heavy-tailed the tensor is (max / rms)."""
import os, sys, tempfile
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import test_reference_stock as T
from animatablegaussians_b200 import avatar, synthetic as S, styleunet_ops as ops
from oracle import ref_stock

def main():
    tmp = tempfile.mkdtemp()
    ops.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    P, size, img = 20000, 1024, int(os.environ.get("DIAG_IMG", "256"))
    can, mats = avatar.synthetic_canonical(P, size=size)
    can = {k: v for k, v in can.items() if k != "dist2"}
    net = avatar.AvatarNet({"with_viewdirs": True}, canonical=can, device="cuda").cuda()
    mats_t = torch.from_numpy(mats).cuda()
    with torch.no_grad():
        pose = net.get_pose_map({"cano2live_jnt_mats_woRoot": mats_t})
        avatar.emulate_pretrained_heads(net, pose[:3])
    state = os.path.join(tmp, "state.pt")
    torch.save(net.state_dict(), state)
    extrs, Ks = S.ring_cameras(8, img=img, focal=275.0 * img / 256.0)
    rng = np.random.default_rng(1)
    N = net.init_points.shape[0]
    inp = os.path.join(tmp, "in.npz")
    for mode in ("normal",):
        if mode == "normal":
            g_rgb = rng.normal(size=(img, img, 3)).astype(np.float32); g_mask = rng.normal(size=(img, img, 1)).astype(np.float32)
        else:
            yy, xx = np.meshgrid(np.linspace(0, 1, img), np.linspace(0, 1, img), indexing="ij")
            g_rgb = np.stack([0.5 + xx, 1.0 - 0.5 * yy, 0.3 + xx * yy], -1).astype(np.float32); g_mask = (0.2 + 0.5 * yy)[..., None].astype(np.float32)
        np.savez(inp, smpl_pos_map=pose.cpu().numpy(), jnt_mats=mats, extr=extrs[3], intr=Ks[3], H=img, W=img, g_rgb=g_rgb, g_mask=g_mask,
                 g_offset=(rng.normal(size=(N, 3)) * 1e-2).astype(np.float32), **can)
        a = T._run("reference", "avatar", inp, os.path.join(tmp, "a.npz"), state)
        a2 = T._run("reference", "avatar", inp, os.path.join(tmp, "a2.npz"), state)
        b = T._run("dropin", "avatar", inp, os.path.join(tmp, "b.npz"), state)
        print("==== upstream gradient:", mode, "img", img, "mask coverage %.3f" % float((a["mask_map"] > 0.5).mean()), "mask mean %.3f" % float(a["mask_map"].mean()))
        for k in ("pg:positions", "pg:scales", "pg:opacity", "scaling", "offset"):
            v = a[k].astype(np.float64)
            print("   %-14s shape %s  min %s  max %s  mean %.4g  std %.4g" % (k, v.shape, np.round(v.min(0), 4) if v.ndim == 2 and v.shape[1] <= 4 else round(float(v.min()), 4),
                  np.round(v.max(0), 4) if v.ndim == 2 and v.shape[1] <= 4 else round(float(v.max()), 4), v.mean(), v.std()))
        for k in sorted(a.files):
            if not (k.startswith("dpg:") or k.startswith("grad:")):
                continue
            x, y, x2 = a[k].astype(np.float64), b[k].astype(np.float64), a2[k].astype(np.float64)
            rms = np.sqrt((x ** 2).mean()) + 1e-300
            print("%-52s relL2 %.2e (ref-vs-ref %.2e)  maxnorm %.2e  max/rms %.1e  frac(|err|>1e-2|x|+1e-6rms) %.2e" % (
                k, np.linalg.norm(x - y) / (np.linalg.norm(x) + 1e-300), np.linalg.norm(x - x2) / (np.linalg.norm(x) + 1e-300),
                np.abs(x - y).max() / (np.abs(x).max() + 1e-300), np.abs(x).max() / rms,
                float((np.abs(x - y) > 1e-2 * np.abs(x) + 1e-6 * rms).mean())))
        sys.stdout.flush()

if __name__ == "__main__":
    main()
