"""Generates tests/golden/camera.npz from the UNMODIFIED reference camera math (utils/graphics_utils.py:51-85
getProjectionMatrix / focal2fov and the matrix composition of gaussians/gaussian_renderer.py:44-52), CPU."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("AGR_REFERENCE_ROOT", "/root/reference")


def cases():
    rng = np.random.default_rng(0)
    out = []
    for i in range(6):
        W, H = int(rng.integers(200, 1600)), int(rng.integers(200, 1600))
        K = np.array([[rng.uniform(300, 2000), 0, W / 2 + rng.uniform(-40, 40)], [0, rng.uniform(300, 2000), H / 2 + rng.uniform(-40, 40)], [0, 0, 1]], np.float32)
        a = rng.normal(0, 1, 3); a /= np.linalg.norm(a); th = rng.uniform(0, 3)
        Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * Kx @ Kx
        E = np.eye(4, dtype=np.float32); E[:3, :3] = R; E[:3, 3] = rng.normal(0, 2, 3)
        out.append((E, K, W, H))
    return out


if __name__ == "__main__":
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_graphics_utils", os.path.join(REF, "utils", "graphics_utils.py"))
    gu = importlib.util.module_from_spec(spec); spec.loader.exec_module(gu)
    d = {}
    for i, (E, K, W, H) in enumerate(cases()):
        extr, intr = torch.from_numpy(E), torch.from_numpy(K)
        FoVx = gu.focal2fov(intr[0, 0].item(), W)
        FoVy = gu.focal2fov(intr[1, 1].item(), H)
        wv = extr.transpose(1, 0)
        proj = gu.getProjectionMatrix(znear=0.1, zfar=100, fovX=FoVx, fovY=FoVy, K=intr, img_w=W, img_h=H).transpose(0, 1)
        full = (wv.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
        d["view%d" % i] = wv.numpy(); d["proj%d" % i] = full.numpy(); d["campos%d" % i] = torch.linalg.inv(extr)[:3, 3].numpy()
        d["tan%d" % i] = np.array([math.tan(FoVx * 0.5), math.tan(FoVy * 0.5)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "camera.npz"), **d)
    print("wrote camera.npz")
