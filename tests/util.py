"""Shared helpers for the parity tests: seeded scenes and the parity metric of SURVEY.md §8(c):
   max|a-b| / max(max|b|, eps) <= 1e-4 per output tensor."""
import math

import numpy as np

from animatablegaussians_b200 import camera as cam

TOL = 1e-4  # north_star: "match the reference's own CUDA rasterizer on identical inputs to <= 1e-4 relative"


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_close(name, a, b, tol=TOL):
    e = rel_err(a, b)
    assert e <= tol, "%s: rel max-norm error %.3e > %.1e (max|ref| = %.3e)" % (name, e, tol, float(np.abs(b).max()))
    return e


def look_at_extr(eye, target=(0, 0, 0), up=(0, -1, 0)):
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(-up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ eye
    return E.astype(np.float32)


def random_scene(P, W, H, seed, focal=None, behind_frac=0.1, big_frac=0.05, sh_degree=None, cov_precomp=False,
                 spread=1.0, opacity_mean=0.0):
    """A small random scene in front of one camera (some points behind it, some huge splats)."""
    rng = np.random.default_rng(seed)
    focal = focal or 0.9 * max(W, H)
    K = np.array([[focal, 0, W / 2 + 0.37], [0, focal * 1.03, H / 2 - 1.21], [0, 0, 1]], np.float32)
    extr = look_at_extr(eye=(0.3, -0.2, -3.0))
    xyz = rng.normal(0, 1, (P, 3)).astype(np.float32) * np.array([0.9, 0.7, 0.6], np.float32) * spread
    nb = int(P * behind_frac)
    if nb:
        xyz[:nb, 2] -= 4.0  # behind the camera / inside the near plane
    scales = np.exp(rng.normal(math.log(0.03), 0.5, (P, 3))).astype(np.float32)
    nbig = int(P * big_frac)
    if nbig:
        scales[-nbig:] *= 8.0
    q = rng.normal(0, 1, (P, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q *= rng.uniform(0.7, 1.3, (P, 1)).astype(np.float32)  # the reference does NOT normalise (forward.cu:127)
    opacity = (1 / (1 + np.exp(-rng.normal(opacity_mean, 2, (P, 1))))).astype(np.float32)
    rgb = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    cb = cam.camera_block(extr, K, W, H)
    sc = dict(P=P, W=W, H=H, xyz=xyz, scales=scales, rotations=q, opacity=opacity, rgb=rgb, extr=extr, K=K,
              bg=rng.uniform(0, 1, 3).astype(np.float32), sh=None, sh_degree=0, cov3D=None, **cb)
    if sh_degree is not None:
        M = (sh_degree + 1) ** 2
        sc["sh"] = rng.normal(0, 0.5, (P, M, 3)).astype(np.float32)
        sc["sh_degree"] = sh_degree
        sc["rgb"] = None
    if cov_precomp:
        A = rng.normal(0, 0.04, (P, 3, 3)).astype(np.float32)
        S = A @ A.transpose(0, 2, 1) + 1e-4 * np.eye(3, dtype=np.float32)
        sc["cov3D"] = np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1).astype(np.float32)
        sc["scales"] = None
        sc["rotations"] = None
    return sc


def upstream_grads(W, H, seed):
    rng = np.random.default_rng(seed + 7)
    return (rng.normal(0, 1, (3, H, W)).astype(np.float32), rng.normal(0, 1, (1, H, W)).astype(np.float32),
            rng.normal(0, 1, (1, H, W)).astype(np.float32))


def run_oracle(sc, grads=None, alpha_from=None):
    from oracle.raster_oracle import RasterOracle
    o = RasterOracle()
    color, radii, depth, alpha = o.forward(sc["bg"], sc["xyz"], sc["rgb"], sc["opacity"], sc["scales"], sc["rotations"],
                                           1.0, sc["cov3D"], sc["viewmatrix"], sc["projmatrix"], sc["tanfovx"],
                                           sc["tanfovy"], sc["H"], sc["W"], sh=sc["sh"], degree=sc["sh_degree"],
                                           campos=sc["campos"])
    out = dict(color=color, radii=radii, depth=depth, alpha=alpha, R=o.num_rendered)
    if grads is not None:
        out["grads"] = o.backward(*grads, alpha_override=alpha_from)
    return out


def assert_close_robust(name, a, b, tol=TOL, outlier_frac=1e-4, outlier_tol=2e-2):
    """For comparisons that cross the rasterizer's discrete decisions (alpha >= 1/255, T < 1e-4): all but a
    vanishing fraction of the elements must agree to `tol`, and no element may be off by more than `outlier_tol`
    (relative to max|b|)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b) / scale
    frac = float((err > tol).mean())
    assert frac <= outlier_frac, "%s: %.2e of the elements differ by more than %.1e" % (name, frac, tol)
    assert err.max() <= outlier_tol, "%s: worst element off by %.2e" % (name, err.max())
