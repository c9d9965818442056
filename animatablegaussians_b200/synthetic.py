"""Synthetic avatar workload of SURVEY.md §8(d) / BASELINE.md §3 (there are no datasets or SMPL-X
model files in the container): a 1.7 m capsule "body" of P Gaussians (front half z>0, back half
z<0), sigma ~ lattice spacing, 55 synthetic joints with top-4 LBS weights, a ring of cameras at
2.5 m with K = [[1100,0,512],[0,1100,512],[0,0,1]] (main_avatar.py:612).  Pure numpy on the host;
used by bench.py, __graft_entry__.smoke() and the tests.
"""
import math

import numpy as np

SEED = 31359  # main_avatar.py:817


def capsule_points(P, rng):
    """P points on a capsule of height 1.7 m, radius 0.15-0.25 m; first half front (z>0), second half back."""
    half = (P + 1) // 2
    n_rows = max(int(math.sqrt(half * 1.7 / 0.45)), 1)
    n_cols = (half + n_rows - 1) // n_rows
    v, u = np.meshgrid((np.arange(n_rows) + 0.5) / n_rows, (np.arange(n_cols) + 0.5) / n_cols * 2 - 1, indexing="ij")
    u, v = u.reshape(-1)[:half], v.reshape(-1)[:half]
    r = 0.15 + 0.10 * np.sin(np.pi * v) ** 2          # radius profile along the height
    cap = np.minimum(1.0, np.minimum(v, 1 - v) / 0.08)  # rounded ends
    r = r * np.sqrt(np.maximum(cap * (2 - cap), 1e-3))
    x = u * r
    y = (v - 0.5) * 1.7
    z = np.sqrt(np.maximum(1 - u * u, 0.0)) * r * 0.6
    front = np.stack([x, y, z], 1)
    back = np.stack([x, y, -z], 1)[: P - half]
    pts = np.concatenate([front, back], 0).astype(np.float32)
    spacing = np.float32(math.sqrt(1.7 * 0.45 / max(half, 1)))
    return pts, spacing


def make_gaussians(P, seed=SEED):
    """Posed-space Gaussian attributes standing in for U-Net outputs (raster-only configs)."""
    rng = np.random.default_rng(seed)
    pts, spacing = capsule_points(P, rng)
    xyz = pts + rng.normal(0, 0.005, (P, 3)).astype(np.float32)
    scales = (spacing * np.exp(rng.normal(0, 0.3, (P, 3)))).astype(np.float32)
    q = rng.normal(0, 1, (P, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opacity = (1 / (1 + np.exp(-rng.normal(1, 2, (P, 1))))).astype(np.float32)
    rgb = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    return dict(xyz=xyz.astype(np.float32), scales=scales, rotations=q, opacity=opacity, rgb=rgb, cano=pts)


def ring_cameras(V, radius=2.5, img=1024, focal=1100.0):
    """V cameras on a ring, built like calc_free_mv (utils/visualize_util.py:133-162):
    extr = T(0,0,radius) @ RotY(2*pi*k/V) @ RotX(pi)."""
    K = np.array([[focal, 0, img / 2], [0, focal, img / 2], [0, 0, 1]], np.float32)
    extrs = []
    for k in range(V):
        a = 2 * math.pi * k / V
        rot_y = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]], np.float32)
        rot_x = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]], np.float32)  # Rodrigues([pi,0,0])
        E = np.identity(4, np.float32)
        E[:3, :3] = rot_y @ rot_x
        E[:3, 3] = [0, 0, radius]
        extrs.append(E)
    return extrs, [K.copy() for _ in range(V)]


def make_skinning(cano_pts, J=55, seed=SEED):
    """Dense (N,J) LBS weights (softmax of distance to random joint centres, top-4 kept, renormalised) and
    (J,4,4) cano->live joint matrices (random small rotations about the centres)."""
    rng = np.random.default_rng(seed + 1)
    N = cano_pts.shape[0]
    centres = np.stack([rng.uniform(-0.2, 0.2, J), rng.uniform(-0.85, 0.85, J), rng.uniform(-0.1, 0.1, J)], 1).astype(np.float32)
    w = np.zeros((N, J), np.float32)
    chunk = 65536
    for s in range(0, N, chunk):
        d2 = ((cano_pts[s:s + chunk, None, :] - centres[None]) ** 2).sum(-1)
        logits = -d2 / (2 * 0.1 ** 2)
        logits -= logits.max(1, keepdims=True)
        e = np.exp(logits)
        idx = np.argpartition(-e, 4, axis=1)[:, :4]
        top = np.take_along_axis(e, idx, 1)
        top /= top.sum(1, keepdims=True)
        ww = np.zeros_like(e)
        np.put_along_axis(ww, idx, top, 1)
        w[s:s + chunk] = ww
    mats = np.zeros((J, 4, 4), np.float32)
    for j in range(J):
        rv = rng.normal(0, 0.3, 3)
        th = np.linalg.norm(rv) + 1e-12
        k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * Kx @ Kx
        mats[j, :3, :3] = R
        mats[j, :3, 3] = centres[j] - R @ centres[j] + rng.normal(0, 0.01, 3)
        mats[j, 3, 3] = 1
    return w, mats
