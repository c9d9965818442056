"""CPU (-m "not gpu") checks of the oracle itself: structural properties the reference guarantees
(SURVEY.md Appendix B.15, §7 hard part (a)) and the golden vectors produced by the UNMODIFIED
reference CUDA kernels on a B200 (tests/golden/make_raster_golden.py)."""
import glob
import os

import numpy as np
import pytest

from tests import util

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_empty_scene_is_background():
    sc = util.random_scene(0, 40, 24, seed=1)
    out = util.run_oracle(sc)
    for c in range(3):
        assert np.allclose(out["color"][c], sc["bg"][c])
    assert np.all(out["alpha"] == 0) and np.all(out["depth"] == 0) and out["R"] == 0


def test_all_behind_camera_culled():
    sc = util.random_scene(64, 32, 32, seed=2, behind_frac=1.0)
    out = util.run_oracle(sc, util.upstream_grads(32, 32, 2))
    assert np.all(out["radii"] == 0) and out["R"] == 0
    for k, v in out["grads"].items():
        assert not np.any(v), k  # culled Gaussians receive exactly-zero gradients (backward.cu:156,369)


def test_alpha_identity_and_bounds():
    sc = util.random_scene(800, 70, 50, seed=3)
    out = util.run_oracle(sc)
    a = out["alpha"]
    assert a.min() >= 0 and a.max() <= 1.0 + 1e-5
    # colour = sum(c_i a_i T_i) + T_final*bg with T_final = 1 - alpha up to rounding  (forward.cu:377-378)
    sc2 = dict(sc)
    sc2["bg"] = np.zeros(3, np.float32)
    out2 = util.run_oracle(sc2)
    recon = out2["color"] + (1 - a) * sc["bg"][:, None, None]
    assert np.abs(recon - out["color"]).max() < 1e-5


def test_backward_matches_finite_differences_on_colors():
    # colour enters linearly: dL/dc is exact, so a directional finite difference must agree tightly
    sc = util.random_scene(300, 48, 40, seed=4)
    g = util.upstream_grads(48, 40, 4)
    out = util.run_oracle(sc, g)
    rng = np.random.default_rng(0)
    d = rng.normal(0, 1, sc["rgb"].shape).astype(np.float32)
    eps = 1e-2
    scp = dict(sc); scp["rgb"] = sc["rgb"] + eps * d
    scm = dict(sc); scm["rgb"] = sc["rgb"] - eps * d
    Lp = (util.run_oracle(scp)["color"].astype(np.float64) * g[0]).sum()
    Lm = (util.run_oracle(scm)["color"].astype(np.float64) * g[0]).sum()
    fd = (Lp - Lm) / (2 * eps)
    an = (out["grads"]["colors"].astype(np.float64) * d).sum()
    assert abs(fd - an) <= 2e-3 * max(abs(an), 1.0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "raster_*.npz"))) or [None])
def test_oracle_matches_reference_golden(path):
    if path is None:
        pytest.skip("no golden vectors committed yet")
    z = np.load(path, allow_pickle=True)
    kw = {k: (None if z[k].dtype == object and z[k].item() is None else z[k]) for k in z["scene_keys"]}
    sc = util.random_scene(**{k: (v.item() if hasattr(v, "item") and np.ndim(v) == 0 else v) for k, v in kw.items()})
    g = (z["g_color"], z["g_depth"], z["g_alpha"])
    # backward starts from the reference's own forward alpha (see RasterOracle.backward)
    out = util.run_oracle(sc, g, alpha_from=z["alpha"])
    assert np.array_equal(out["radii"], z["radii"])
    assert out["R"] == int(z["num_rendered"])
    util.assert_close("color", out["color"], z["color"])
    util.assert_close("depth", out["depth"], z["depth"])
    util.assert_close("alpha", out["alpha"], z["alpha"])
    for k in ("means3D", "means2D", "colors", "opacity", "scales", "rotations", "cov3D", "sh"):
        gk = "grad_" + k
        if gk in z and z[gk].size and out["grads"][k].size:
            util.assert_close(gk, out["grads"][k], z[gk])


def test_oracle_backward_matches_float64_autograd():
    """Independent pin of the hand-derived backward (backward.cu:144-601): a float64 autograd restatement of
    the FORWARD semantics must produce the same gradients as the C oracle's restated backward."""
    import torch
    from oracle import raster_autograd as ra
    W, H = 32, 32
    sc = util.random_scene(40, W, H, seed=11, behind_frac=0.1, big_frac=0.1)
    g = util.upstream_grads(W, H, 11)
    out = util.run_oracle(sc, g)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    m, c, o, s, q = t64(sc["xyz"]), t64(sc["rgb"]), t64(sc["opacity"]), t64(sc["scales"]), t64(sc["rotations"])
    color, depth, alpha = ra.render(m, c, o, s, q, sc["viewmatrix"], sc["projmatrix"], sc["tanfovx"], sc["tanfovy"],
                                    H, W, sc["bg"])
    util.assert_close("color", out["color"], color.detach().numpy(), 2e-5)
    util.assert_close("depth", out["depth"], depth.detach().numpy(), 2e-5)
    util.assert_close("alpha", out["alpha"], alpha.detach().numpy(), 2e-5)
    L = (color * torch.tensor(g[0], dtype=torch.float64)).sum() + (depth * torch.tensor(g[1], dtype=torch.float64)).sum() + \
        (alpha * torch.tensor(g[2], dtype=torch.float64)).sum()
    L.backward()
    util.assert_close("grad_colors", out["grads"]["colors"], c.grad.numpy(), 1e-4)
    util.assert_close("grad_opacity", out["grads"]["opacity"], o.grad.numpy(), 1e-4)
    util.assert_close("grad_means3D", out["grads"]["means3D"], m.grad.numpy(), 1e-3)
    util.assert_close("grad_scales", out["grads"]["scales"], s.grad.numpy(), 1e-3)
    util.assert_close("grad_rotations", out["grads"]["rotations"], q.grad.numpy(), 1e-3)
