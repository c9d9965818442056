"""render3's camera setup (gaussian_renderer.py:44-52, graphics_utils.py:51-85): host mirror vs golden values produced by
the unmodified reference functions (tests/golden/make_camera_golden.py)."""
import os

import numpy as np

from animatablegaussians_b200 import camera
from tests import util
from tests.golden.make_camera_golden import cases


def test_camera_block_matches_reference_golden():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera.npz"))
    for i, (E, K, W, H) in enumerate(cases()):
        cb = camera.camera_block(E, K, W, H)
        util.assert_close("view", cb["viewmatrix"], z["view%d" % i], 1e-6)
        util.assert_close("proj", cb["projmatrix"], z["proj%d" % i], 2e-6)
        util.assert_close("campos", cb["campos"], z["campos%d" % i], 1e-5)
        assert abs(cb["tanfovx"] - z["tan%d" % i][0]) < 1e-6 and abs(cb["tanfovy"] - z["tan%d" % i][1]) < 1e-6


def test_projection_reduces_to_pinhole():
    """SURVEY.md §8a cheat-sheet: K-aware projection + ndc2Pix == fx*x/z + cx - 0.5."""
    E, K, W, H = cases()[0]
    cb = camera.camera_block(E, K, W, H)
    rng = np.random.default_rng(1)
    pc = rng.normal(0, 1, (50, 3)); pc[:, 2] = np.abs(pc[:, 2]) + 1.0          # camera-space points
    pw = (np.linalg.inv(E) @ np.c_[pc, np.ones(50)].T).T                          # world-space, homogeneous
    hom = pw @ cb["projmatrix"].astype(np.float64)
    ndc = hom[:, :2] / hom[:, 3:4]
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    assert np.allclose(px, K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 2] - 0.5, atol=2e-3)
    assert np.allclose(py, K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2] - 0.5, atol=2e-3)
