"""BASELINE config 0: SMPL-X LBS (smplx/lbs.py) — oracle pinned by the reference's own output; kernel vs both."""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.golden.make_smpl_lbs_golden import synthetic_model

GOLD = os.path.join(os.path.dirname(__file__), "golden", "smpl_lbs.npz")


def test_oracle_matches_reference_golden_cpu():
    from oracle import smpl_lbs_oracle as so
    z = np.load(GOLD)
    m = synthetic_model()
    verts, joints, A = so.lbs(m["betas"], m["pose"], m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"],
                              m["parents"], m["lbs_weights"])
    util.assert_close("verts", verts[::10].numpy(), z["verts"], 1e-5)
    util.assert_close("joints", joints.numpy(), z["joints"], 1e-5)
    util.assert_close("A", A.numpy(), z["A"], 1e-5)


@pytest.mark.gpu
def test_kernel_matches_reference_golden(built_lib):
    from animatablegaussians_b200 import smpl_lbs
    z = np.load(GOLD)
    m = {k: v.cuda() for k, v in synthetic_model().items()}
    verts, joints, A = smpl_lbs.lbs(m["betas"], m["pose"], m["v_template"][None], m["shapedirs"], m["posedirs"], m["J_regressor"],
                                    m["parents"], m["lbs_weights"], pose2rot=True, return_affine_mat=True)
    util.assert_close("verts", verts[0, ::10].cpu().numpy(), z["verts"], 1e-4)
    util.assert_close("joints", joints[0].cpu().numpy(), z["joints"], 1e-4)
    util.assert_close("A", A[0].cpu().numpy(), z["A"], 1e-4)
    # pose given as rotation matrices (pose2rot=False) takes the same path
    rot, _, _ = smpl_lbs.joint_chain(m["pose"].view(-1, 3), joints[0], m["parents"])
    v2, j2 = smpl_lbs.lbs(m["betas"], rot.reshape(1, -1, 9), m["v_template"][None], m["shapedirs"], m["posedirs"], m["J_regressor"],
                          m["parents"], m["lbs_weights"], pose2rot=False)
    util.assert_close("verts(rotmat)", v2.cpu().numpy(), verts.cpu().numpy(), 1e-5)
