"""SURVEY.md §8f rank 4 + rank 2 host logic (CPU): PLY export / import in the reference's layout, trainer checkpoints, the
optimizer state in torch.optim.Adam's layout, the cosine schedule."""
import math
import os

import numpy as np
import torch

from animatablegaussians_b200 import formats, optim


def _vals(P=7):
    g = torch.Generator().manual_seed(0)
    return {"positions": torch.randn(P, 3, generator=g), "colors": torch.rand(P, 3, generator=g),
            "opacity": torch.rand(P, 1, generator=g) * 0.9 + 0.05, "scales": torch.rand(P, 3, generator=g) * 0.02 + 1e-3,
            "rotations": torch.nn.functional.normalize(torch.randn(P, 4, generator=g))}


def test_ply_layout_known_answer_and_round_trip(tmp_path):
    vals = _vals()
    path = os.path.join(tmp_path, "sub", "posed.ply")
    formats.save_gaussians_as_ply(path, vals)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 7"]
    props = [l.split()[-1] for l in lines if l.startswith("property float ")]
    assert props == formats.gaussian_ply_fields() and len(props) == 62          # 3+3+3+45+1+3+4 (gaussians/obj_io.py:9-21)
    assert len(body) == 7 * 62 * 4
    v = formats.read_ply_vertices(path)
    # known answers for Gaussian 0: BGR swap + RGB2SH, logit opacity, log scale, zero normals / higher SH
    c = vals["colors"][0].numpy()
    assert np.allclose([v["f_dc_0"][0], v["f_dc_1"][0], v["f_dc_2"][0]], (c[[2, 1, 0]] - 0.5) / 0.28209479177387814, atol=1e-6)
    o = float(vals["opacity"][0])
    assert abs(v["opacity"][0] - math.log(o / (1 - o))) < 1e-5
    assert np.allclose([v["scale_%d" % i][0] for i in range(3)], np.log(vals["scales"][0].numpy()), atol=1e-6)
    assert all(float(np.abs(v[n]).max()) == 0 for n in ["nx", "ny", "nz"] + ["f_rest_%d" % i for i in range(45)])
    back = formats.load_gaussians_from_ply(path)
    for k in ("positions", "colors", "opacity", "scales", "rotations"):
        assert torch.allclose(back[k], vals[k].float(), atol=2e-6), k
    assert back["features_extr"].shape == (7, 3, 15)


def test_checkpoint_files_round_trip(tmp_path):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    net(torch.randn(5, 4)).sum().backward(); opt.step()
    formats.save_ckpt(str(tmp_path), net, opt, epoch_idx=3, iter_idx=1234)
    assert sorted(os.listdir(tmp_path)) == ["net.pt", "optm.pt"]
    assert set(torch.load(os.path.join(tmp_path, "net.pt"), weights_only=False)) == {"epoch_idx", "iter_idx", "avatar_net"}
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    flat = optim.FlatAdam(net2.parameters(), lr=1e-3)
    assert formats.load_ckpt(str(tmp_path), net2, flat) == (3, 1234)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    assert flat.lr == 5e-4 and flat.t == 1


def test_flat_adam_state_dict_is_torch_adam_layout():
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    net[0].bias.requires_grad_(False)                       # torch keeps frozen parameters in the index space
    ref = torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.8, 0.95), eps=1e-7)
    for _ in range(3):
        ref.zero_grad(); net(torch.randn(4, 6)).pow(2).sum().backward(); ref.step()
    sd = ref.state_dict()
    net2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    net2[0].bias.requires_grad_(False)
    flat = optim.FlatAdam(net2.parameters())
    flat.load_state_dict(sd)
    assert flat.t == 3 and flat.lr == 2e-4 and flat.betas == (0.8, 0.95) and flat.eps == 1e-7
    out = flat.state_dict()
    assert sorted(out["state"]) == sorted(sd["state"]) == [0, 2, 3]
    for i in sd["state"]:
        assert float(out["state"][i]["step"]) == 3.0
        assert torch.equal(out["state"][i]["exp_avg"], sd["state"][i]["exp_avg"])
        assert torch.equal(out["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"])
    assert out["param_groups"][0]["params"] == [0, 1, 2, 3]
    fresh = torch.optim.Adam(net2.parameters())
    fresh.load_state_dict(out)                               # and torch's own optimizer accepts ours
    assert fresh.param_groups[0]["lr"] == 2e-4
    # the trainer's update_lr() idiom (main_avatar.py:61-68)
    for group in flat.param_groups:
        group["lr"] = optim.cosine_lr(5e-4, 400000, 800000)
    assert abs(flat.lr - 5e-4 * (0.5 * 0.95 + 0.05)) < 1e-12
    assert abs(optim.cosine_lr(5e-4, 0, 800000) - 5e-4) < 1e-15 and abs(optim.cosine_lr(5e-4, 800000, 800000) - 2.5e-5) < 1e-12


# ---- OpenEXR position maps (dataset/dataset_mv_rgb.py:146-151) -----------------------------------------------------------
def test_exr_reader_matches_opencv_golden():
    """formats.read_exr vs what OpenCV — the reference's reader — decoded from the same files (fixtures written and read back
    by tests/golden/make_exr_golden.py): FLOAT / HALF, NONE / RLE / ZIPS / ZIP, raw (incompressible) blocks, a ragged last
    block, one channel.  Bit-exact."""
    import os
    import numpy as np
    from animatablegaussians_b200 import formats
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = np.load(os.path.join(gold_dir, "exr_golden.npz"))
    assert len(gold.files) >= 7
    for k in gold.files:
        got = formats.read_exr(os.path.join(gold_dir, "exr_%s.exr" % k))
        assert got.dtype == np.float32 and got.shape == gold[k].shape, k
        assert np.array_equal(got, gold[k]), k


def test_exr_reader_matches_live_opencv(tmp_path):
    import os
    import numpy as np
    import pytest
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    from animatablegaussians_b200 import formats
    rng = np.random.default_rng(11)
    img = (rng.normal(size=(64, 128, 3)) * (rng.random((64, 128, 1)) > 0.4)).astype(np.float32)
    path = str(tmp_path / "00000000.exr")
    if not cv2.imwrite(path, img):
        pytest.skip("this OpenCV build cannot write OpenEXR")
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert np.array_equal(formats.read_exr(path), ref)
    # the dataset's split of the (H, 2H, 3) map into the (6, H, H) network input
    want = np.concatenate([ref[:, :64], ref[:, 64:]], 2).transpose(2, 0, 1)
    assert np.array_equal(formats.load_smpl_pos_map(path).numpy(), want)


def test_exr_reader_rejects_what_it_cannot_read(tmp_path):
    import pytest
    from animatablegaussians_b200 import formats
    p = tmp_path / "bad.exr"
    p.write_bytes(b"not an exr file at all")
    with pytest.raises(ValueError):
        formats.read_exr(str(p))
