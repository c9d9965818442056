"""DualStyleUNet: state_dict compatibility and numerical parity with the UNMODIFIED reference module
(golden vectors generated on CPU by tests/golden/make_styleunet_golden.py; fp32, tolerance 1e-4 rel max-norm)."""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.golden import make_styleunet_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(cfg):
    from animatablegaussians_b200.styleunet import DualStyleUNet
    torch.manual_seed(0)
    net = DualStyleUNet(**cfg)
    G.fill_state(net)
    return net


def _check(name, device, tol=1e-4):
    from animatablegaussians_b200 import styleunet_ops as ops
    ops.set_compute_dtype(torch.float32)
    cfg, use_view, stride = G.CASES[name]
    z = np.load(os.path.join(GOLD, "styleunet_%s.npz" % name))
    net = _build(cfg).to(device)
    cond, style, vf1, vf2, up = (t.to(device) if t is not None else None for t in G.inputs(cfg, use_view))
    cond.requires_grad_(True)
    out, _ = net([style], cond, randomize_noise=False, view_feature1=vf1, view_feature2=vf2)
    (out * up).sum().backward()
    util.assert_close(name + ":out", out.detach().cpu().numpy()[..., ::stride, ::stride], z["out"], tol)
    assert abs(float(out.detach().double().mean()) - float(z["out_mean"])) <= tol * max(abs(float(z["out_std"])), 1e-6)
    util.assert_close(name + ":grad_cond", cond.grad.cpu().numpy(), z["grad_cond"], 10 * tol)
    named = dict(net.named_parameters())
    for k in G.GRAD_KEYS:
        if "grad:" + k in z.files:
            util.assert_close(name + ":grad:" + k, named[k].grad.cpu().numpy(), z["grad:" + k], 10 * tol)
    return net


def test_state_dict_keys_match_reference_listing():
    """363 entries per net = 216 parameters + 147 buffers for the full-size net."""
    from animatablegaussians_b200.styleunet import DualStyleUNet
    net = DualStyleUNet(inp_size=512, inp_ch=3, out_ch=3, out_size=1024, style_dim=512, n_mlp=2)
    sd = net.state_dict()
    assert len(sd) == 363 and len(list(net.parameters())) == 216
    n_par = sum(p.numel() for p in net.parameters())
    assert abs(n_par / 1e6 - 74.48) < 0.01
    for k in ("style.1.weight", "conv_in.0.kernel", "conv_in.1.weight", "conv_in.2.bias", "from_rgbs.0.downsample.kernel",
              "from_rgbs.0.conv.0.weight", "from_rgbs.0.conv.1.bias", "cond_convs.0.conv1.0.weight",
              "cond_convs.0.conv2.0.kernel", "cond_convs.0.conv2.1.weight", "comb_convs.0.0.weight", "comb_convs.0.1.bias",
              "convs1.0.conv.weight", "convs1.0.conv.blur.kernel", "convs1.0.conv.modulation.weight", "convs1.0.noise.weight",
              "convs1.0.activate.bias", "to_rgbs1.0.bias", "to_rgbs1.0.conv.weight", "to_rgbs1.0.upsample.kernel",
              "to_rgbs1.0.iwt.ll", "to_rgbs1.0.dwt.hh", "noises.noise_11", "iwt.lh"):
        assert k in sd, k
    assert sd["convs1.0.conv.weight"].shape == (1, 512, 512, 3, 3) and sd["noises.noise_11"].shape == (1, 1, 512, 512)


@pytest.mark.parametrize("name", ["small", "small8"])
def test_matches_reference_golden_cpu(name):
    _check(name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "small8", "full_view"])
def test_matches_reference_golden_gpu(name, built_lib):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    _check(name, "cuda", tol=2e-4)


@pytest.mark.gpu
def test_prefix_tail_split_is_exact(built_lib):
    """forward_prefix + forward_view_tail == forward (the per-pose prefix cache of SURVEY.md §7 (f))."""
    from animatablegaussians_b200 import styleunet_ops as ops
    ops.set_compute_dtype(torch.float32)
    cfg, use_view, _ = G.CASES["full_view"]
    net = _build(cfg).cuda()
    cond, style, vf1, vf2, _ = (t.cuda() if t is not None else None for t in G.inputs(cfg, True))
    with torch.no_grad():
        full, _ = net([style], cond, randomize_noise=False, view_feature1=vf1, view_feature2=vf2)
        pre = net.forward_prefix([style], cond)
        split = net.forward_view_tail(pre, vf1, vf2)
    util.assert_close("split", split.cpu().numpy(), full.cpu().numpy(), 1e-6)
