"""DualStyleUNet: state_dict compatibility and numerical parity with the UNMODIFIED reference module
(golden vectors generated on CPU by tests/golden/make_styleunet_golden.py; fp32, tolerance 1e-4 rel max-norm)."""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.golden import make_styleunet_golden as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(cfg, oracle=False):
    if oracle:
        from oracle.styleunet_oracle import DualStyleUNet
    else:
        from animatablegaussians_b200.styleunet import DualStyleUNet
    torch.manual_seed(0)
    net = DualStyleUNet(**cfg)
    G.fill_state(net)
    return net


def _check(name, device, tol=1e-4, oracle=False):
    if oracle:
        from oracle import styleunet_oracle as so
        so.set_compute_dtype(torch.float32)
    else:
        from animatablegaussians_b200 import styleunet_ops as ops
        ops.set_compute_dtype(torch.float32)
    cfg, use_view, stride = G.CASES[name]
    z = np.load(os.path.join(GOLD, "styleunet_%s.npz" % name))
    net = _build(cfg, oracle).to(device)
    cond, style, vf1, vf2, up = (t.to(device) if t is not None else None for t in G.inputs(cfg, use_view))
    cond.requires_grad_(True)
    out, _ = net([style], cond, randomize_noise=False, view_feature1=vf1, view_feature2=vf2)
    (out * up).sum().backward()
    util.assert_close(name + ":out", out.detach().cpu().numpy()[..., ::stride, ::stride], z["out"], tol)
    assert abs(float(out.detach().double().mean()) - float(z["out_mean"])) <= tol * max(abs(float(z["out_std"])), 1e-6)
    # Whole-network gradients cross ~40 leaky-ReLU kinks: a pre-activation within rounding of 0 flips its slope
    # between two fp32 implementations (CPU reference vs GPU), which moves individual gradient entries by O(1e-2)
    # while leaving the bulk untouched.  Same-device runs agree to 1e-3 in max-norm; across devices the robust
    # measure is the relative L2 error.  The strict per-operator gradient checks are the tests below.
    def l2(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    gtol = 10 * tol if device == "cpu" else 2e-2
    err = (util.rel_err if device == "cpu" else l2)
    assert err(cond.grad.cpu().numpy(), z["grad_cond"]) <= gtol, name + ":grad_cond"
    named = dict(net.named_parameters())
    for k in G.GRAD_KEYS:
        if "grad:" + k in z.files:
            assert err(named[k].grad.cpu().numpy(), z["grad:" + k]) <= gtol, name + ":grad:" + k
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
def test_oracle_matches_reference_golden_cpu(name):
    """Pins oracle/styleunet_oracle.py (the per-operator checker of the CUDA kernels) to the real reference."""
    _check(name, "cpu", oracle=True)


def test_product_has_no_cpu_path():
    from animatablegaussians_b200 import styleunet_ops as ops
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.haar_dwt(torch.zeros(1, 4, 8, 8))


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


# ------------------------------------------------------------------------------------------ per-operator parity
def _cmp(a, b, tol):
    util.assert_close("op", a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy(), tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("C,H,W,kind", [(64, 32, 48, "blur1"), (3, 33, 17, "down"), (12, 16, 16, "up"), (128, 9, 9, "blur2"),
                                         (32, 20, 12, "up"), (8, 15, 15, "blur1"),
                                         (16, 517, 600, "blur1"), (64, 300, 300, "blur2")])   # last two: blur_strip_kernel, ragged last strip
def test_upfirdn2d_matches_oracle(C, H, W, kind, dtype, tol, built_lib):
    from animatablegaussians_b200 import styleunet_ops as ops
    from oracle import styleunet_oracle as so
    k = so.make_kernel([1, 3, 3, 1]).cuda()
    kd = k.double()
    cfg = {"blur1": dict(kernel=k * 4, up=1, down=1, pad=(1, 1)), "blur2": dict(kernel=k, up=1, down=1, pad=(2, 2)),
           "up": dict(kernel=k * 4, up=2, down=1, pad=(2, 1)), "down": dict(kernel=k, up=1, down=2, pad=(1, 1))}[kind]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(1, C, H, W, device="cuda", generator=g)
    xa = x.to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xb = x.to(dtype).double().requires_grad_(True)
    ya = ops.upfirdn2d(xa, **cfg)
    yb = so.upfirdn2d(xb, **dict(cfg, kernel=cfg['kernel'].double()))
    _cmp(ya, yb, tol)
    up = torch.randn(yb.shape, device="cuda", generator=g)
    ya.backward(up.to(dtype)); yb.backward(up.to(dtype).double())
    _cmp(xa.grad, xb.grad, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])  # oracle = cuDNN fp32 conv (may pick Winograd)
@pytest.mark.parametrize("C", [3, 8, 12, 64])
def test_haar_and_wavelet_upsample_match_oracle(C, dtype, tol, built_lib):
    from animatablegaussians_b200 import styleunet_ops as ops
    from oracle import styleunet_oracle as so
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(1, C, 24, 16, device="cuda", generator=g).to(dtype)
    xa = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xb = x.double().requires_grad_(True)
    ya, yb = ops.haar_dwt(xa), so.haar_dwt(xb)
    _cmp(ya, yb, tol)
    _cmp(ops.haar_iwt(ya), so.haar_iwt(yb), tol)
    _cmp(ops.haar_iwt(ya), x, 2 * tol)  # orthonormal round trip
    k = (so.make_kernel([1, 3, 3, 1]) * 4).cuda()
    if C % 4 == 0:
        za, zb = ops.wavelet_upsample(xa, k), so.wavelet_upsample(xb, k.double())
        _cmp(za, zb, 2 * tol)
        up = torch.randn(zb.shape, device="cuda", generator=g)
        za.backward(up.to(dtype)); zb.backward(up.to(dtype).double())
        _cmp(xa.grad, xb.grad, 3 * tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("C,H,activate,noise", [(512, 16, True, True), (64, 64, True, False), (12, 32, False, False), (3, 8, True, True)])
def test_bias_act_matches_oracle(C, H, activate, noise, dtype, tol, built_lib):
    from animatablegaussians_b200 import styleunet_ops as ops
    from oracle import styleunet_oracle as so
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(1, C, H, H, device="cuda", generator=g).to(dtype)
    ba = torch.randn(C, device="cuda", generator=g).requires_grad_(True)
    bb = ba.detach().double().requires_grad_(True)
    nz = torch.randn(1, 1, H, H, device="cuda", generator=g) if noise else None
    wa = torch.tensor([0.37], device="cuda", requires_grad=True) if noise else None
    wb = torch.tensor([0.37], device="cuda", dtype=torch.float64, requires_grad=True) if noise else None
    xa = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xb = x.double().requires_grad_(True)
    ya = ops.bias_act(xa, ba, nz, wa, activate)
    yb = so.bias_act(xb, bb, nz.double() if noise else None, wb, activate)
    _cmp(ya, yb, tol)
    up = torch.randn(yb.shape, device="cuda", generator=g)
    ya.backward(up.to(dtype)); yb.backward(up.to(dtype).double())
    _cmp(xa.grad, xb.grad, tol)
    _cmp(ba.grad, bb.grad, 5 * tol)
    if noise:
        _cmp(wa.grad, wb.grad, 5 * tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("Cout,Cin,k,demod,tr", [(64, 32, 3, True, False), (12, 64, 1, False, False), (32, 48, 3, True, True), (128, 128, 3, False, False),
                                                 (16, 1024, 3, True, False), (16, 3, 3, True, False), (8, 2048, 3, True, True)])  # smem row / unaligned row / row > 48 KB
def test_modweight_matches_oracle(Cout, Cin, k, demod, tr, dtype, tol, built_lib):
    from animatablegaussians_b200 import styleunet_ops as ops
    from oracle import styleunet_oracle as so
    g = torch.Generator(device="cuda").manual_seed(3)
    w = torch.randn(1, Cout, Cin, k, k, device="cuda", generator=g)
    s = 1 + 0.3 * torch.randn(1, Cin, device="cuda", generator=g)
    scale = 1 / (Cin * k * k) ** 0.5
    wa, sa = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
    wb, sb = w.double().requires_grad_(True), s.double().requires_grad_(True)
    oa, ha = ops.mod_weight(wa, sa, scale, demod, dtype, transpose_io=tr)   # operand + fp32 gradient handle
    ob = so.prepare_modulated_weight(wb, sb, scale, demod)[0]
    if tr:
        ob = ob.transpose(0, 1)
    _cmp(oa, ob, tol)
    up = torch.randn(ob.shape, device="cuda", generator=g)
    ha.backward(up.to(dtype).float()); ob.backward(up.to(dtype).double())
    _cmp(wa.grad, wb.grad, tol)
    _cmp(sa.grad, sb.grad, 5 * tol)


def _ref_conv(x, w, k, stride, pad, transposed):
    if transposed:
        return torch.nn.functional.conv_transpose2d(x, w.transpose(0, 1), stride=stride, padding=pad)
    return torch.nn.functional.conv2d(x, w, stride=stride, padding=pad)


_CONV_CASES = [  # N, H, W, Cin, Cout, k, stride, pad, transposed, act, noise
    (1, 16, 16, 64, 64, 3, 1, 1, False, 1, True), (1, 8, 16, 128, 128, 3, 1, 1, False, 0, False), (1, 32, 64, 512, 256, 3, 1, 1, False, 1, False),
    (1, 64, 64, 64, 128, 1, 1, 0, False, 0, False), (1, 128, 128, 128, 64, 3, 1, 1, False, 1, True), (1, 24, 48, 1024, 512, 3, 1, 1, False, 1, False),
    (1, 256, 256, 64, 64, 3, 1, 1, False, 0, True), (2, 128, 256, 128, 128, 3, 1, 1, False, 0, False),   # no lrelu on 4M outputs: kink flips dominate the max-norm
    (16, 512, 512, 64, 64, 3, 1, 1, False, 0, False),                                                     # the benched top level: 64->64 @512^2 x 16 views
    (1, 8, 8, 512, 512, 3, 1, 1, False, 1, False),                                                        # 8x8 level: tile wider than the map
    (1, 33, 33, 128, 256, 3, 2, 0, False, 1, False), (1, 129, 129, 256, 512, 3, 2, 0, False, 0, False), (1, 17, 17, 512, 512, 3, 2, 0, False, 1, False),   # blur -> stride 2
    (1, 8, 8, 512, 512, 3, 2, 0, True, 0, False), (2, 32, 32, 128, 64, 3, 2, 0, True, 0, False), (1, 64, 64, 512, 256, 3, 2, 0, True, 0, False),        # transposed stride 2
    (4, 64, 64, 64, 128, 4, 2, 1, False, 0, False), (4, 128, 128, 1, 64, 4, 2, 1, False, 2, False),                                                     # viewdir_net
    (1, 65, 65, 3, 128, 3, 2, 0, False, 1, False), (1, 32, 32, 3, 128, 1, 1, 0, False, 1, False),                                                       # 3-channel inputs
    (2, 32, 32, 64, 12, 1, 1, 0, False, 0, False), (1, 16, 16, 512, 32, 1, 1, 0, False, 0, False),                                                      # ToRGB
]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad,transposed,act,noise", _CONV_CASES)
def test_conv_matches_fp32_reference(N, H, W, Cin, Cout, k, stride, pad, transposed, act, noise, dtype, built_lib):
    """Every layer geometry of the path (include/agr_conv.h) — tcgen05 implicit GEMM (bf16, wide channels) or the CUDA-core
    kernels (fp32 / narrow layers) — forward, data gradient, weight gradient, bias / noise-weight gradients vs an fp32
    convolution of the SAME (bf16-rounded) operands: only the final rounding of the outputs differs."""
    from animatablegaussians_b200 import styleunet_ops as ops
    if dtype == torch.float32 and N * H * W * Cin * Cout > (1 << 31):
        pytest.skip("fp32 CUDA-core path: covered at smaller sizes")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).to(dtype).contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, device="cuda", generator=g)
    geom = ops.conv_geom(x.shape, Cout, k, stride, pad, transposed)
    wide = Cin % 64 == 0 and Cout % 64 == 0
    assert [ops.conv_path(x, geom, i) for i in range(3)] == [1 if (wide and dtype == torch.bfloat16) else 2] * 3
    nz = torch.randn(1, 1, geom.OH, geom.OW, device="cuda", generator=g) if noise else None
    nw = torch.tensor([0.5], device="cuda", requires_grad=True) if noise else None
    xa, wa, ba = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ya = ops.conv2d(xa, wa, None, bias=ba, noise=nz, noise_weight=nw, activate=act, stride=stride, pad=pad, transposed=transposed)
    xb, wb, bb = x.float().requires_grad_(True), w.float().requires_grad_(True), b.clone().requires_grad_(True)
    nwb = nw.detach().clone().requires_grad_(True) if noise else None
    yb = _ref_conv(xb, wb, k, stride, pad, transposed)
    if noise:
        yb = yb + nwb * nz
    yb = yb + bb.view(1, -1, 1, 1)
    if act:
        yb = torch.nn.functional.leaky_relu(yb, 0.2) * (2 ** 0.5 if act == 1 else 1.0)
    bf = dtype == torch.bfloat16
    _cmp(ya, yb, 8e-3 if bf else 2e-5)
    up = torch.randn(yb.shape, device="cuda", generator=g).to(dtype)
    ya.backward(up)
    # reference backward starts from the product's own output sign pattern (lrelu kink) -> use float grads of yb
    yb.backward(up.float())
    tol = 2e-2 if bf else 1e-4
    _cmp(xa.grad, xb.grad, tol)
    _cmp(wa.grad, wb.grad, tol)
    _cmp(ba.grad, bb.grad, tol)
    if noise:
        _cmp(nw.grad, nwb.grad, tol)


@pytest.mark.gpu
def test_bf16_network_tracks_fp32(built_lib):
    """bf16 compute (tcgen05 convolutions + bf16 glue kernels) stays within bf16 noise of the fp32 path end to end."""
    from animatablegaussians_b200 import styleunet_ops as ops
    cfg, use_view, _ = G.CASES["small"]
    outs, grads = [], []
    for dt in (torch.float32, torch.bfloat16):
        ops.set_compute_dtype(dt)
        net = _build(cfg).cuda()
        cond, style, vf1, vf2, up = (t.cuda() if t is not None else None for t in G.inputs(cfg, use_view))
        cond.requires_grad_(True)
        out, _ = net([style], cond, randomize_noise=False)
        (out * up).sum().backward()
        outs.append(out.detach().double()); grads.append(net.convs1[1].conv.weight.grad.detach().double())
    ops.set_compute_dtype(torch.float32)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(outs[1], outs[0]) < 3e-2, rel(outs[1], outs[0])
    assert rel(grads[1], grads[0]) < 8e-2, rel(grads[1], grads[0])


@pytest.mark.gpu
@pytest.mark.parametrize("V,H,W,Ca,Cb,Cout", [(3, 16, 32, 128, 128, 128), (2, 32, 32, 64, 128, 64)])
def test_split_contraction_equals_conv_of_concat(V, H, W, Ca, Cb, Cout, built_lib):
    """conv(cat([a_v, b]), w) == conv(a_v, w[:, :Ca]) + conv(b, w[:, Ca:]) with b shared by the view batch."""
    from animatablegaussians_b200 import styleunet_ops as ops
    g = torch.Generator(device="cuda").manual_seed(5)
    bf = lambda t: t.to(torch.bfloat16)
    a = bf(torch.randn(V, Ca, H, W, device="cuda", generator=g)).contiguous(memory_format=torch.channels_last)
    b = bf(torch.randn(1, Cb, H, W, device="cuda", generator=g)).contiguous(memory_format=torch.channels_last)
    w = bf(torch.randn(Cout, Ca + Cb, 3, 3, device="cuda", generator=g) / ((Ca + Cb) * 9) ** 0.5).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(Cout, device="cuda", generator=g)
    a1, b1, w1, c1 = a.clone().requires_grad_(True), b.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y1 = ops._SplitConvAct.apply(a1, b1, w1, None, c1, 1)
    a2, b2, w2, c2 = a.float().requires_grad_(True), b.float().requires_grad_(True), w.float().requires_grad_(True), bias.clone().requires_grad_(True)
    y2 = torch.nn.functional.conv2d(torch.cat([a2, b2.expand(V, -1, -1, -1)], 1), w2, None, padding=1) + c2.view(1, -1, 1, 1)
    y2 = torch.nn.functional.leaky_relu(y2, 0.2) * 2 ** 0.5
    _cmp(y1, y2, 1.2e-2)
    up = bf(torch.randn(y2.shape, device="cuda", generator=g))
    y1.backward(up); y2.backward(up.float())
    _cmp(a1.grad, a2.grad, 2e-2); _cmp(b1.grad, b2.grad, 2e-2); _cmp(w1.grad, w2.grad, 2e-2); _cmp(c1.grad, c2.grad, 2e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1.2e-2)])
@pytest.mark.parametrize("V,Vb", [(3, 1), (2, 2)])
def test_add_view_feature_matches_interpolate(V, Vb, dtype, tol, built_lib):
    """out + F.interpolate(vf, 2x, 'bilinear') fused, incl. the adjoint resampling, vs ATen in float64."""
    from animatablegaussians_b200 import styleunet_ops as ops
    g = torch.Generator(device="cuda").manual_seed(6)
    vf = torch.randn(V, 16, 12, 10, device="cuda", generator=g)
    base = torch.randn(Vb, 16, 24, 20, device="cuda", generator=g).to(dtype)
    va, ba = vf.clone().requires_grad_(True), base.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ya = ops.add_view_feature(ba, va)
    vb, bb = vf.double().requires_grad_(True), base.double().requires_grad_(True)
    yb = bb + torch.nn.functional.interpolate(vb, (24, 20), mode="bilinear")
    _cmp(ya, yb, tol)
    up = torch.randn(yb.shape, device="cuda", generator=g).to(dtype)
    ya.backward(up); yb.backward(up.double())
    _cmp(va.grad, vb.grad, tol)
    _cmp(ba.grad, bb.grad, 2 * tol)


@pytest.mark.gpu
@pytest.mark.parametrize("out_dim,in_dim,bias,lr_mul", [(512, 512, True, 1.0), (64, 512, True, 1.0), (33, 70, False, 0.01), (3, 5, True, 0.5)])
def test_equal_linear_matches_reference_expression(out_dim, in_dim, bias, lr_mul, built_lib):
    """F.linear(x, W * scale, bias * lr_mul) (dual_styleunet.py:155-158) in float64 vs the fused style-vector kernels."""
    from animatablegaussians_b200 import styleunet_ops as ops
    g = torch.Generator(device="cuda").manual_seed(5)
    W = torch.randn(out_dim, in_dim, device="cuda", generator=g)
    b = torch.randn(out_dim, device="cuda", generator=g) if bias else None
    x = torch.randn(1, in_dim, device="cuda", generator=g)
    scale = lr_mul / in_dim ** 0.5
    Wa, xa = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ba = b.clone().requires_grad_(True) if bias else None
    Wb, xb = W.double().requires_grad_(True), x.double().requires_grad_(True)
    bb = b.double().requires_grad_(True) if bias else None
    with ops.step_arena():
        ya = ops.equal_linear(xa, Wa, ba, scale, lr_mul)
        yb = torch.nn.functional.linear(xb, Wb * scale, bias=bb * lr_mul if bias else None)
        _cmp(ya, yb, 1e-5)
        up = torch.randn(1, out_dim, device="cuda", generator=g)
        ya.backward(up); yb.backward(up.double())
    _cmp(Wa.grad, Wb.grad, 1e-5)
    _cmp(xa.grad, xb.grad, 1e-5)
    if bias:
        _cmp(ba.grad, bb.grad, 1e-6)


@pytest.mark.gpu
def test_step_arena_hands_out_disjoint_zeroed_slices(built_lib):
    from animatablegaussians_b200 import styleunet_ops as ops
    dev = torch.device("cuda:0")
    with ops.step_arena(chunk_floats=64):
        a, b = ops._zeros(5, dev), ops._zeros(7, dev)
        a.fill_(1.0); b.fill_(2.0)
        c = ops._zeros(60, dev)          # does not fit the open chunk: new chunk
        d = ops._zeros(100, dev)         # larger than a chunk: plain allocation
        assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and b.data_ptr() - a.data_ptr() == 32
        assert float(c.abs().sum()) == 0 and float(d.abs().sum()) == 0 and float(a.sum()) == 5 and float(b.sum()) == 14
    assert ops._arena is None
    e = ops._zeros(3, dev)
    assert float(e.abs().sum()) == 0


def test_wavelet_upsample_banks_reproduce_the_chain():
    """CPU: the parity-dependent filter banks agr_wavelet_upsample takes (host-derived) against the oracle's
    dwt(upsample(iwt(.))) chain in float64 -- forward and adjoint, borders included."""
    import numpy as np
    from animatablegaussians_b200 import styleunet_ops as ops
    from oracle import styleunet_oracle as so
    k = (so.make_kernel([1, 3, 3, 1]) * 4).double()
    f, a = ops.wavelet_upsample_taps(k.numpy())
    f, a = f.reshape(2, 2, 4, 4, 2, 2), a.reshape(4, 4, 4, 4)
    torch.manual_seed(0)
    Ci, h, w = 3, 5, 7
    x = torch.randn(1, 4 * Ci, h, w, dtype=torch.float64, requires_grad=True)
    y = so.wavelet_upsample(x, k)
    g = torch.randn_like(y)
    y.backward(g)
    xp = np.pad(x.detach().numpy()[0].reshape(4, Ci, h, w), ((0, 0), (0, 0), (1, 1), (1, 1)))
    gp = np.pad(g.numpy()[0].reshape(4, Ci, 2 * h, 2 * w), ((0, 0), (0, 0), (1, 2), (1, 2)))
    out, dx = np.zeros((4, Ci, 2 * h, 2 * w)), np.zeros((4, Ci, h, w))
    for m in range(h):
        for n in range(w):
            for pi in range(2):
                for pj in range(2):
                    win = xp[:, :, m + pi:m + pi + 2, n + pj:n + pj + 2]              # (bi, c, di, dj)
                    out[:, :, 2 * m + pi, 2 * n + pj] = np.einsum("obij,bcij->oc", f[pi, pj], win)
            dx[:, :, m, n] = np.einsum("ioab,ocab->ic", a, gp[:, :, 2 * m:2 * m + 4, 2 * n:2 * n + 4])
    assert np.abs(out.reshape(1, 4 * Ci, 2 * h, 2 * w) - y.detach().numpy()).max() < 1e-12
    assert np.abs(dx.reshape(1, 4 * Ci, h, w) - x.grad.numpy()).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_grouped_modweight_equals_per_layer(dtype, built_lib):
    """agr_modweight_group_* (all layers of a net in ceil(L/40) launches) against the per-layer op: identical operands
    and weight gradients (same row bodies), style gradients up to atomic-add order.  45 layers -> two launches."""
    from animatablegaussians_b200 import styleunet_ops as ops
    g = torch.Generator(device="cuda").manual_seed(11)
    shapes = [(64, 32, 3, True, False, True), (12, 64, 1, False, False, True), (32, 48, 3, True, False, True),
              (16, 3, 3, False, False, False), (8, 2048, 3, True, False, True), (128, 128, 3, False, False, False),
              (16, 1024, 3, True, False, True), (5, 7, 1, True, False, True)]
    shapes = (shapes * 6)[:45]
    entries, ref = [], []
    for Cout, Cin, k, demod, tr, styled in shapes:
        w = torch.randn(1, Cout, Cin, k, k, device="cuda", generator=g).requires_grad_(True)
        s = (1 + 0.3 * torch.randn(1, Cin, device="cuda", generator=g)).requires_grad_(True) if styled else None
        scale = 1 / (Cin * k * k) ** 0.5
        entries.append((w, s, scale, demod, tr))
        w2 = w.detach().clone().requires_grad_(True)
        s2 = s.detach().clone().requires_grad_(True) if styled else None
        ref.append((w2, s2, scale, demod, tr))
    with ops.step_arena():
        plan = ops.prepare_weights([e[:4] for e in entries], dtype)          # {id(weight): (operand, fp32 gradient handle)}
        outs, handles = zip(*[plan[id(e[0])] for e in entries])
        singles, handles2 = zip(*[ops.mod_weight(w, s, scale, demod, dtype) for w, s, scale, demod, tr in ref])
        ups = [torch.randn(o.shape, device="cuda", generator=g).to(dtype).float() for o in outs]
        torch.autograd.backward(handles, ups)
        torch.autograd.backward(handles2, ups)
    for (w, s, *_), (w2, s2, *_), o, o2 in zip(entries, ref, outs, singles):
        assert torch.equal(o, o2)
        assert torch.equal(w.grad, w2.grad)
        if s is not None:
            _cmp(s.grad, s2.grad, 1e-5)


@pytest.mark.gpu
def test_grouped_equal_linear_equals_per_layer(built_lib):
    """agr_equal_linear_group_* against the per-layer op on slices of one latent; 43 layers -> two launches; the style
    gradients of all layers land in one d_latent."""
    from animatablegaussians_b200 import styleunet_ops as ops
    g = torch.Generator(device="cuda").manual_seed(12)
    D, n_lat = 96, 7

    class Lin:
        def __init__(self, out_dim, bias, lr_mul):
            self.weight = torch.randn(out_dim, D, device="cuda", generator=g).requires_grad_(True)
            self.bias = torch.randn(out_dim, device="cuda", generator=g).requires_grad_(True) if bias else None
            self.scale, self.lr_mul = lr_mul / D ** 0.5, lr_mul

    layers = [(Lin(o, b, lr), i % n_lat) for i, (o, b, lr) in enumerate([(64, True, 1.0), (5, False, 0.5), (130, True, 1.0), (32, True, 0.01)] * 11)][:43]
    lat = torch.randn(1, n_lat, D, device="cuda", generator=g).requires_grad_(True)
    lat2 = lat.detach().clone().requires_grad_(True)
    with ops.step_arena():
        ys = ops.equal_linear_group(lat, layers)
        ups = [torch.randn(1, m.weight.shape[0], device="cuda", generator=g) for m, _ in layers]
        torch.autograd.backward(ys, ups)
        grads = [(m.weight.grad.clone(), None if m.bias is None else m.bias.grad.clone()) for m, _ in layers]
        for m, _ in layers:
            m.weight.grad = None
            if m.bias is not None:
                m.bias.grad = None
        ys2 = [ops.equal_linear(lat2[:, i], m.weight, m.bias, m.scale, m.lr_mul) for m, i in layers]
        torch.autograd.backward(ys2, ups)
    for (m, _), y, y2, (gw, gb) in zip(layers, ys, ys2, grads):
        assert torch.equal(y, y2) and torch.equal(gw, m.weight.grad)
        if gb is not None:
            assert torch.equal(gb, m.bias.grad)
    _cmp(lat.grad, lat2.grad, 1e-5)


def test_tail_state_exchange_bookkeeping():
    """CPU: the tensors a rank receives from the prefix owner (parallel.py) are exactly what forward_view_tail() reads — both
    decoders' (out, skip) and the skip feature of the tail level — and with_tail_state() puts them back where _decode looks."""
    from animatablegaussians_b200 import styleunet
    net = styleunet.DualStyleUNet(inp_size=512, inp_ch=3, out_ch=3, out_size=1024, style_dim=512, n_mlp=2)
    n_cond = 1 + len(net.from_rgbs)
    assert net._tail_cond_index() == [-n_cond]          # view_level 8 -> tail = decoder level 5 -> the finest encoder feature
    mk = lambda tag: torch.full((1,), float(tag))
    prefix = dict(latent=None, noise=None, plan=None, cond_list=[mk(10 + i) for i in range(n_cond)], s1=(mk(1), mk(2)), s2=(mk(3), mk(4)))
    state = net.tail_state(prefix)
    assert [float(t) for t in state] == [1, 2, 3, 4, 10]
    received = [t + 100 for t in state]
    empty = dict(latent=None, noise=None, plan=None, cond_list=[None] * n_cond, s1=None, s2=None)     # what tail_prefix() provides
    got = net.with_tail_state(empty, received)
    assert float(got["s1"][0]) == 101 and float(got["s1"][1]) == 102 and float(got["s2"][0]) == 103 and float(got["s2"][1]) == 104
    assert float(got["cond_list"][-n_cond]) == 110 and all(c is None for c in got["cond_list"][1:])
    # the level the tail starts at is the one _decode enters with start = view_level + 2
    assert (net.view_level + 2) // 2 == len(net.to_rgbs1) - 1
