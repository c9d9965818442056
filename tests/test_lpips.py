"""Perceptual loss (second part): animatablegaussians_b200.lpips.LPIPS against the reference's own
`network.lpips.LPIPS` (network/lpips/lpips.py:23-124) imported unmodified from the installed reference copy and run on
torch / cuDNN — same state_dict (random VGG-16 trunk: the ImageNet weights are a download; the reference's shipped `lin`
weights), same inputs: value, per-layer values and the gradient w.r.t. the rendered image."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "baseline", "_ref", "AnimatableGaussians")


def _reference_lpips():
    if not os.path.exists(os.path.join(REF_PY, "network", "lpips", "lpips.py")):
        pytest.skip("baseline/_ref not installed (python -m oracle.install_ref needs /root/reference)")
    pytest.importorskip("torchvision")
    if REF_PY not in sys.path:
        sys.path.insert(0, REF_PY)
    from network.lpips import LPIPS as RefLPIPS
    torch.manual_seed(0)
    return RefLPIPS(net="vgg", pnet_rand=True, verbose=False)


def test_state_dict_keys_match_reference():
    """CPU: same keys / shapes as the reference module, and the reference's lin weights load."""
    from animatablegaussians_b200 import lpips
    ref = _reference_lpips()
    ours = lpips.LPIPS(net="vgg", pnet_rand=True)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    ours.load_state_dict(a, strict=True)
    assert all(torch.equal(a[k], ours.state_dict()[k]) for k in a)
    assert not any(p.requires_grad for p in ours.net.parameters())
    with pytest.raises(ValueError):
        lpips.LPIPS(net="alex", pnet_rand=True)
    with pytest.raises(RuntimeError, match="GPU only"):
        ours(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,H,W,N", [(torch.float32, 64, 64, 1), (torch.float32, 96, 80, 2), (torch.bfloat16, 128, 128, 1)])
def test_lpips_matches_reference(dtype, H, W, N, built_lib):
    from animatablegaussians_b200 import lpips, styleunet_ops as ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = _reference_lpips().cuda()
    ours = lpips.LPIPS(net="vgg", pnet_rand=True).cuda()
    ours.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand(N, 3, H, W, device="cuda", generator=g)
    gt = (img + 0.25 * torch.randn(N, 3, H, W, device="cuda", generator=g)).clamp(0, 1)
    a = img.clone().requires_grad_(True)
    val_ref, per_ref = ref.forward(a, gt, retPerLayer=True, normalize=True)
    val_ref.mean().backward()
    ops.set_compute_dtype(dtype)
    try:
        b = img.clone().requires_grad_(True)
        val, per = ours.forward(b, gt, retPerLayer=True, normalize=True)
        val.mean().backward()
    finally:
        ops.set_compute_dtype(torch.float32)
    tol_v, tol_g = (1e-4, 6e-3) if dtype == torch.float32 else (3e-2, 2.5e-1)   # gradients cross the ReLU / max-pool kinks of 13 layers
    assert val.shape == val_ref.shape == (N, 1, 1, 1)
    assert float((val - val_ref).abs().max() / val_ref.abs().max()) < tol_v
    # the reference accumulates the total IN PLACE into its first per-layer tensor (lpips.py:105-107: `val = res[0]; val += res[l]`),
    # so its res[0] is the total; layer 0 on its own is the total minus the other four
    per_ref = [val_ref - sum(per_ref[1:])] + list(per_ref[1:])
    for x, y in zip(per, per_ref):
        assert float((x - y).abs().max() / val_ref.abs().max()) < 3 * tol_v
    rel = float((b.grad - a.grad).norm() / a.grad.norm())
    assert rel < tol_g, rel


def test_head_formulas_match_reference_autograd():
    """CPU: the closed-form forward / backward of the LPIPS head as the CUDA kernel computes it (oracle/lpips_oracle.py) against
    autograd of the reference's own `normalize_tensor` + squared difference + lin weights + spatial mean, in float64."""
    from oracle import lpips_oracle as lo
    ref = _reference_lpips()
    import network.lpips as ref_pkg
    g = torch.Generator().manual_seed(9)
    for C, H, W in ((64, 9, 7), (512, 4, 4)):
        f = torch.relu(torch.randn(2, C, H, W, generator=g, dtype=torch.float64)).requires_grad_(True)
        w = torch.rand(C, generator=g, dtype=torch.float64)
        d = (ref_pkg.normalize_tensor(f[0:1]) - ref_pkg.normalize_tensor(f[1:2])) ** 2
        val = (d * w[None, :, None, None]).sum(1, keepdim=True).mean([2, 3]).sum()
        val.backward()
        assert abs(float(lo.layer_forward(f.detach(), w)) - float(val)) < 1e-12
        got = lo.layer_backward(f.detach(), w, 1.0)
        assert torch.allclose(got, f.grad, rtol=1e-9, atol=1e-14)
