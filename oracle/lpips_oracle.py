"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the LPIPS head of ONE tapped layer exactly as csrc/lpips.cu computes it
(include/agr_lpips.h): closed-form forward and backward per pixel.  Reference semantics: network/lpips/lpips.py:88-103
(`normalize_tensor` of both feature stacks, squared difference, 1x1 `lin` weights, spatial mean) with
network/lpips/__init__.py:40-42 (`sqrt(sum x^2 + eps)`, `x / (norm + eps)`, eps = 1e-10).  Pinned on CPU against autograd of that
reference formula (tests/test_lpips.py::test_head_formulas_match_reference_autograd); the CUDA kernels are checked against the
reference module itself on the GPU."""
import torch

EPS = 1e-10


def layer_forward(f, w):
    """f (2, C, H, W): features of image 0 / image 1; w (C,) -> scalar."""
    f0, f1 = f[0], f[1]
    n0 = torch.sqrt((f0 * f0).sum(0) + EPS)
    n1 = torch.sqrt((f1 * f1).sum(0) + EPS)
    e = f0 / (n0 + EPS) - f1 / (n1 + EPS)
    return (w[:, None, None] * e * e).sum(0).mean()


def layer_backward(f, w, g=1.0):
    """d(layer_forward)/df * g, with the kernel's closed form: u = f i, i = 1/(n + eps), du_c/df_k = i delta_ck - f_c f_k i^2 / n."""
    f0, f1 = f[0], f[1]
    pixels = f0.shape[1] * f0.shape[2]
    n0 = torch.sqrt((f0 * f0).sum(0) + EPS)
    n1 = torch.sqrt((f1 * f1).sum(0) + EPS)
    i0, i1 = 1.0 / (n0 + EPS), 1.0 / (n1 + EPS)
    e = (2.0 * g / pixels) * w[:, None, None] * (f0 * i0 - f1 * i1)
    t0 = (e * f0).sum(0) * i0 * i0 / n0
    t1 = (e * f1).sum(0) * i1 * i1 / n1
    return torch.stack([e * i0 - f0 * t0, -e * i1 + f1 * t1], 0)
