"""Generates tests/golden/styleunet_*.npz by importing the UNMODIFIED reference DualStyleUNet
(/root/reference/network/styleunet/dual_styleunet.py) in THIS container (CPU, fp32; the reference's `fused` and
`upfirdn2d` CUDA extensions are absent on CPU, where fused_act.py:118-129 / upfirdn2d.py:177-227 take their
pure-PyTorch branches — stub modules satisfy the import).   python tests/golden/make_styleunet_golden.py
Weights are not stored: both sides fill every state_dict entry from a name-seeded generator (fill_state)."""
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("AGR_REFERENCE_ROOT", "/root/reference")

CASES = {
    # name: (ctor kwargs, use view features, output stride for storage)
    "small": (dict(inp_size=64, inp_ch=3, out_ch=3, out_size=128, style_dim=64, n_mlp=2), False, 1),
    "small8": (dict(inp_size=32, inp_ch=3, out_ch=8, out_size=64, style_dim=32, n_mlp=2), False, 1),
    "full_view": (dict(inp_size=512, inp_ch=3, out_ch=3, out_size=1024, style_dim=512, n_mlp=2), True, 8),
}
GRAD_KEYS = ["conv_in.2.bias", "convs1.0.conv.modulation.bias", "convs1.1.noise.weight", "to_rgbs2.0.bias",
             "style.1.bias", "comb_convs.0.1.bias", "convs2.2.activate.bias"]


def fill_state(module, gain=1.0):
    """Deterministic, name-seeded values for every parameter and every noise buffer (FIR/Haar buffers keep
    their constructed values).  Shared by the generator and tests/test_styleunet.py."""
    sd = module.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith("kernel") or k.split(".")[-1] in ("ll", "lh", "hl", "hh"):
            new[k] = v
            continue
        seed = int(hashlib.sha256(k.encode()).hexdigest()[:8], 16)
        g = torch.Generator().manual_seed(seed)
        t = torch.randn(v.shape, generator=g, dtype=torch.float32)
        if k.endswith("modulation.bias"):
            t = 1.0 + 0.1 * t
        elif k.endswith("bias"):
            t = 0.1 * t
        elif k.endswith("noise.weight"):
            t = 0.3 * t
        elif k.startswith("style."):
            t = t * 100.0 if k.endswith("weight") else t   # lr_mul = 0.01 parametrisation
        new[k] = t * gain if k.endswith("weight") and not k.startswith("style.") and "noise" not in k else t
    module.load_state_dict(new, strict=True)


def inputs(cfg, use_view):
    g = torch.Generator().manual_seed(1234)
    cond = torch.randn(1, cfg["inp_ch"], cfg["inp_size"], cfg["inp_size"], generator=g) * 0.5
    style = torch.ones(1, cfg["style_dim"]) / np.sqrt(cfg["style_dim"])
    vf1 = torch.randn(1, 128, 128, 128, generator=g) * 0.2 if use_view else None
    vf2 = torch.randn(1, 128, 128, 128, generator=g) * 0.2 if use_view else None
    up = torch.randn(1, 2 * cfg["out_ch"], cfg["out_size"], cfg["out_size"], generator=g)
    return cond, style, vf1, vf2, up


def import_reference():
    for name in ("fused", "upfirdn2d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    from network.styleunet.dual_styleunet import DualStyleUNet
    return DualStyleUNet


if __name__ == "__main__":
    DualStyleUNet = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    torch.set_num_threads(8)
    for name, (cfg, use_view, stride) in CASES.items():
        torch.manual_seed(0)
        net = DualStyleUNet(**cfg)
        fill_state(net)
        cond, style, vf1, vf2, up = inputs(cfg, use_view)
        cond.requires_grad_(True)
        out, _ = net([style], cond, randomize_noise=False, view_feature1=vf1, view_feature2=vf2)
        (out * up).sum().backward()
        named = dict(net.named_parameters())
        d = dict(out=out.detach()[..., ::stride, ::stride].numpy(), out_mean=np.float64(out.detach().double().mean()),
                 out_std=np.float64(out.detach().double().std()), grad_cond=cond.grad.numpy(), stride=np.int64(stride))
        for k in GRAD_KEYS:
            if k in named and named[k].grad is not None:
                d["grad:" + k] = named[k].grad.numpy()
        np.savez_compressed(os.path.join(out_dir, "styleunet_%s.npz" % name), **d)
        print("wrote", name, out.shape, float(out.abs().max()))
