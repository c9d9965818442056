"""Operators of the B200 DualStyleUNet (see styleunet.py).

Every operator has ONE implementation.  Operators whose hand-written sm_100a kernel exists call it through
the C ABI (include/agr_styleunet.h); operators listed in LIBRARY_OPS below still call a vendor library
(cuDNN through torch) and are reported as such by bench.py — they are the next kernels to replace, not a
fallback: there is no runtime switch between a kernel and a library path.

Semantics restated from the reference:
  bias_act            network/styleunet/fused_act.py:100-132, fused_bias_act_kernel.cu:18-65 (act=3)
  upfirdn2d           network/styleunet/upfirdn2d.py:105-227, upfirdn2d_kernel.cu:107-207
  haar_dwt/haar_iwt   dual_styleunet.py:374-425
  modulated_conv2d    dual_styleunet.py:225-300 (fused branch) + :303-313 (noise) + fused lrelu
  equal_conv2d        dual_styleunet.py:93-122 + ConvLayer :329-371
"""
import math

import torch
import torch.nn.functional as F

# operators that are still vendor-library calls (cuDNN via torch) in this round
LIBRARY_OPS = ("conv2d(cuDNN)", "conv_transpose2d(cuDNN)")

_COMPUTE_DTYPE = torch.float32
_SQRT2 = math.sqrt(2.0)


def set_compute_dtype(dtype):
    """torch.float32 (parity tests) or torch.bfloat16 (BASELINE config 4: 'bf16 StyleUNet')."""
    global _COMPUTE_DTYPE
    assert dtype in (torch.float32, torch.bfloat16)
    _COMPUTE_DTYPE = dtype


def compute_dtype():
    return _COMPUTE_DTYPE


def to_compute(x):
    x = x.to(_COMPUTE_DTYPE)
    if x.is_cuda and x.ndim == 4:
        x = x.contiguous(memory_format=torch.channels_last)
    return x


def from_compute(x):
    return x.float().contiguous()


def _w(t):
    return t.to(_COMPUTE_DTYPE)


# ------------------------------------------------------------------------------------------ elementwise
def bias_act(x, bias=None, noise=None, noise_weight=None, activate=True):
    """y = lrelu(x + w_noise * noise + bias[c], 0.2) * sqrt(2)   (activate=False: no lrelu / gain)."""
    if noise is not None:
        x = x + _w(noise_weight) * _w(noise)
    if bias is not None:
        x = x + _w(bias).view(1, -1, *([1] * (x.ndim - 2)))
    if activate:
        x = F.leaky_relu(x, 0.2) * _SQRT2
    return x


def bilinear_resize(x, size):
    """F.interpolate(mode='bilinear') of the view feature (dual_styleunet.py:882,901)."""
    return F.interpolate(x.float(), size, mode="bilinear").to(_COMPUTE_DTYPE)


# ------------------------------------------------------------------------------------------ FIR resampling
def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """Per-channel: zero-insert upsample by `up`, pad, convolve with `kernel` (true convolution), decimate."""
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    kh, kw = kernel.shape
    y = x.reshape(B * C, 1, H, W)
    if up > 1:
        z = y.new_zeros(B * C, 1, H, up, W, up)
        z[:, :, :, 0, :, 0] = y
        y = z.reshape(B * C, 1, H * up, W * up)
    y = F.pad(y, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    y = y[:, :, max(-py0, 0): y.shape[2] - max(-py1, 0), max(-px0, 0): y.shape[3] - max(-px1, 0)]
    w = torch.flip(kernel, [0, 1]).view(1, 1, kh, kw).to(y.dtype)
    y = F.conv2d(y, w, stride=down)
    return y.reshape(B, C, y.shape[2], y.shape[3])


def _haar_kernels(dev, inverse):
    s = 0.5
    ll = torch.tensor([[s, s], [s, s]], device=dev)
    lh = torch.tensor([[-s, -s], [s, s]], device=dev)
    hl = torch.tensor([[-s, s], [-s, s]], device=dev)
    hh = torch.tensor([[s, -s], [-s, s]], device=dev)
    return (ll, -lh, -hl, hh) if inverse else (ll, lh, hl, hh)


def haar_dwt(x):
    """HaarTransform.forward (dual_styleunet.py:398-404): C -> 4C at half resolution, order ll|lh|hl|hh."""
    ks = _haar_kernels(x.device, False)
    return torch.cat([upfirdn2d(x, k, down=2) for k in ks], 1)


def haar_iwt(x):
    """InverseHaarTransform.forward (dual_styleunet.py:418-425): 4C -> C at double resolution."""
    ks = _haar_kernels(x.device, True)
    parts = x.chunk(4, 1)
    out = None
    for p, k in zip(parts, ks):
        y = upfirdn2d(p, k, up=2, pad=(1, 0, 1, 0))
        out = y if out is None else out + y
    return out


def wavelet_upsample(skip, up_kernel):
    """ToRGB skip path (dual_styleunet.py:624-631): dwt(upsample(iwt(skip)))."""
    k = up_kernel
    p = k.shape[0] - 2
    y = haar_iwt(skip)
    y = upfirdn2d(y, k, up=2, pad=((p + 1) // 2 + 1, p // 2))
    return haar_dwt(y)


# ------------------------------------------------------------------------------------------ dense contractions
def equal_conv2d(x, weight, scale, stride, padding, act_bias=None, activate=True):
    out = F.conv2d(x, _w(weight * scale), None, stride=stride, padding=padding)
    return bias_act(out, act_bias, activate=activate)


def prepare_modulated_weight(weight, s, scale, demodulate):
    """(1,Cout,Cin,k,k), (B,Cin) -> (B,Cout,Cin,k,k): scale * w * s, optionally demodulated
    (dual_styleunet.py:256-261)."""
    B = s.shape[0]
    w = scale * weight * s.view(B, 1, -1, 1, 1)
    if demodulate:
        w = w * torch.rsqrt(w.pow(2).sum([2, 3, 4]) + 1e-8).view(B, -1, 1, 1, 1)
    return w


def modulated_conv2d(x, weight, s, scale, demodulate=True, upsample=False, downsample=False, blur=None, padding=1,
                     noise=None, noise_weight=None, act_bias=None, activate=True):
    B, Cin, H, W = x.shape
    Cout, k = weight.shape[1], weight.shape[-1]
    w = _w(prepare_modulated_weight(weight, s.float(), scale, demodulate))
    if upsample:
        wt = w.transpose(1, 2).reshape(B * Cin, Cout, k, k)
        out = F.conv_transpose2d(x.reshape(1, B * Cin, H, W), wt, padding=0, stride=2, groups=B)
        out = blur(out.reshape(B, Cout, out.shape[2], out.shape[3]))
    elif downsample:
        x = blur(x)
        out = F.conv2d(x.reshape(1, B * Cin, x.shape[2], x.shape[3]), w.reshape(B * Cout, Cin, k, k), padding=0, stride=2, groups=B)
        out = out.reshape(B, Cout, out.shape[2], out.shape[3])
    else:
        out = F.conv2d(x.reshape(1, B * Cin, H, W), w.reshape(B * Cout, Cin, k, k), padding=padding, groups=B)
        out = out.reshape(B, Cout, out.shape[2], out.shape[3])
    return bias_act(out, act_bias, noise=noise, noise_weight=noise_weight, activate=activate)
