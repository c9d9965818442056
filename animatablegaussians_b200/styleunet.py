"""DualStyleUNet for B200 — host-side mirror of the reference's network/styleunet/dual_styleunet.py
(`DualStyleUNet`, dual_styleunet.py:636-911) with the same constructor, forward signature and
state_dict keys/shapes, so the authors' checkpoints load unchanged.

The module tree only *describes* the network (parameters, buffers, wiring).  Arithmetic is delegated to
animatablegaussians_b200.styleunet_ops, which owns the CUDA kernels:
    modulated_conv2d / equal_conv2d   dense 3x3 / 1x1 contractions (implicit GEMM)
    bias_act                          + noise, + bias, leaky-ReLU(0.2) * sqrt(2)     (fused_act.py:100-132)
    upfirdn2d / haar_dwt / haar_iwt   FIR resampling                                  (upfirdn2d.py:105-183)
Differences from the reference that do not change results:
  * modulation, demodulation and the 1/sqrt(fan_in) scale are folded into ONE weight-preparation pass
    per layer (the reference's `fused` branch materialises the same weight with 5 elementwise launches,
    dual_styleunet.py:256-265);
  * noise injection + bias + activation are one epilogue instead of three passes (dual_styleunet.py:598-604);
  * the shared encoder runs once and the per-view colour decoders can reuse the view-independent prefix
    (`forward_prefix` / `forward_view_tail`; SURVEY.md §7 hard part (f)).
"""
import math

import torch
from torch import nn

from . import styleunet_ops as ops

_CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32, 2048: 32, 4096: 32}


def make_kernel(k):
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


class _Fir(nn.Module):
    """Holds a `kernel` buffer (state_dict key compatibility) and applies upfirdn2d."""

    def __init__(self, kernel, up=1, down=1, pad=(0, 0)):
        super().__init__()
        self.register_buffer("kernel", kernel)
        self.up, self.down, self.pad = up, down, pad

    def forward(self, x):
        return ops.upfirdn2d(x, self.kernel, up=self.up, down=self.down, pad=self.pad)


def Blur(blur_kernel, pad, upsample_factor=1):
    k = make_kernel(blur_kernel)
    if upsample_factor > 1:
        k = k * (upsample_factor ** 2)
    return _Fir(k, pad=pad)


def Upsample(blur_kernel, factor=2):
    k = make_kernel(blur_kernel) * (factor ** 2)
    p = k.shape[0] - factor
    return _Fir(k, up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def Downsample(blur_kernel, factor=2):
    k = make_kernel(blur_kernel)
    p = k.shape[0] - factor
    return _Fir(k, down=factor, pad=((p + 1) // 2, p // 2))


def _haar():
    lo = torch.ones(1, 2) / math.sqrt(2)
    hi = torch.tensor([[-1.0, 1.0]]) / math.sqrt(2)
    return lo.T * lo, hi.T * lo, lo.T * hi, hi.T * hi  # ll, lh, hl, hh


class HaarTransform(nn.Module):
    def __init__(self, in_channels=None):
        super().__init__()
        for n, k in zip(("ll", "lh", "hl", "hh"), _haar()):
            self.register_buffer(n, k)

    def forward(self, x):
        return ops.haar_dwt(x)


class InverseHaarTransform(nn.Module):
    def __init__(self, in_channels=None):
        super().__init__()
        ll, lh, hl, hh = _haar()
        for n, k in zip(("ll", "lh", "hl", "hh"), (ll, -lh, -hl, hh)):
            self.register_buffer(n, k)

    def forward(self, x):
        return ops.haar_iwt(x)


class PixelNorm(nn.Module):
    def forward(self, x):
        return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0.0, lr_mul=1.0, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, x):
        y = ops.equal_linear(x, self.weight, self.bias, self.scale, self.lr_mul)
        return ops.bias_act(y, None) if self.activation else y   # fused_leaky_relu(x W^T * scale, bias * lr_mul)


class EqualConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride, self.padding = stride, padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None

    def forward(self, x):
        return ops.bias_act(x, self.bias)


class ConvLayer(nn.Sequential):
    """[Blur ->] EqualConv2d -> FusedLeakyReLU, keys `0/1/2` like the reference (dual_styleunet.py:329-371).
    forward() runs conv + bias + activation as one op."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=(1, 3, 3, 1), bias=True,
                 activate=True):
        layers = []
        if downsample:
            p = (len(blur_kernel) - 2) + (kernel_size - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride, padding = 2, 0
        else:
            stride, padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel, bias=bias))
        super().__init__(*layers)
        self.has_blur, self.activate = downsample, activate

    def forward(self, x):
        i = 0
        if self.has_blur:
            x = self[0](x)
            i = 1
        conv = self[i]
        if self.activate:
            return ops.equal_conv2d(x, conv.weight, conv.scale, conv.stride, conv.padding, act_bias=self[i + 1].bias, activate=True)
        return ops.equal_conv2d(x, conv.weight, conv.scale, conv.stride, conv.padding, act_bias=conv.bias, activate=False)


class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.kernel_size, self.in_channel, self.out_channel = kernel_size, in_channel, out_channel
        self.upsample, self.downsample, self.demodulate = upsample, downsample, demodulate
        if upsample:
            p = (len(blur_kernel) - 2) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + 1, p // 2 + 1), upsample_factor=2)
        if downsample:
            p = (len(blur_kernel) - 2) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class StyledConv(nn.Module):
    """modulated conv -> + w_noise * noise -> + bias -> lrelu*sqrt2 (dual_styleunet.py:570-604), one fused op."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=(1, 3, 3, 1),
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, x, style, noise=None):
        c = self.conv
        s = lambda: c.modulation(style)   # only evaluated outside a weight plan
        if noise is None:  # randomize_noise=True path of the reference (dual_styleunet.py:309-313)
            h = x.shape[2] * (2 if c.upsample else 1)
            noise = x.new_empty(1, 1, h, h).normal_()
        return ops.modulated_conv2d(x, c.weight, s, c.scale, demodulate=c.demodulate, upsample=c.upsample,
                                    downsample=c.downsample, blur=getattr(c, "blur", None), padding=c.padding,
                                    noise=noise, noise_weight=self.noise.weight, act_bias=self.activate.bias,
                                    activate=True)


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, out_channel=12, upsample=True, blur_kernel=(1, 3, 3, 1), use_wt=True):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
            self.iwt = InverseHaarTransform(3)
            self.dwt = HaarTransform(3)
        self.out_channel = out_channel
        self.conv = ModulatedConv2d(in_channel, out_channel, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, out_channel, 1, 1))

    def forward(self, x, style, skip=None):
        c = self.conv
        out = ops.modulated_conv2d(x, c.weight, lambda: c.modulation(style), c.scale, demodulate=False, padding=0,
                                   act_bias=self.bias.view(-1), activate=False)
        if skip is not None:
            out = out + ops.wavelet_upsample(skip, self.upsample.kernel)  # dwt(upsample(iwt(skip)))
        return out


class FromRGB(nn.Module):
    def __init__(self, out_channel, in_channel, downsample=True, blur_kernel=(1, 3, 3, 1), use_wt=False):
        super().__init__()
        assert not use_wt, "DualStyleUNet only instantiates FromRGB(use_wt=False) (dual_styleunet.py:694)"
        self.downsample = Downsample(blur_kernel) if downsample else None
        self.conv = ConvLayer(in_channel, out_channel, 1)

    def forward(self, img, skip=None):
        if self.downsample is not None:
            img = self.downsample(img)
        out = self.conv(img)
        if skip is not None:
            out = out + skip
        return img, out


class ConvBlock(nn.Module):
    def __init__(self, in_channel, out_channel, blur_kernel=(1, 3, 3, 1), downsample=True):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=downsample)

    def forward(self, x):
        return self.conv2(self.conv1(x))


class DualStyleUNet(nn.Module):
    def __init__(self, inp_size, inp_ch, out_ch, out_size, style_dim, n_mlp, middle_size=8, c_dim=0,
                 channel_multiplier=2, blur_kernel=(1, 3, 3, 1), lr_mlp=0.01):
        super().__init__()
        assert channel_multiplier == 2 and c_dim == 0
        self.inp_size, self.style_dim = inp_size, style_dim
        self.middle_log_size = int(math.log(middle_size, 2))
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.channels = dict(_CHANNELS)
        self.log_size = int(math.log(out_size, 2)) - 1

        in_channel = self.channels[inp_size // 2]
        self.from_rgbs, self.cond_convs, self.comb_convs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.comb_convs.append(ConvLayer(in_channel * 2, in_channel, 3))
        self.conv_in = ConvLayer(inp_ch, in_channel, 3, downsample=True)
        for i in range(int(math.log(inp_size, 2)) - 2, self.middle_log_size - 1, -1):
            out_channel = self.channels[2 ** i]
            self.from_rgbs.append(FromRGB(in_channel, inp_ch, downsample=True, use_wt=False))
            self.cond_convs.append(ConvBlock(in_channel, out_channel, blur_kernel))
            self.comb_convs.append(ConvLayer(out_channel * 2 if i > self.middle_log_size else out_channel, out_channel, 3))
            in_channel = out_channel

        self.convs1, self.convs2 = nn.ModuleList(), nn.ModuleList()
        self.to_rgbs1, self.to_rgbs2 = nn.ModuleList(), nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[middle_size]
        self.num_layers = (self.log_size - self.middle_log_size) * 2
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 8) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(self.middle_log_size + 1, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            for convs, rgbs in ((self.convs1, self.to_rgbs1), (self.convs2, self.to_rgbs2)):
                convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
                convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
                rgbs.append(ToRGB(in_channel=out_channel, style_dim=style_dim, out_channel=out_ch * 4))
            in_channel = out_channel
        self.iwt = InverseHaarTransform(out_ch)
        self.n_latent = self.log_size * 2 - (self.middle_log_size * 2 - 1) + 1
        self.view_level = 8  # loop index after which the view feature is added (dual_styleunet.py:881,900)
        # the two decoders share the encoder output and nothing else: the second one runs on a side stream, forward AND
        # backward (autograd replays each node on its forward stream) — their batch-1 layers launch 2 - 128 CTAs each
        import os
        self.concurrent_decoders = os.environ.get("AGR_SERIAL_DECODERS", "0") != "1"
        self._side = None

    # ------------------------------------------------------------------ pieces
    def make_noise(self, device, zero_noise=False):
        f = torch.zeros if zero_noise else torch.randn
        return [f(1, 1, 2 ** i, 2 ** i, device=device) for i in range(self.middle_log_size + 1, self.log_size + 1) for _ in range(2)]

    def get_latent(self, x):
        return self.style(x)

    def _latent(self, styles, input_is_latent, truncation, truncation_latent, inject_index):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            return styles[0].unsqueeze(1).repeat(1, self.n_latent, 1) if styles[0].ndim < 3 else styles[0]
        if inject_index is None:
            import random
            inject_index = random.randint(1, self.n_latent - 1)
        a = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        b = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
        return torch.cat([a, b], 1)

    def encode(self, condition_img):
        """Shared encoder (dual_styleunet.py:853-862): returns cond_list, finest first."""
        cond_img = condition_img
        cond_out = self.conv_in(cond_img)
        cond_list = [cond_out]
        for from_rgb, cond_conv in zip(self.from_rgbs, self.cond_convs):
            cond_img, cond_out = from_rgb(cond_img, cond_out)
            cond_out = cond_conv(cond_out)
            cond_list.append(cond_out)
        return cond_list

    def _decode(self, convs, to_rgbs, cond_list, latent, noise, view_feature, start=0, state=None, stop=None):
        """One decoder (dual_styleunet.py:867-885). `start/state/stop` let callers split it at a level."""
        out, skip = state if state is not None else (None, None)
        n_levels = len(to_rgbs)
        for lvl in range(start // 2, n_levels):
            i = 2 * lvl
            if stop is not None and i >= stop:
                return out, skip
            if i == 0:
                out = self.comb_convs[-1](cond_list[-1])
            elif i < 2 * len(self.comb_convs):
                cond = cond_list[-1 - lvl]
                comb = self.comb_convs[-1 - lvl]
                if cond.shape[0] != out.shape[0]:   # view batch: the skip feature is shared -> split contraction
                    conv = comb[0]
                    out = ops.equal_conv2d_split(out, cond, conv.weight, conv.scale, act_bias=comb[1].bias, activate=True)
                else:
                    out = comb(torch.cat([out, cond], dim=1))
            out = convs[i](out, latent[:, i], noise=noise[i])
            out = convs[i + 1](out, latent[:, i + 1], noise=noise[i + 1])
            skip = to_rgbs[lvl](out, latent[:, i + 2], skip)
            if view_feature is not None and i == self.view_level:
                out = ops.add_view_feature(out, view_feature)
        if stop is not None:
            return out, skip
        return self.iwt(skip)

    def _both_decoders(self, run1, run2, ref):
        """(run1(), run2()) with run2 on this net's side stream when `ref` lives on a GPU."""
        if not (self.concurrent_decoders and ref.is_cuda):
            return run1(), run2()
        if self._side is None or self._side.device != ref.device:
            self._side = torch.cuda.Stream(ref.device)
        main = torch.cuda.current_stream(ref.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            r2 = run2()
        r1 = run1()
        main.wait_stream(self._side)
        for t in (r2 if isinstance(r2, (tuple, list)) else (r2,)):
            if isinstance(t, torch.Tensor):
                t.record_stream(main)
        return r1, r2

    def weight_plan(self, latent, tail_only=False):
        """Conv-ready operands of EVERY layer for this latent, prepared by the grouped kernels (a few launches instead
        of one per layer): encoder / combiner equalised convs (scale only) and the two decoders' modulated convs
        (style modulation of latent[:, i] as _decode indexes it).  Only for the single-style case the avatar runs.
        `tail_only`: just the layers forward_view_tail() runs (a rank that receives the prefix state from its owner)."""
        if latent.shape[0] != 1 or not latent.is_cuda:
            return None
        entries = []
        first = (self.view_level + 2) // 2 if tail_only else 0      # first decoder level of the plan

        def plain(layer):   # ConvLayer: [Blur,] EqualConv2d [, FusedLeakyReLU]
            conv = layer[1] if layer.has_blur else layer[0]
            entries.append((conv.weight, None, conv.scale, False))

        if not tail_only:
            plain(self.conv_in)
            for from_rgb, cond_conv in zip(self.from_rgbs, self.cond_convs):
                plain(from_rgb.conv); plain(cond_conv.conv1); plain(cond_conv.conv2)
            for comb in self.comb_convs:
                plain(comb)
        else:
            for lvl in range(first, len(self.to_rgbs1)):
                if 0 < 2 * lvl < 2 * len(self.comb_convs):
                    plain(self.comb_convs[-1 - lvl])
        mods = []   # (ModulatedConv2d, latent index) in _decode's order
        for convs, rgbs in ((self.convs1, self.to_rgbs1), (self.convs2, self.to_rgbs2)):
            mods += [(sc.conv, i) for i, sc in enumerate(convs) if i >= 2 * first]
            mods += [(rgb.conv, 2 * lvl + 2) for lvl, rgb in enumerate(rgbs) if lvl >= first]
        if latent.dtype == torch.float32 and latent.dim() == 3:
            styles = ops.equal_linear_group(latent, [(c.modulation, i) for c, i in mods])
        else:
            styles = [c.modulation(latent[:, i]) for c, i in mods]
        for (c, _), s in zip(mods, styles):
            entries.append((c.weight, s, c.scale, c.demodulate))
        return ops.prepare_weights(entries, ops.compute_dtype())

    # ------------------------------------------------------------------ reference-shaped forward
    def forward(self, styles, condition_img, cond=None, return_latents=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True, view_feature1=None,
                view_feature2=None):
        assert cond is None
        latent = self._latent(styles, input_is_latent, truncation, truncation_latent, inject_index)
        if noise is None:
            noise = [None] * self.num_layers if randomize_noise else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        x = ops.to_compute(condition_img)
        with ops.weight_plan(self.weight_plan(latent)):
            cond_list = self.encode(x)
            image1 = self._decode(self.convs1, self.to_rgbs1, cond_list, latent, noise, view_feature1)
            image2 = self._decode(self.convs2, self.to_rgbs2, cond_list, latent, noise, view_feature2)
        images = ops.from_compute(torch.cat([image1, image2], 1))
        return (images, latent) if return_latents else (images, None)

    def forward_maps(self, styles, condition_img, view_feature1=None, view_feature2=None):
        """forward() with the fixed noise buffers, returning the two decoder outputs (front, back) separately in the
        compute dtype / NHWC — what AvatarNet's fused gather consumes (no channel cat, no fp32 NCHW copy)."""
        latent = self._latent(styles, False, 1, None, None)
        noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        with ops.weight_plan(self.weight_plan(latent)):
            cond_list = self.encode(ops.to_compute(condition_img))
            return self._both_decoders(lambda: self._decode(self.convs1, self.to_rgbs1, cond_list, latent, noise, view_feature1),
                                       lambda: self._decode(self.convs2, self.to_rgbs2, cond_list, latent, noise, view_feature2), cond_list[0])

    # ------------------------------------------------------------------ view-batch split (exact)
    def forward_prefix(self, styles, condition_img):
        """Everything that does not depend on the view feature: encoder + both decoders up to and including
        level `view_level` (the addition of the view feature happens at the START of the tail)."""
        latent = self._latent(styles, False, 1, None, None)
        noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        plan = self.weight_plan(latent)
        with ops.weight_plan(plan):
            cond_list = self.encode(ops.to_compute(condition_img))
            stop = self.view_level + 2
            s1, s2 = self._both_decoders(lambda: self._decode(self.convs1, self.to_rgbs1, cond_list, latent, noise, None, stop=stop),
                                         lambda: self._decode(self.convs2, self.to_rgbs2, cond_list, latent, noise, None, stop=stop), cond_list[0])
        return dict(latent=latent, noise=noise, cond_list=cond_list, s1=s1, s2=s2, plan=plan)

    # ---- prefix state exchange (animatablegaussians_b200/parallel.py: the prefix runs on one owner rank) -----------------------
    def _tail_cond_index(self):
        """(Negative) indices into cond_list that the tail reads: `cond_list[-1 - lvl]` of each tail level (see _decode)."""
        first = (self.view_level + 2) // 2
        return [-1 - lvl for lvl in range(first, len(self.to_rgbs1)) if 0 < 2 * lvl < 2 * len(self.comb_convs)]

    def tail_state(self, prefix):
        """The tensors forward_view_tail() reads from a prefix: both decoders' (out, skip) + the skip features of the tail levels."""
        return [prefix["s1"][0], prefix["s1"][1], prefix["s2"][0], prefix["s2"][1]] + [prefix["cond_list"][k] for k in self._tail_cond_index()]

    def with_tail_state(self, prefix, tensors):
        """`prefix` (from forward_prefix or tail_prefix) with the exchanged tensors in place of its own."""
        cond = list(prefix["cond_list"])
        for k, t in zip(self._tail_cond_index(), tensors[4:]):
            cond[k] = t
        return dict(prefix, s1=(tensors[0], tensors[1]), s2=(tensors[2], tensors[3]), cond_list=cond)

    def tail_prefix(self, styles):
        """What a rank that does NOT run the prefix needs before forward_view_tail(): latent, noise buffers, the tail layers'
        conv operands; the state tensors come from the owner (with_tail_state)."""
        latent = self._latent(styles, False, 1, None, None)
        noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        return dict(latent=latent, noise=noise, cond_list=[None] * (1 + len(self.from_rgbs)), s1=None, s2=None,
                    plan=self.weight_plan(latent, tail_only=True))

    def forward_view_tail(self, prefix, view_feature1, view_feature2, as_pair=False):
        """View-dependent remainder for a BATCH of V views (view features (V,128,h,w)): add the (bilinearly resized)
        view feature to the shared prefix state, run the last decoder level(s) once with batch V — every layer's
        modulated weight is prepared once and shared by the V views."""
        outs = []
        V = view_feature1.shape[0] if view_feature1 is not None else 1
        for convs, rgbs, st, vf in ((self.convs1, self.to_rgbs1, prefix["s1"], view_feature1),
                                    (self.convs2, self.to_rgbs2, prefix["s2"], view_feature2)):
            out, skip = st
            if vf is not None and self.view_level < 2 * len(rgbs):  # smaller nets never reach the view level
                out = ops.add_view_feature(out, vf)                  # (1|V) + resize(V) -> V, one pass
            elif V > 1:
                out = ops.expand_batch(out, V)
            if V > 1 and skip is not None:
                skip = ops.expand_batch(skip, V)
            with ops.weight_plan(prefix.get("plan")):
                outs.append(self._decode(convs, rgbs, prefix["cond_list"], prefix["latent"], prefix["noise"], None,
                                         start=self.view_level + 2, state=(out, skip)))
        if as_pair:
            return outs[0], outs[1]
        return ops.from_compute(torch.cat(outs, 1))
