"""LPIPS perceptual loss on a VGG-16 trunk for B200 — host-side mirror of the reference's `network.lpips.LPIPS`
(network/lpips/lpips.py:23-124, trunk network/lpips/pretrained_networks.py:96-134) as the trainer uses it
(main_avatar.py:117-124,342: `LPIPS(net='vgg')`, `forward(img[None, [2,1,0]], gt[None, [2,1,0]], normalize=True).mean()`).

Same constructor arguments, same state_dict keys (`net.slice{1..5}.{idx}.weight|bias` with torchvision's vgg16.features
indices, `lin{0..4}.model.1.weight`, `scaling_layer.shift|scale`), so the reference's `weights/v0.1/vgg.pth` and a
torchvision VGG-16 state load unchanged.  The arithmetic runs on this package's kernels:
    13 x (3x3 conv + bias + ReLU)   include/agr_conv.h, epilogue activate = 3 (tcgen05 in bf16, CUDA cores in fp32 / for Cin = 3)
    4 x max-pool 2x2                include/agr_lpips.h agr_maxpool2x2_*
    5 x LPIPS head                  include/agr_lpips.h agr_lpips_layer_* (normalise, difference^2, lin weights, spatial mean)
Both images go through the trunk as ONE batch of two.  The trunk is frozen (`pnet_tune=False`, as the trainer builds it):
no weight gradients are computed.  `spatial=True` (per-pixel maps) is not part of the trainer's path and is rejected."""
import ctypes as C
import os

import torch
from torch import nn

from . import _lib, stats, styleunet_ops as ops

_p = C.c_void_p
_lib.register_symbols({
    "agr_maxpool2x2_forward": (C.c_int, [C.c_int32, _p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p]),
    "agr_maxpool2x2_backward": (C.c_int, [C.c_int32, _p, _p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _p]),
    "agr_lpips_layer_forward": (C.c_int, [C.c_int32, _p, _p, C.c_int64, C.c_int32, C.c_float, _p, _p]),
    "agr_lpips_layer_backward": (C.c_int, [C.c_int32, _p, _p, C.c_int64, C.c_int32, C.c_float, _p, _p, _p]),
})

_CL = torch.channels_last
# torchvision vgg16.features: conv indices per LPIPS slice (the ReLU / MaxPool modules between them hold no parameters)
_VGG_SLICES = (((0, 3, 64), (2, 64, 64)), ((5, 64, 128), (7, 128, 128)), ((10, 128, 256), (12, 256, 256), (14, 256, 256)),
               ((17, 256, 512), (19, 512, 512), (21, 512, 512)), ((24, 512, 512), (26, 512, 512), (28, 512, 512)))
_CHNS = (64, 128, 256, 512, 512)


def _check(st, what):
    if st != _lib.AGR_OK:
        raise RuntimeError("%s failed: %d" % (what, st))


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x = ops._nhwc(x)
        N, Cc, H, W = x.shape
        y = torch.empty((N, Cc, H // 2, W // 2), dtype=x.dtype, device=x.device, memory_format=_CL)
        with torch.cuda.device(x.device), stats.stage("lpips", launches=1):
            _check(lib.agr_maxpool2x2_forward(ops._code(x), ops._ptr(x), ops._ptr(y), N, H, W, Cc, ops._stream(x)), "agr_maxpool2x2_forward")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        g = ops._nhwc(g)
        N, Cc, H, W = x.shape
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device), stats.stage("lpips", launches=1):
            _check(lib.agr_maxpool2x2_backward(ops._code(x), ops._ptr(x), ops._ptr(g), ops._ptr(dx), N, H, W, Cc, ops._stream(x)),
                   "agr_maxpool2x2_backward")
        return dx


class _LpipsLayer(torch.autograd.Function):
    """feats (2, C, H, W) NHWC = features of image 0 and image 1; w (C,) fp32 -> scalar (1,) fp32."""

    @staticmethod
    def forward(ctx, feats, w):
        lib = _lib.load()
        f = ops._nhwc(feats)
        Cc, pixels = f.shape[1], f.shape[2] * f.shape[3]
        wc = w.detach().float().contiguous().view(-1)
        out = torch.zeros(1, dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device), stats.stage("lpips", launches=1):
            _check(lib.agr_lpips_layer_forward(ops._code(f), ops._ptr(f), ops._ptr(wc), pixels, Cc, 1.0 / pixels, ops._ptr(out), ops._stream(f)),
                   "agr_lpips_layer_forward")
        ctx.save_for_backward(f, wc)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        f, wc = ctx.saved_tensors
        Cc, pixels = f.shape[1], f.shape[2] * f.shape[3]
        gc = g.detach().float().contiguous().view(-1)
        df = torch.empty_like(f)
        with torch.cuda.device(f.device), stats.stage("lpips", launches=1):
            _check(lib.agr_lpips_layer_backward(ops._code(f), ops._ptr(f), ops._ptr(wc), pixels, Cc, 1.0 / pixels, ops._ptr(gc), ops._ptr(df),
                                                ops._stream(f)), "agr_lpips_layer_backward")
        return df, None


class _Slice(nn.Module):
    """Convolutions of one trunk slice under torchvision's `features` indices (state_dict keys `<idx>.weight|bias`)."""

    def __init__(self, spec):
        super().__init__()
        self.idx = [i for i, _, _ in spec]
        for i, cin, cout in spec:
            self.add_module(str(i), nn.Conv2d(cin, cout, 3, padding=1))


class _VGG16(nn.Module):
    def __init__(self, requires_grad=False):
        super().__init__()
        for k, spec in enumerate(_VGG_SLICES):
            setattr(self, "slice%d" % (k + 1), _Slice(spec))
        self.N_slices = 5
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False
        self._cache = {}

    def _operand(self, conv, dtype):
        """conv-ready KRSC operand + fp32 bias of a frozen conv, cached per dtype (re-made if the parameter was reloaded)."""
        key = (id(conv), dtype)
        hit = self._cache.get(key)
        ver = (conv.weight._version, conv.weight.data_ptr(), conv.bias._version)
        if hit is None or hit[0] != ver:
            hit = (ver, conv.weight.detach().to(dtype).contiguous(memory_format=_CL), conv.bias.detach().float().contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def forward(self, x):
        outs = []
        for k in range(5):
            sl = getattr(self, "slice%d" % (k + 1))
            if k > 0:
                x = _MaxPool.apply(x)          # features[4, 9, 16, 23] open slices 2..5
            for i in sl.idx:
                conv = getattr(sl, str(i))
                if conv.weight.requires_grad:
                    raise RuntimeError("LPIPS: the VGG trunk is frozen on this path (pnet_tune=False, as main_avatar.py builds it)")
                w, b = self._operand(conv, x.dtype)
                x = ops.conv2d(x, w, None, bias=b, activate=3, k=3, stride=1, pad=1)   # conv + bias + ReLU in one kernel
            outs.append(x)
        return outs


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """Holds the learned 1x1 `lin` weights under the reference's key (`model.1.weight` behind a Dropout, `model.0.weight` without)."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)

    @property
    def weight(self):
        return self.model[-1].weight


def default_lin_weights():
    """The reference's `weights/v0.1/vgg.pth` where a copy of the reference is installed (it is data, not shipped here)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base in (os.path.join(root, "baseline", "_ref", "AnimatableGaussians"), "/root/reference"):
        p = os.path.join(base, "network", "lpips", "weights", "v0.1", "vgg.pth")
        if os.path.exists(p):
            return p
    return None


class LPIPS(nn.Module):
    def __init__(self, pretrained=True, net="vgg", version="0.1", lpips=True, spatial=False, pnet_rand=False, pnet_tune=False,
                 use_dropout=True, model_path=None, eval_mode=True, verbose=False):
        super().__init__()
        if net not in ("vgg", "vgg16"):
            raise ValueError("LPIPS: only the VGG-16 trunk is on this path (main_avatar.py:342 builds LPIPS(net='vgg'))")
        if spatial:
            raise ValueError("LPIPS: spatial=True (per-pixel maps) is not on the trainer's path")
        if pnet_tune:
            raise ValueError("LPIPS: the trunk is frozen on this path (pnet_tune=False)")
        self.pnet_type, self.pnet_tune, self.pnet_rand = net, pnet_tune, pnet_rand
        self.spatial, self.lpips, self.version = spatial, lpips, version
        self.scaling_layer = ScalingLayer()
        self.chns = list(_CHNS)
        self.L = len(self.chns)
        self.net = _VGG16(requires_grad=False)
        if not pnet_rand:
            self._load_torchvision_trunk()
        if lpips:
            for k, c in enumerate(self.chns):
                setattr(self, "lin%d" % k, NetLinLayer(c, use_dropout=use_dropout))
            self.lins = nn.ModuleList([getattr(self, "lin%d" % k) for k in range(self.L)])
            if pretrained:
                path = model_path or default_lin_weights()
                if path is None:
                    raise FileNotFoundError("LPIPS(pretrained=True): pass model_path= the reference's network/lpips/weights/v0.1/vgg.pth")
                self.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
        if eval_mode:
            self.eval()

    def _load_torchvision_trunk(self):
        """ImageNet weights of torchvision's vgg16 (what `pn.vgg16(pretrained=True)` loads); needs them in the local torch hub
        cache — there is no network here, so a missing file is an error, not a silent random trunk."""
        import torchvision
        tv = torchvision.models.vgg16(weights=torchvision.models.VGG16_Weights.IMAGENET1K_V1).features.state_dict()
        own = {}
        for k, spec in enumerate(_VGG_SLICES):
            for i, _, _ in spec:
                own["slice%d.%d.weight" % (k + 1, i)] = tv["%d.weight" % i]
                own["slice%d.%d.bias" % (k + 1, i)] = tv["%d.bias" % i]
        self.net.load_state_dict(own, strict=True)

    def forward(self, in0, in1, retPerLayer=False, normalize=False):
        """in0, in1: (N, 3, H, W) in [-1, 1] (or [0, 1] with normalize=True), BGR/RGB as the caller orders them -> (N,1,1,1)."""
        if not in0.is_cuda:
            raise RuntimeError("LPIPS runs on the GPU only (no CPU fallback)")
        if in0.shape != in1.shape or in0.dim() != 4 or in0.shape[1] != 3:
            raise ValueError("LPIPS: expected two (N, 3, H, W) batches of the same shape")
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        if self.version == "0.1":
            in0, in1 = self.scaling_layer(in0), self.scaling_layer(in1)
        N = in0.shape[0]
        x = ops.to_compute(torch.cat([in0, in1], 0))               # ONE trunk pass over the 2N images
        feats = self.net(x)
        ones = None
        vals, per_layer = [], [[] for _ in range(self.L)]
        for n in range(N):
            tot = None
            for k in range(self.L):
                f = feats[k]
                pair = f[[n, N + n]] if N > 1 else f
                if self.lpips:
                    w = self.lins[k].weight.view(-1)
                else:
                    ones = torch.ones(self.chns[k], device=f.device) if ones is None or ones.numel() != self.chns[k] else ones
                    w = ones
                r = _LpipsLayer.apply(pair, w)
                per_layer[k].append(r)
                tot = r if tot is None else tot + r
            vals.append(tot)
        val = torch.stack(vals, 0).view(N, 1, 1, 1)
        if retPerLayer:
            return val, [torch.stack(pl, 0).view(N, 1, 1, 1) for pl in per_layer]
        return val
