"""Operators of the B200 DualStyleUNet (see styleunet.py): autograd wrappers around the sm_100a kernels of
include/agr_styleunet.h.  Activations are NHWC (torch channels_last), fp32 (parity tests) or bf16.

No operator has a CPU / PyTorch fallback; the per-layer and the grouped (whole-net) forms of the weight preparation
share their device code.  Every dense contraction (convolutions and transposed convolutions, forward / data gradient /
weight gradient: include/agr_conv.h) and everything around them (weight modulation/demodulation, noise + bias +
activation, FIR resampling, Haar transforms) is hand-written CUDA; no cuDNN / cuBLAS call remains on the path.

Semantics restated from the reference:
  bias_act            network/styleunet/fused_act.py:100-132, fused_bias_act_kernel.cu:18-65 (act=3)
  upfirdn2d           network/styleunet/upfirdn2d.py:105-227, upfirdn2d_kernel.cu:107-207
  haar_dwt/haar_iwt   dual_styleunet.py:374-425
  modulated_conv2d    dual_styleunet.py:225-300 (fused branch) + :303-313 (noise) + fused lrelu
  equal_conv2d        dual_styleunet.py:93-122 + ConvLayer :329-371
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _lib, stats

# what is still a vendor-library call on the StyleUNet / avatar path (reported by bench.py): no convolution, no GEMM
LIBRARY_OPS = ("CUB DeviceScan + DeviceRadixSort in the rasterizer binning (as in the reference)",)

_p = C.c_void_p
class AgrModWeightItem(C.Structure):
    """include/agr_styleunet.h"""
    _fields_ = [("w", _p), ("s", _p), ("w_out", _p), ("demod", _p), ("d_wout", _p), ("d_w", _p), ("d_s", _p),
                ("scale", C.c_float), ("Cout", C.c_int32), ("Cin", C.c_int32), ("k", C.c_int32),
                ("demodulate", C.c_int32), ("transpose_io", C.c_int32)]


class AgrEqualLinearItem(C.Structure):
    """include/agr_styleunet.h"""
    _fields_ = [("w", _p), ("bias", _p), ("x", _p), ("dy", _p), ("y", _p), ("d_w", _p), ("d_bias", _p), ("d_x", _p),
                ("scale", C.c_float), ("lr_mul", C.c_float), ("out_dim", C.c_int32), ("in_dim", C.c_int32)]


_lib.register_symbols({
    "agr_equal_linear_group_forward": (C.c_int, [C.POINTER(AgrEqualLinearItem), C.c_int32, _p]),
    "agr_equal_linear_group_backward": (C.c_int, [C.POINTER(AgrEqualLinearItem), C.c_int32, _p]),
    "agr_modweight_group_forward": (C.c_int, [C.c_int32, C.POINTER(AgrModWeightItem), C.c_int32, _p]),
    "agr_modweight_group_backward": (C.c_int, [C.c_int32, C.POINTER(AgrModWeightItem), C.c_int32, _p]),
    "agr_upfirdn2d": (C.c_int, [C.c_int32, _p, _p] + [C.c_int32] * 6 + [C.POINTER(C.c_float)] + [C.c_int32] * 6 + [_p]),
    "agr_haar": (C.c_int, [C.c_int32, C.c_int32, _p, _p] + [C.c_int32] * 4 + [_p]),
    "agr_fused_bias_act": (C.c_int, [C.c_int32, _p, _p, _p, _p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, _p]),
    "agr_bias_act_forward": (C.c_int, [C.c_int32, _p, _p, C.c_int64, C.c_int32, _p, _p, _p, C.c_int64, C.c_int32, _p]),
    "agr_bias_act_backward": (C.c_int, [C.c_int32, _p, _p, _p, C.c_int64, C.c_int32, _p, C.c_int64, _p, _p, C.c_int32, _p]),
    "agr_modweight_forward": (C.c_int, [C.c_int32, _p, _p, C.c_float] + [C.c_int32] * 5 + [_p, _p, _p]),
    "agr_modweight_backward": (C.c_int, [C.c_int32, _p, _p, C.c_float] + [C.c_int32] * 5 + [_p, _p, _p, _p, _p]),
    "agr_sum_batch": (C.c_int, [C.c_int32, _p, _p, C.c_int32, C.c_int64, _p]),
    "agr_bilinear2x_add_forward": (C.c_int, [C.c_int32, _p, _p, _p] + [C.c_int32] * 5 + [_p]),
    "agr_bilinear2x_backward": (C.c_int, [C.c_int32, _p, _p] + [C.c_int32] * 4 + [_p]),
    "agr_wavelet_upsample": (C.c_int, [C.c_int32, C.c_int32, _p, _p] + [C.c_int32] * 4 + [C.POINTER(C.c_float), _p]),
    "agr_equal_linear_forward": (C.c_int, [_p, _p, _p, C.c_float, C.c_float, C.c_int32, C.c_int32, _p, _p]),
    "agr_equal_linear_backward": (C.c_int, [_p, _p, _p, C.c_float, C.c_float, C.c_int32, C.c_int32, _p, _p, _p, _p]),
})

_COMPUTE_DTYPE = torch.float32
_CL = torch.channels_last


def set_compute_dtype(dtype):
    """torch.float32 (parity tests) or torch.bfloat16 (BASELINE config 4: 'bf16 StyleUNet')."""
    global _COMPUTE_DTYPE
    assert dtype in (torch.float32, torch.bfloat16)
    _COMPUTE_DTYPE = dtype


def compute_dtype():
    return _COMPUTE_DTYPE


def to_compute(x):
    return x.to(_COMPUTE_DTYPE).contiguous(memory_format=_CL)


def from_compute(x):
    return x.float().contiguous()


def _code(t):
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise TypeError("StyleUNet kernels take fp32 or bf16 activations, got %s" % t.dtype)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(st, what):
    if st != _lib.AGR_OK:
        raise RuntimeError("%s failed (status %d) %s" % (what, st, _lib.cuda_error_string() if st == _lib.AGR_ERR_CUDA else ""))


_TRACE_COPIES = bool(int(__import__("os").environ.get("AGR_TRACE_COPIES", "0")))


def _nhwc(x):
    if not x.is_cuda:
        raise RuntimeError("animatablegaussians_b200 StyleUNet operators are CUDA-only (no CPU fallback)")
    if _TRACE_COPIES and x.ndim == 4 and not x.is_contiguous(memory_format=_CL) and x.numel() > (1 << 24):
        import traceback
        fr = [f for f in traceback.extract_stack(limit=6)][:-1]
        print("[agr copy] %s %s strides=%s  <- %s" % (tuple(x.shape), x.dtype, x.stride(), " < ".join("%s:%d" % (f.name, f.lineno) for f in reversed(fr))))
    return x.contiguous(memory_format=_CL)


def _new_like(x, C_, H, W):
    return torch.empty((x.shape[0], C_, H, W), dtype=x.dtype, device=x.device, memory_format=_CL)


# ------------------------------------------------------------------------------------------ FIR resampling
import contextlib  # noqa: E402
import weakref  # noqa: E402

# ------------------------------------------------------------------------------------------ zero-initialised scratch
# The reduction targets of the backward kernels (bias / noise-weight / style gradients, accumulated with atomics) must
# start at zero.  One torch.zeros per target is ~530 fill launches per train step (r01 profile); inside `step_arena()`
# they are carved from chunks that are zero-filled once.  A chunk is never handed out twice and dies with its last
# slice, so nothing carries over between steps or across CUDA-graph capture boundaries.
_arena = None


@contextlib.contextmanager
def step_arena(chunk_floats=1 << 18):
    """Wrap ONE forward+backward (eager, or the body of a CUDA-graph capture)."""
    global _arena
    prev, _arena = _arena, {"bufs": {}, "chunk": int(chunk_floats)}
    try:
        yield
    finally:
        _arena = prev


def _zeros(n, dev):
    """Chunks are per (device, stream): a chunk is zero-filled on the stream that asked for it and only handed to work
    queued on that same stream (the nets of a step run on parallel streams, forward and backward)."""
    a = _arena
    if a is None or n > a["chunk"]:
        return torch.zeros(n, dtype=torch.float32, device=dev)
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    slot = a["bufs"].get(key)
    if slot is None or slot[1] + n > a["chunk"]:
        slot = [torch.zeros(a["chunk"], dtype=torch.float32, device=dev), 0]
        a["bufs"][key] = slot
    out = slot[0][slot[1]:slot[1] + n]
    slot[1] += (n + 3) // 4 * 4   # keep every slice 16-byte aligned
    return out


_taps_cache = {}


def _host_taps(kernel, flip):
    """(kh,kw) device buffer -> ctypes float array (flipped or not).  Cached per tensor OBJECT (weak reference +
    version counter, so a recycled address or an in-place update can never serve stale taps): one D2H per FIR buffer
    of the module, none afterwards (required for CUDA-graph capture)."""
    key = (id(kernel), flip)
    hit = _taps_cache.get(key)
    if hit is not None and hit[0]() is kernel and hit[1] == kernel._version:
        return hit[2]
    k = kernel.detach().float().cpu()
    if flip:
        k = torch.flip(k, [0, 1])
    vals = [float(v) for v in k.reshape(-1)]
    out = ((C.c_float * len(vals))(*vals), int(k.shape[0]), int(k.shape[1]))
    if len(_taps_cache) > 4096:
        _taps_cache.clear()
    _taps_cache[key] = (weakref.ref(kernel), kernel._version, out)
    return out


def _fir_launch(x, taps, kh, kw, up, down, px0, py0, out_h, out_w):
    lib = _lib.load()
    y = _new_like(x, x.shape[1], out_h, out_w)
    with torch.cuda.device(x.device), stats.stage("styleunet_fir", launches=1):
        _check(lib.agr_upfirdn2d(_code(x), _ptr(x), _ptr(y), x.shape[0], x.shape[2], x.shape[3], x.shape[1], out_h, out_w,
                                 C.cast(taps, C.POINTER(C.c_float)), kh, kw, up, down, px0, py0, _stream(x)), "agr_upfirdn2d")
    return y


class _UpFirDn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, up, down, pad):
        x = _nhwc(x)
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        H, W = x.shape[2], x.shape[3]
        out_h = (H * up + py0 + py1 - kh + down) // down
        out_w = (W * up + px0 + px1 - kw + down) // down
        taps, _, _ = _host_taps(kernel, True)
        ctx.kernel, ctx.cfg = kernel, (up, down, px0, py0, kh, kw, H, W)
        return _fir_launch(x, taps, kh, kw, up, down, px0, py0, out_h, out_w)

    @staticmethod
    def backward(ctx, g):
        up, down, px0, py0, kh, kw, H, W = ctx.cfg
        taps, _, _ = _host_taps(ctx.kernel, False)  # adjoint: correlate with the un-flipped kernel
        gx = _fir_launch(_nhwc(g), taps, kh, kw, down, up, kw - px0 - 1, kh - py0 - 1, H, W)
        return gx, None, None, None, None


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    return _UpFirDn.apply(x, kernel, up, down, tuple(int(p) for p in pad))


class _Haar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, analysis):
        x = _nhwc(x)
        ctx.analysis = analysis
        return _Haar._run(x, analysis)

    @staticmethod
    def _run(x, analysis):
        lib = _lib.load()
        B, Cc, H, W = x.shape
        y = _new_like(x, Cc * 4, H // 2, W // 2) if analysis else _new_like(x, Cc // 4, H * 2, W * 2)
        with torch.cuda.device(x.device), stats.stage("styleunet_fir", launches=1):
            _check(lib.agr_haar(_code(x), 0 if analysis else 2, _ptr(x), _ptr(y), B, H, W, Cc, _stream(x)), "agr_haar")
        return y

    @staticmethod
    def backward(ctx, g):
        return _Haar._run(_nhwc(g), not ctx.analysis), None  # orthonormal: adjoint == inverse


def haar_dwt(x):
    """HaarTransform.forward (dual_styleunet.py:398-404): C -> 4C at half resolution, order ll|lh|hl|hh."""
    return _Haar.apply(x, True)


def haar_iwt(x):
    """InverseHaarTransform.forward (dual_styleunet.py:418-425): 4C -> C at double resolution."""
    return _Haar.apply(x, False)


def wavelet_upsample_taps(kernel):
    """Collapse dwt(upfirdn2d(iwt(.), kernel, up=2, pad=(2,1))) into the two 256-entry filter banks of
    agr_wavelet_upsample.  `kernel`: the (4,4) FIR as the Upsample module holds it (numpy).  The operator is assembled
    from the definitions of its stages on a small grid in float64; an interior row / column of it is the bank.
    Returns (forward[pi][pj][bo][bi][di][dj], adjoint[bi][bo][a][b]) as flat float64 arrays."""
    import numpy as np
    k = np.asarray(kernel, np.float64)
    assert k.shape == (4, 4)
    taps = k[::-1, ::-1]                      # upfirdn2d convolves: correlation with the flipped kernel
    sgn = np.array([[[1, 1], [1, 1]], [[1, 1], [-1, -1]], [[1, -1], [1, -1]], [[1, -1], [-1, 1]]], np.float64)  # [band][row][col]
    h = 6

    def chain(skip):                          # (h,h,4) -> (2h,2h,4)
        img = np.zeros((2 * h, 2 * h))
        for p in range(2):
            for q in range(2):                # InverseHaarTransform: pixel (p,q) of each 2x2 block
                img[p::2, q::2] = 0.5 * sum(sgn[b, p, q] * skip[:, :, b] for b in range(4))
        z = np.zeros((4 * h + 3, 4 * h + 3))  # zero-inserted image, pad (2, 1)
        z[2:2 + 4 * h:2, 2:2 + 4 * h:2] = img
        up = np.zeros((4 * h, 4 * h))
        for ky in range(4):
            for kx in range(4):
                up += taps[ky, kx] * z[ky:ky + 4 * h, kx:kx + 4 * h]
        out = np.zeros((2 * h, 2 * h, 4))
        for b in range(4):                    # HaarTransform
            out[:, :, b] = 0.5 * sum(sgn[b, p, q] * up[p::2, q::2] for p in range(2) for q in range(2))
        return out

    A = np.zeros((2 * h, 2 * h, 4, h, h, 4))
    for m in range(h):
        for n in range(h):
            for b in range(4):
                e = np.zeros((h, h, 4)); e[m, n, b] = 1.0
                A[:, :, :, m, n, b] = chain(e)
    m0 = 2
    fwd = np.zeros((2, 2, 4, 4, 2, 2))
    for pi in range(2):
        for pj in range(2):
            row = A[2 * m0 + pi, 2 * m0 + pj].copy()          # (bo, m, n, bi)
            for di in range(2):
                for dj in range(2):
                    mm, nn = m0 + di + pi - 1, m0 + dj + pj - 1
                    fwd[pi, pj, :, :, di, dj] = row[:, mm, nn, :]
                    row[:, mm, nn, :] = 0
            assert np.abs(row).max() < 1e-12, "wavelet upsample: support is not 2x2"
    adj = np.zeros((4, 4, 4, 4))
    col = A[:, :, :, m0, m0, :].copy()                         # (i, j, bo, bi)
    for a in range(4):
        for b in range(4):
            adj[:, :, a, b] = col[2 * m0 - 1 + a, 2 * m0 - 1 + b].T
            col[2 * m0 - 1 + a, 2 * m0 - 1 + b] = 0
    assert np.abs(col).max() < 1e-12, "wavelet upsample: adjoint support is not 4x4"
    return fwd.reshape(-1), adj.reshape(-1)


_wavelet_cache = {}


def _wavelet_taps(kernel):
    """ctypes banks for a module's FIR buffer, cached like _host_taps (one D2H per buffer, none under graph capture)."""
    key = id(kernel)
    hit = _wavelet_cache.get(key)
    if hit is not None and hit[0]() is kernel and hit[1] == kernel._version:
        return hit[2]
    f, a = wavelet_upsample_taps(kernel.detach().double().cpu().numpy())
    out = ((C.c_float * 256)(*[float(v) for v in f]), (C.c_float * 256)(*[float(v) for v in a]))
    if len(_wavelet_cache) > 1024:
        _wavelet_cache.clear()
    _wavelet_cache[key] = (weakref.ref(kernel), kernel._version, out)
    return out


class _WaveletUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, skip, kernel):
        skip = _nhwc(skip)
        ctx.kernel = kernel
        return _WaveletUp._run(skip, kernel, False)

    @staticmethod
    def _run(x, kernel, adjoint):
        lib = _lib.load()
        B, C4, H, W = x.shape
        h, w = (H // 2, W // 2) if adjoint else (H, W)
        y = _new_like(x, C4, h, w) if adjoint else _new_like(x, C4, 2 * h, 2 * w)
        taps = _wavelet_taps(kernel)[1 if adjoint else 0]
        with torch.cuda.device(x.device), stats.stage("styleunet_fir", launches=1):
            _check(lib.agr_wavelet_upsample(_code(x), int(adjoint), _ptr(x), _ptr(y), B, h, w, C4 // 4,
                                            C.cast(taps, C.POINTER(C.c_float)), _stream(x)), "agr_wavelet_upsample")
        return y

    @staticmethod
    def backward(ctx, g):
        return _WaveletUp._run(_nhwc(g), ctx.kernel, True), None


def wavelet_upsample(skip, up_kernel):
    """ToRGB skip path (dual_styleunet.py:624-631): dwt(upsample(iwt(skip))), one fused pass for the 4-tap FIR."""
    if tuple(up_kernel.shape) == (4, 4) and skip.shape[1] % 4 == 0:
        return _WaveletUp.apply(skip, up_kernel)
    p = up_kernel.shape[0] - 2
    y = haar_iwt(skip)
    y = upfirdn2d(y, up_kernel, up=2, pad=((p + 1) // 2 + 1, p // 2))
    return haar_dwt(y)


class _ExpandBatch(torch.autograd.Function):
    """(1,C,H,W) -> (V,C,H,W) broadcast view of the shared prefix state; the adjoint sums the V gradients with one
    streaming kernel."""

    @staticmethod
    def forward(ctx, x, V):
        ctx.V = V
        return x.expand(V, -1, -1, -1)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        g = _nhwc(g)
        y = _new_like(g[:1], g.shape[1], g.shape[2], g.shape[3])
        n = y.numel()
        with torch.cuda.device(g.device), stats.stage("styleunet_act", launches=1):
            _check(lib.agr_sum_batch(_code(g), _ptr(g), _ptr(y), ctx.V, n, _stream(g)), "agr_sum_batch")
        return y, None


def expand_batch(x, V):
    if V == 1 or x.shape[0] == V:
        return x
    if x.numel() % 8:
        return x.expand(V, -1, -1, -1)
    return _ExpandBatch.apply(_nhwc(x), V)


# ------------------------------------------------------------------------------------------ bias + noise + act
class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, noise, noise_weight, activate):
        lib = _lib.load()
        is4 = x.ndim == 4
        x = _nhwc(x) if is4 else x.contiguous()
        Cc = x.shape[1]
        pixels = x.numel() // Cc
        y = torch.empty_like(x)
        b = bias.detach().float().contiguous() if bias is not None else None
        nz = noise.detach().float().contiguous() if noise is not None else None
        nw = noise_weight.detach().float().contiguous() if noise_weight is not None else None
        nper = nz.numel() if nz is not None else 1
        if nz is not None and pixels % nper != 0:
            raise RuntimeError("noise must have one value per pixel (one image shared by the batch)")
        with torch.cuda.device(x.device), stats.stage("styleunet_act", launches=1):
            _check(lib.agr_bias_act_forward(_code(x), _ptr(x), _ptr(y), pixels, Cc, _ptr(b), _ptr(nz), _ptr(nw), nper, int(activate),
                                            _stream(x)), "agr_bias_act_forward")
        ctx.save_for_backward(y if activate else None, nz)
        ctx.meta = (activate, is4, bias is not None, noise_weight is not None and noise is not None, Cc, pixels,
                    None if bias is None else bias.shape, None if noise_weight is None else noise_weight.shape)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        y, nz = ctx.saved_tensors
        activate, is4, has_b, has_n, Cc, pixels, bshape, nshape = ctx.meta
        g = _nhwc(g) if is4 else g.contiguous()
        dx = torch.empty_like(g)
        db = _zeros(Cc, g.device) if has_b else None
        dn = _zeros(1, g.device) if has_n else None
        with torch.cuda.device(g.device), stats.stage("styleunet_act", launches=1):
            _check(lib.agr_bias_act_backward(_code(g), _ptr(g), _ptr(y), _ptr(dx), pixels, Cc, _ptr(nz) if has_n else None,
                                             nz.numel() if has_n else 1, _ptr(db), _ptr(dn), int(activate), _stream(g)),
                   "agr_bias_act_backward")
        return dx, (db.view(bshape) if has_b else None), None, (dn.view(nshape) if has_n else None), None


def bias_act(x, bias=None, noise=None, noise_weight=None, activate=True):
    """y = lrelu(x + w_noise * noise + bias[c], 0.2) * sqrt(2)   (activate=False: no lrelu / gain)."""
    return _BiasAct.apply(x, bias, noise, noise_weight, activate)


class _AddViewFeature(torch.autograd.Function):
    """out + bilinear_2x(view_feature) (dual_styleunet.py:881-883) in one pass; out may be the shared (batch-1) state."""

    @staticmethod
    def forward(ctx, base, vf):
        lib = _lib.load()
        base = _nhwc(base)
        vf32 = vf.detach().float().contiguous(memory_format=_CL)
        V, Cc, h, w = vf32.shape
        y = torch.empty((V, Cc, 2 * h, 2 * w), dtype=base.dtype, device=base.device, memory_format=_CL)
        with torch.cuda.device(base.device), stats.stage("styleunet_act", launches=1):
            _check(lib.agr_bilinear2x_add_forward(_code(base), _ptr(vf32), _ptr(base), _ptr(y), V, base.shape[0], h, w, Cc, _stream(base)),
                   "agr_bilinear2x_add_forward")
        ctx.meta = (base.shape[0], V, Cc, h, w, vf.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        Vb, V, Cc, h, w, vdt = ctx.meta
        g = _nhwc(g)
        d_base = d_vf = None
        if ctx.needs_input_grad[0]:
            if Vb == V:
                d_base = g
            else:
                d_base = _new_like(g[:1], Cc, 2 * h, 2 * w)
                with torch.cuda.device(g.device), stats.stage("styleunet_act", launches=1):
                    _check(lib.agr_sum_batch(_code(g), _ptr(g), _ptr(d_base), V, d_base.numel(), _stream(g)), "agr_sum_batch")
        if ctx.needs_input_grad[1]:
            d_vf = torch.empty((V, Cc, h, w), dtype=torch.float32, device=g.device, memory_format=_CL)
            with torch.cuda.device(g.device), stats.stage("styleunet_act", launches=1):
                _check(lib.agr_bilinear2x_backward(_code(g), _ptr(g), _ptr(d_vf), V, h, w, Cc, _stream(g)), "agr_bilinear2x_backward")
            d_vf = d_vf.to(vdt)
        return d_base, d_vf


def add_view_feature(out, vf):
    """out (1 or V, C, H, W) + F.interpolate(vf (V, C, H/2, W/2), (H, W), mode='bilinear')  ->  (V, C, H, W)."""
    if out.shape[2] == 2 * vf.shape[2] and out.shape[3] == 2 * vf.shape[3] and vf.shape[1] % 8 == 0 and out.shape[0] in (1, vf.shape[0]):
        return _AddViewFeature.apply(out, vf)
    return expand_batch(out, vf.shape[0]) + bilinear_resize(vf, out.shape[-2:])


def bilinear_resize(x, size):
    """F.interpolate(mode='bilinear') of the view feature (dual_styleunet.py:882,901)."""
    return F.interpolate(x.float(), size, mode="bilinear").to(_COMPUTE_DTYPE).contiguous(memory_format=_CL)


# ------------------------------------------------------------------------------------------ weight preparation
class _ModWeight(torch.autograd.Function):
    """(Cout,Cin,k,k) fp32 master weight + (Cin,) style -> conv-ready weight in the compute dtype, KRSC memory
    (torch channels_last), appended to `sink`; the autograd output is its fp32 gradient HANDLE (see _ModWeightGroup).
    transpose_io -> (Cin,Cout,k,k) layout."""

    @staticmethod
    def forward(ctx, weight, s, scale, demodulate, transpose_io, dtype, sink):
        lib = _lib.load()
        w = weight.detach().float().contiguous()
        Cout, Cin, k = w.shape[-4], w.shape[-3], w.shape[-1]
        sv = s.detach().float().contiguous().view(-1)
        shape = (Cin, Cout, k, k) if transpose_io else (Cout, Cin, k, k)
        out = torch.empty(shape, dtype=dtype, device=w.device, memory_format=_CL)
        demod = torch.empty(Cout, dtype=torch.float32, device=w.device) if demodulate else None
        with torch.cuda.device(w.device), stats.stage("styleunet_weight", launches=1):
            _check(lib.agr_modweight_forward(_code(out), _ptr(w), _ptr(sv), float(scale), Cout, Cin, k, int(demodulate),
                                             int(transpose_io), _ptr(out), _ptr(demod), _stream(w)), "agr_modweight_forward")
        sink.append(out)
        ctx.save_for_backward(w, sv, demod)
        ctx.meta = (float(scale), Cout, Cin, k, demodulate, transpose_io, weight.shape, s.shape)
        return _zero_scalar(w.device).expand(shape)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        w, sv, demod = ctx.saved_tensors
        scale, Cout, Cin, k, demodulate, transpose_io, wshape, sshape = ctx.meta
        g = g.contiguous(memory_format=_CL)
        dw = torch.empty_like(w)
        ds = _zeros(Cin, w.device)
        with torch.cuda.device(w.device), stats.stage("styleunet_weight", launches=1):
            _check(lib.agr_modweight_backward(_code(g), _ptr(w), _ptr(sv), scale, Cout, Cin, k, int(demodulate), int(transpose_io),
                                              _ptr(g), _ptr(demod), _ptr(dw), _ptr(ds), _stream(w)), "agr_modweight_backward")
        return dw.view(wshape), ds.view(sshape), None, None, None, None, None


def mod_weight(weight, s, scale, demodulate, dtype, transpose_io=False):
    """-> (conv-ready operand, fp32 gradient handle) of one layer (per-layer form of prepare_weights)."""
    if s is None:
        s = _ones(weight.shape[-3], weight.device)
    sink = []
    handle = _ModWeight.apply(weight, s, scale, demodulate, transpose_io, dtype, sink)
    return sink[0], handle


def _addr(t):
    return None if t is None else t.data_ptr()


class _ModWeightGroup(torch.autograd.Function):
    """_ModWeight for every conv layer of a U-Net at once (agr_modweight_group_*): `specs[l] = (scale, demodulate,
    transpose_io)`, `tensors = (weight_0, s_0, weight_1, s_1, ...)` with s_l = None for a plain equalised conv.

    The conv-ready operands (compute dtype) are appended to `sink`; the autograd outputs are fp32 HANDLES of the
    operands' logical shape (stride-0 views of one zero: no memory, never read).  The convolutions return their fp32
    weight gradients against the handle, so dW reaches the modulation backward in fp32 without a cast through the
    operand's bf16 dtype (autograd would otherwise cast every gradient to the dtype of the forward output)."""

    @staticmethod
    def forward(ctx, specs, dtype, sink, *tensors):
        lib = _lib.load()
        L = len(specs)
        ws = [tensors[2 * l].detach().float().contiguous() for l in range(L)]
        dev = ws[0].device
        svs, outs, demods, handles = [], [], [], []
        items = (AgrModWeightItem * L)()
        zero = _zero_scalar(dev)
        for l, (scale, demodulate, transpose_io) in enumerate(specs):
            w, s = ws[l], tensors[2 * l + 1]
            Cout, Cin, k = w.shape[-4], w.shape[-3], w.shape[-1]
            sv = _ones(Cin, dev) if s is None else s.detach().float().contiguous().view(-1)
            shape = (Cin, Cout, k, k) if transpose_io else (Cout, Cin, k, k)
            out = torch.empty(shape, dtype=dtype, device=dev, memory_format=_CL)
            demod = torch.empty(Cout, dtype=torch.float32, device=dev) if demodulate else None
            svs.append(sv); outs.append(out); demods.append(demod); handles.append(zero.expand(shape))
            it = items[l]
            it.w, it.s, it.w_out, it.demod = w.data_ptr(), sv.data_ptr(), out.data_ptr(), _addr(demod)
            it.scale, it.Cout, it.Cin, it.k = float(scale), Cout, Cin, k
            it.demodulate, it.transpose_io = int(bool(demodulate)), int(bool(transpose_io))
        with torch.cuda.device(dev), stats.stage("styleunet_weight", launches=(L + 39) // 40):
            _check(lib.agr_modweight_group_forward(_code(outs[0]), items, L, _stream(ws[0])), "agr_modweight_group_forward")
        sink.extend(outs)
        ctx.save_for_backward(*ws, *svs, *[d for d in demods if d is not None])
        ctx.meta = (specs, [d is not None for d in demods], [tensors[2 * l].shape for l in range(L)],
                    [None if tensors[2 * l + 1] is None else tensors[2 * l + 1].shape for l in range(L)])
        return tuple(handles)

    @staticmethod
    def backward(ctx, *gs):
        lib = _lib.load()
        specs, has_demod, wshapes, sshapes = ctx.meta
        L = len(specs)
        saved = ctx.saved_tensors
        ws, svs, rest = saved[:L], saved[L:2 * L], list(saved[2 * L:])
        dev = ws[0].device
        items = (AgrModWeightItem * L)()
        keep, grads = [], [None, None, None]
        for l, (scale, demodulate, transpose_io) in enumerate(specs):
            w = ws[l]
            Cout, Cin, k = w.shape[-4], w.shape[-3], w.shape[-1]
            demod = rest.pop(0) if has_demod[l] else None
            g = gs[l]
            if g is None:   # a layer this forward never used
                shape = (Cin, Cout, k, k) if transpose_io else (Cout, Cin, k, k)
                g = torch.zeros(shape, dtype=torch.float32, device=dev)
            g = g.float().contiguous(memory_format=_CL)
            dw = torch.empty_like(w)
            ds = _zeros(Cin, dev) if sshapes[l] is not None and ctx.needs_input_grad[4 + 2 * l] else None
            keep.append((g, dw, ds))
            it = items[l]
            it.w, it.s, it.demod, it.d_wout = w.data_ptr(), svs[l].data_ptr(), _addr(demod), g.data_ptr()
            it.d_w, it.d_s = dw.data_ptr(), _addr(ds)
            it.scale, it.Cout, it.Cin, it.k = float(scale), Cout, Cin, k
            it.demodulate, it.transpose_io = int(bool(demodulate)), int(bool(transpose_io))
            grads.append(dw.view(wshapes[l]))
            grads.append(ds.view(sshapes[l]) if ds is not None else None)
        with torch.cuda.device(dev), stats.stage("styleunet_weight", launches=(L + 39) // 40):
            _check(lib.agr_modweight_group_backward(0, items, L, _stream(ws[0])), "agr_modweight_group_backward")
        return tuple(grads)


_zero_cache = {}


def _zero_scalar(dev):
    key = str(dev)
    if key not in _zero_cache:
        _zero_cache[key] = torch.zeros(1, dtype=torch.float32, device=dev)
    return _zero_cache[key]


# A "weight plan" holds the conv-ready operands of every layer of one U-Net for one step, prepared by the grouped
# kernels; layers look their weight up by parameter identity and fall back to the per-layer op outside a plan.
_plan = None


@contextlib.contextmanager
def weight_plan(plan):
    global _plan
    prev, _plan = _plan, plan
    try:
        yield
    finally:
        _plan = prev


def planned_weight(weight, dtype):
    """-> (operand, gradient handle) of this parameter in the active weight plan, or None."""
    if _plan is None:
        return None
    hit = _plan.get(id(weight))
    return hit if hit is not None and hit[0].dtype == dtype else None


def prepare_weights(entries, dtype):
    """entries: [(weight parameter, style modulation s (1,Cin) or None, scale, demodulate)] ->
    {id(weight): (conv-ready KRSC operand in `dtype`, fp32 gradient handle)}."""
    if not entries:
        return {}
    specs = tuple((float(e[2]), bool(e[3]), False) for e in entries)
    flat = []
    for e in entries:
        flat += [e[0], e[1]]
    sink = []
    handles = _ModWeightGroup.apply(specs, dtype, sink, *flat)
    return {id(e[0]): (o, h) for e, o, h in zip(entries, sink, handles)}


class _EqualLinearVec(torch.autograd.Function):
    """EqualLinear (no activation) on one style vector: lr_mul * bias + scale * W x, one launch each way."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, lr_mul):
        lib = _lib.load()
        out_dim, in_dim = weight.shape
        xv, w = x.detach().contiguous().view(-1), weight.detach().contiguous()
        b = bias.detach().contiguous() if bias is not None else None
        y = torch.empty(out_dim, dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device), stats.stage("styleunet_weight", launches=1):
            _check(lib.agr_equal_linear_forward(_ptr(w), _ptr(b), _ptr(xv), float(scale), float(lr_mul), out_dim, in_dim, _ptr(y),
                                                _stream(w)), "agr_equal_linear_forward")
        ctx.save_for_backward(xv, w)
        ctx.meta = (float(scale), float(lr_mul), bias is not None, x.shape)
        return y.view(1, out_dim)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        xv, w = ctx.saved_tensors
        scale, lr_mul, has_b, xshape = ctx.meta
        out_dim, in_dim = w.shape
        g = g.contiguous().view(-1)
        dw = torch.empty_like(w)
        db = torch.empty(out_dim, dtype=torch.float32, device=w.device) if has_b else None
        dx = _zeros(in_dim, w.device) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(w.device), stats.stage("styleunet_weight", launches=1):
            _check(lib.agr_equal_linear_backward(_ptr(w), _ptr(xv), _ptr(g), scale, lr_mul, out_dim, in_dim, _ptr(dw), _ptr(db),
                                                 _ptr(dx), _stream(w)), "agr_equal_linear_backward")
        return (dx.view(xshape) if dx is not None else None), dw, db, None, None


def equal_linear(x, weight, bias, scale, lr_mul):
    """dual_styleunet.py:155-158.  The fused kernels cover the case the U-Nets run 108 times per step (one fp32 style
    vector); a batch of styles is a plain library GEMM."""
    if x.is_cuda and x.dim() == 2 and x.shape[0] == 1 and x.dtype == torch.float32 and weight.dtype == torch.float32:
        return _EqualLinearVec.apply(x, weight, bias, scale, lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul if bias is not None else None)


class _EqualLinearGroup(torch.autograd.Function):
    """y_l = EqualLinear_l(latent[0, idx_l]) for every modulation layer of a U-Net in one launch each way;
    `tensors = (weight_0, bias_0, weight_1, bias_1, ...)`, `specs[l] = (scale, lr_mul)`.  The style gradients of all
    layers accumulate straight into one d_latent buffer."""

    @staticmethod
    def forward(ctx, latent, idx, specs, *tensors):
        lib = _lib.load()
        L = len(idx)
        lat = latent.detach().float().contiguous()
        D = lat.shape[-1]
        dev = lat.device
        ws = [tensors[2 * l].detach().contiguous() for l in range(L)]
        items = (AgrEqualLinearItem * L)()
        ys, keep = [], []
        for l in range(L):
            w, b = ws[l], tensors[2 * l + 1]
            b = b.detach().contiguous() if b is not None else None
            out_dim, in_dim = w.shape
            if in_dim != D or w.dtype != torch.float32:
                raise RuntimeError("equal_linear group: weight %s does not match the latent width %d" % (tuple(w.shape), D))
            y = torch.empty(out_dim, dtype=torch.float32, device=dev)
            ys.append(y); keep.append(b)
            it = items[l]
            it.w, it.bias, it.x, it.y = w.data_ptr(), _addr(b), lat.data_ptr() + 4 * D * int(idx[l]), y.data_ptr()
            it.scale, it.lr_mul, it.out_dim, it.in_dim = float(specs[l][0]), float(specs[l][1]), out_dim, in_dim
        with torch.cuda.device(dev), stats.stage("styleunet_weight", launches=(L + 39) // 40):
            _check(lib.agr_equal_linear_group_forward(items, L, _stream(lat)), "agr_equal_linear_group_forward")
        ctx.save_for_backward(lat, *ws)
        ctx.meta = (tuple(int(i) for i in idx), specs, [tensors[2 * l + 1] is not None for l in range(L)], latent.shape)
        return tuple(y.view(1, -1) for y in ys)

    @staticmethod
    def backward(ctx, *gs):
        lib = _lib.load()
        idx, specs, has_b, lshape = ctx.meta
        L = len(idx)
        lat, ws = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        D = lat.shape[-1]
        dev = lat.device
        d_lat = _zeros(lat.numel(), dev) if ctx.needs_input_grad[0] else None
        items = (AgrEqualLinearItem * L)()
        keep, grads = [], [None, None, None]
        for l in range(L):
            w = ws[l]
            out_dim, in_dim = w.shape
            g = gs[l]
            g = torch.zeros(out_dim, dtype=torch.float32, device=dev) if g is None else g.float().contiguous().view(-1)
            dw = torch.empty_like(w)
            db = torch.empty(out_dim, dtype=torch.float32, device=dev) if has_b[l] else None
            keep.append(g)
            it = items[l]
            it.w, it.x, it.dy = w.data_ptr(), lat.data_ptr() + 4 * D * idx[l], g.data_ptr()
            it.d_w, it.d_bias = dw.data_ptr(), _addr(db)
            it.d_x = None if d_lat is None else d_lat.data_ptr() + 4 * D * idx[l]
            it.scale, it.lr_mul, it.out_dim, it.in_dim = float(specs[l][0]), float(specs[l][1]), out_dim, in_dim
            grads += [dw, db]
        with torch.cuda.device(dev), stats.stage("styleunet_weight", launches=(L + 39) // 40):
            _check(lib.agr_equal_linear_group_backward(items, L, _stream(lat)), "agr_equal_linear_group_backward")
        grads[0] = d_lat.view(lshape) if d_lat is not None else None
        return tuple(grads)


def equal_linear_group(latent, layers):
    """layers: [(EqualLinear-like module with weight / bias / scale / lr_mul, index into latent[0])] -> [(1, out_dim)]."""
    flat = []
    for m, _ in layers:
        flat += [m.weight, m.bias]
    return _EqualLinearGroup.apply(latent, tuple(i for _, i in layers), tuple((m.scale, m.lr_mul) for m, _ in layers), *flat)


_ones_cache = {}


def _ones(n, dev):
    key = (n, str(dev))
    if key not in _ones_cache:
        _ones_cache[key] = torch.ones(n, dtype=torch.float32, device=dev)
    return _ones_cache[key]


# ------------------------------------------------------------------------------------------ dense contractions
class AgrConvGeom(C.Structure):
    """include/agr_conv.h"""
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "Cin", "OH", "OW", "Cout", "ksize", "stride", "pad", "transposed")]


class AgrConvEpilogue(C.Structure):
    """include/agr_conv.h"""
    _fields_ = [("bias", _p), ("noise", _p), ("noise_w", _p), ("residual", _p), ("activate", C.c_int32), ("out_fp32", C.c_int32),
                ("w_cin_total", C.c_int32), ("w_cin_offset", C.c_int32)]


_lib.register_symbols({
    "agr_conv2d_path": (C.c_int, [C.c_int32, C.POINTER(AgrConvGeom), C.c_int32]),
    "agr_conv2d_forward": (C.c_int, [C.c_int32, C.POINTER(AgrConvGeom), _p, _p, _p, C.POINTER(AgrConvEpilogue), _p]),
    "agr_conv2d_dgrad": (C.c_int, [C.c_int32, C.POINTER(AgrConvGeom), _p, _p, _p, _p]),
    "agr_conv2d_dgrad_krsc": (C.c_int, [C.c_int32, C.POINTER(AgrConvGeom), _p, _p, C.c_int32, C.c_int32, _p, _p]),
    "agr_conv2d_wgrad": (C.c_int, [C.c_int32, C.POINTER(AgrConvGeom), _p, _p, _p, C.c_int32, C.c_int32, C.c_int32, _p]),
    "agr_weight_transpose": (C.c_int, [C.c_int32, _p, _p, C.c_int32, C.c_int32, C.c_int32, _p]),
    "agr_conv2d_set_generation": (C.c_int, [C.c_int32]),
    "agr_conv2d_set_wgrad_split": (C.c_int, [C.c_int32, C.c_int32]),
})

_STAGE = {1: "styleunet_conv_tc", 2: "styleunet_conv_direct"}


def conv_geom(x_shape, Cout, k, stride=1, pad=None, transposed=False):
    """Geometry of one layer (include/agr_conv.h): x_shape = (N, Cin, H, W), the logical NCHW shape of the input."""
    N, Cin, H, W = (int(v) for v in x_shape)
    pad = k // 2 if pad is None else int(pad)
    if transposed:
        OH, OW = (H - 1) * stride - 2 * pad + k, (W - 1) * stride - 2 * pad + k
    else:
        OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    return AgrConvGeom(N, H, W, Cin, OH, OW, int(Cout), int(k), int(stride), pad, int(bool(transposed)))


def _label(what, g):
    return "%s N%d %dx%d %d->%d k%d s%d%s" % (what, g.N, g.H, g.W, g.Cin, g.Cout, g.ksize, g.stride, " T" if g.transposed else "")


def _flops(g):
    pix = g.N * (g.H * g.W if g.transposed else g.OH * g.OW)
    return 2.0 * pix * g.Cin * g.Cout * g.ksize * g.ksize


def conv_path(x, g, what=0):
    """1 = tcgen05, 2 = CUDA-core direct (agr_conv2d_path); what: 0 forward, 1 data gradient, 2 weight gradient."""
    return int(_lib.load().agr_conv2d_path(_code(x), C.byref(g), what))


def conv_forward(x, w, g, bias=None, noise=None, noise_w=None, activate=0, residual=None, out_fp32=False, cin_total=0, cin_offset=0):
    """x (N,Cin,H,W) NHWC, w KRSC (Cout, cin_total or Cin, k, k) channels_last -> y (N,Cout,OH,OW) NHWC."""
    lib = _lib.load()
    y = torch.empty((g.N, g.Cout, g.OH, g.OW), dtype=torch.float32 if out_fp32 else x.dtype, device=x.device, memory_format=_CL)
    ep = AgrConvEpilogue(_addr(bias), _addr(noise), _addr(noise_w if noise is not None else None), _addr(residual), int(activate),
                         int(bool(out_fp32)), int(cin_total), int(cin_offset))
    stage = _STAGE.get(conv_path(x, g, 0), "styleunet_conv_direct")
    stats.add_work(stage, _flops(g))
    with torch.cuda.device(x.device), stats.stage(stage, launches=1, label=_label("fwd", g)):
        _check(lib.agr_conv2d_forward(_code(x), C.byref(g), _ptr(x), _ptr(w), _ptr(y), C.byref(ep), _stream(x)), "agr_conv2d_forward")
    return y


def weight_transpose(w):
    """KRSC (Cout,Cin,k,k) channels_last -> (Cin,Cout,k,k) channels_last: the operand of the data-gradient calls."""
    lib = _lib.load()
    Cout, Cin, k = w.shape[0], w.shape[1], w.shape[-1]
    wt = torch.empty((Cin, Cout, k, k), dtype=w.dtype, device=w.device, memory_format=_CL)
    with torch.cuda.device(w.device), stats.stage("styleunet_weight", launches=1):
        _check(lib.agr_weight_transpose(_code(w), _ptr(w), _ptr(wt), Cout, Cin, k, _stream(w)), "agr_weight_transpose")
    return wt


def conv_dgrad(dy, wt, g):
    """dx (N,Cin,H,W) of the layer `g` from dy (N,Cout,OH,OW) and wt = weight_transpose(w)."""
    lib = _lib.load()
    dx = torch.empty((g.N, g.Cin, g.H, g.W), dtype=dy.dtype, device=dy.device, memory_format=_CL)
    stage = _STAGE.get(conv_path(dy, g, 1), "styleunet_conv_direct")
    stats.add_work(stage, _flops(g))
    with torch.cuda.device(dy.device), stats.stage(stage, launches=1, label=_label("dgrad", g)):
        _check(lib.agr_conv2d_dgrad(_code(dy), C.byref(g), _ptr(dy), _ptr(wt), _ptr(dx), _stream(dy)), "agr_conv2d_dgrad")
    return dx


def conv_dgrad_w(dy, w, g, cin_total=0, cin_offset=0):
    """dx of the layer `g` from dy and the layer's KRSC operand `w` itself.  Tensor-core path: the weight is read in place
    (agr_conv2d_dgrad_krsc); CUDA-core path: through a transposed copy, as before.  `cin_total/cin_offset`: `g.Cin` is a
    slice of the weight's input channels (split contraction)."""
    if conv_path(dy, g, 1) != 1:
        wt = weight_transpose(w)
        return conv_dgrad(dy, wt if not cin_total else wt[cin_offset:cin_offset + g.Cin], g)
    lib = _lib.load()
    dx = torch.empty((g.N, g.Cin, g.H, g.W), dtype=dy.dtype, device=dy.device, memory_format=_CL)
    stats.add_work("styleunet_conv_tc", _flops(g))
    with torch.cuda.device(dy.device), stats.stage("styleunet_conv_tc", launches=1, label=_label("dgrad", g)):
        _check(lib.agr_conv2d_dgrad_krsc(_code(dy), C.byref(g), _ptr(dy), _ptr(w), int(cin_total), int(cin_offset), _ptr(dx), _stream(dy)),
               "agr_conv2d_dgrad_krsc")
    return dx


def conv_wgrad(x, dy, g, dw=None, ci_total=0, ci_offset=0):
    """fp32 weight gradient as a (Cout, k, k, ci_total or Cin) buffer (= KRSC memory).  `dw` given: accumulate into it."""
    lib = _lib.load()
    ct = int(ci_total) if ci_total else g.Cin
    zero = dw is None
    if dw is None:
        dw = torch.empty((g.Cout, g.ksize, g.ksize, ct), dtype=torch.float32, device=x.device)
    stage = _STAGE.get(conv_path(x, g, 2), "styleunet_conv_direct")
    stats.add_work(stage, _flops(g))
    with torch.cuda.device(x.device), stats.stage(stage, launches=1 + int(zero), label=_label("wgrad", g)):
        _check(lib.agr_conv2d_wgrad(_code(x), C.byref(g), _ptr(x), _ptr(dy), _ptr(dw), ct, int(ci_offset), int(zero), _stream(x)),
               "agr_conv2d_wgrad")
    return dw


def _as_kcrs(dw):
    """(Cout,k,k,Cin) buffer -> logical (Cout,Cin,k,k) tensor in channels_last memory (no copy)."""
    return dw.permute(0, 3, 1, 2)


def _act_backward(g, y, nz, activate, has_b, has_n, Cout):
    """Gradient through act(z + noise_w*noise + bias): dz, d_bias, d_noise_w (one pass; reductions by atomics)."""
    lib = _lib.load()
    if not (activate or has_b or has_n):
        return g, None, None
    pixels = g.numel() // Cout
    dz = torch.empty_like(g) if activate else None     # identity activation: dz is g itself, only the reductions run
    db = _zeros(Cout, g.device) if has_b else None
    dn = _zeros(1, g.device) if has_n else None
    with torch.cuda.device(g.device), stats.stage("styleunet_act", launches=1):
        _check(lib.agr_bias_act_backward(_code(g), _ptr(g), _ptr(y), _ptr(dz), pixels, Cout, _ptr(nz) if has_n else None,
                                         nz.numel() if has_n else 1, _ptr(db), _ptr(dn), int(activate), _stream(g)),
               "agr_bias_act_backward")
    return (dz if activate else g), db, dn


class _Conv(torch.autograd.Function):
    """y = act(conv(x, w) + noise_w * noise + bias) for every layer geometry of the path (include/agr_conv.h): forward,
    data gradient and weight gradient on the tcgen05 kernels (bf16, channel counts that are multiples of 64) or the
    CUDA-core kernels (fp32 parity mode, narrow layers).  `w` is the conv-ready KRSC operand; its gradient is produced
    in fp32 and handed to `handle` when the operand comes from a weight plan (see _ModWeightGroup), to `w` otherwise."""

    @staticmethod
    def forward(ctx, x, w, handle, bias, noise, noise_weight, activate, k, stride, pad, transposed):
        x = _nhwc(x)
        w = w.contiguous(memory_format=_CL)
        if w.dtype != x.dtype:
            raise RuntimeError("conv: weight operand %s does not match the activation dtype %s" % (w.dtype, x.dtype))
        g = conv_geom(x.shape, w.shape[0], k, stride, pad, transposed)
        b = bias.detach().float().contiguous() if bias is not None else None
        nz = noise.detach().float().contiguous() if noise is not None else None
        nw = noise_weight.detach().float().contiguous() if noise_weight is not None else None
        y = conv_forward(x, w, g, b, nz, nw, activate)
        ctx.save_for_backward(x, w, y if activate else None, nz)
        ctx.g = g
        ctx.meta = (int(activate), bias is not None, noise is not None and noise_weight is not None,
                    None if bias is None else bias.shape, None if noise_weight is None else noise_weight.shape, handle is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y, nz = ctx.saved_tensors
        activate, has_b, has_n, bshape, nshape, has_handle = ctx.meta
        g = ctx.g
        gy = _nhwc(gy)
        dz, db, dn = _act_backward(gy, y, nz, activate, has_b, has_n, g.Cout)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad_w(dz, w, g)
        if ctx.needs_input_grad[2 if has_handle else 1]:
            dw = _as_kcrs(conv_wgrad(x, dz, g))
        return (dx, None if has_handle else dw, dw if has_handle else None, (db.view(bshape) if has_b else None), None,
                (dn.view(nshape) if has_n else None), None, None, None, None, None)


def conv2d(x, w, handle=None, bias=None, noise=None, noise_weight=None, activate=0, k=None, stride=1, pad=None, transposed=False):
    k = int(w.shape[-1]) if k is None else k
    pad = k // 2 if pad is None else pad
    return _Conv.apply(x, w, handle, bias, noise, noise_weight, int(activate), k, int(stride), int(pad), bool(transposed))


class _SplitConvAct(torch.autograd.Function):
    """act(conv(cat([a (V,Ca,H,W), b (1,Cb,H,W) broadcast]), w) + bias) without the concatenation: the view-independent
    half conv(b, w[:, Ca:]) runs once and enters the per-view half as a residual in the tcgen05 epilogue."""

    @staticmethod
    def forward(ctx, a, b, w, handle, bias, activate):
        a, b = _nhwc(a), _nhwc(b)
        w = w.contiguous(memory_format=_CL)
        Cout, k, Ca, Cb = w.shape[0], w.shape[-1], a.shape[1], b.shape[1]
        bb = bias.detach().float().contiguous() if bias is not None else None
        ga, gb = conv_geom(a.shape, Cout, k), conv_geom(b.shape, Cout, k)
        zb = conv_forward(b, w, gb, out_fp32=True, cin_total=Ca + Cb, cin_offset=Ca)   # fp32 partial sum
        y = conv_forward(a, w, ga, bb, None, None, activate, residual=zb, cin_total=Ca + Cb, cin_offset=0)
        ctx.save_for_backward(a, b, w, y if activate else None)
        ctx.geoms = (ga, gb)
        ctx.meta = (int(activate), k, bias is not None, None if bias is None else bias.shape, handle is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        a, b, w, y = ctx.saved_tensors
        activate, k, has_b, bshape, has_handle = ctx.meta
        ga, gb = ctx.geoms
        Cout, Ca, Cb, V = w.shape[0], a.shape[1], b.shape[1], a.shape[0]
        g = _nhwc(g)
        dz, db, _ = _act_backward(g, y, None, activate, has_b, False, Cout)
        dzs = _new_like(dz[:1], Cout, dz.shape[2], dz.shape[3])
        with torch.cuda.device(g.device), stats.stage("styleunet_act", launches=1):
            _check(lib.agr_sum_batch(_code(dz), _ptr(dz), _ptr(dzs), V, dzs.numel(), _stream(g)), "agr_sum_batch")
        da = conv_dgrad_w(dz, w, ga, Ca + Cb, 0) if ctx.needs_input_grad[0] else None     # w (Cout, Ca+Cb, k, k): slices of its input channels
        dbb = conv_dgrad_w(dzs, w, gb, Ca + Cb, Ca) if ctx.needs_input_grad[1] else None
        dw = None
        if ctx.needs_input_grad[3 if has_handle else 2]:
            dwf = conv_wgrad(a, dz, ga, ci_total=Ca + Cb, ci_offset=0)
            conv_wgrad(b, dzs, gb, dw=dwf, ci_total=Ca + Cb, ci_offset=Ca)
            dw = _as_kcrs(dwf)
        return da, dbb, None if has_handle else dw, dw if has_handle else None, (db.view(bshape) if has_b else None), None


def _operand(weight, s, scale, demodulate, dtype):
    """-> (conv-ready KRSC operand, gradient handle or None)."""
    hit = planned_weight(weight, dtype)
    if hit is not None:
        return hit
    return mod_weight(weight, s, scale, demodulate, dtype)


def equal_conv2d_split(a, b, weight, scale, act_bias=None, activate=True):
    """ConvLayer on cat([a, b], 1) where b (batch 1) is shared by the batch of a; falls back to the plain path when the
    shapes are outside the tensor-core kernel's coverage (the split needs its fp32 residual epilogue)."""
    Ca, Cb = a.shape[1], b.shape[1]
    k = weight.shape[-1]
    ok = (b.shape[0] == 1 and Ca % 64 == 0 and Cb % 64 == 0 and a.dtype == torch.bfloat16 and
          conv_path(a, conv_geom(a.shape, weight.shape[0], k), 0) == 1 and conv_path(b, conv_geom(b.shape, weight.shape[0], k), 0) == 1)
    if not ok:
        return equal_conv2d(torch.cat([a, expand_batch(b, a.shape[0])], 1), weight, scale, 1, k // 2, act_bias, activate)
    w, handle = _operand(weight, None, scale, False, a.dtype)
    return _SplitConvAct.apply(a, b, w, handle, act_bias, int(bool(activate)))


def equal_conv2d(x, weight, scale, stride, padding, act_bias=None, activate=True):
    """EqualConv2d [+ FusedLeakyReLU] (dual_styleunet.py:93-122, 329-371) as one fused op."""
    w, handle = _operand(weight, None, scale, False, x.dtype)
    return conv2d(x, w, handle, bias=act_bias, activate=int(bool(activate)), stride=stride, pad=padding)


def modulated_conv2d(x, weight, s, scale, demodulate=True, upsample=False, downsample=False, blur=None, padding=1,
                     noise=None, noise_weight=None, act_bias=None, activate=True):
    """ModulatedConv2d's fused branch (dual_styleunet.py:256-300) + NoiseInjection + FusedLeakyReLU.  `s` may be a
    callable returning the (1, Cin) style modulation: it is only evaluated when no weight plan holds this layer's operand."""
    hit = planned_weight(weight, x.dtype)
    if hit is None:
        if callable(s):
            s = s()
        if s.shape[0] != 1:
            raise RuntimeError("one style per call: the batch (views of one pose) shares the modulated weight")
        hit = mod_weight(weight, s, scale, demodulate, x.dtype)
    w, handle = hit
    if upsample:      # conv_transpose2d(stride 2) -> Blur -> noise, bias, activation   (dual_styleunet.py:266-282)
        out = blur(conv2d(x, w, handle, stride=2, pad=0, transposed=True))
        return bias_act(out, act_bias, noise=noise, noise_weight=noise_weight, activate=activate)
    if downsample:    # Blur -> conv2d(stride 2)                                          (dual_styleunet.py:283-290)
        x, stride, padding = blur(x), 2, 0
    else:
        stride = 1
    return conv2d(x, w, handle, bias=act_bias, noise=noise, noise_weight=noise_weight, activate=int(bool(activate)),
                  stride=stride, pad=padding)
