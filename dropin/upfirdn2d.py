"""Drop-in for the reference's native module `upfirdn2d` (network/styleunet/upfirdn2d.cpp:17-30, loaded by
network/styleunet/upfirdn2d.py:30 `import upfirdn2d as upfirdn2d_op`): same function and argument order, on
libagr_b200.so (agr_upfirdn2d, include/agr_styleunet.h).  input is (major, in_h, in_w, minor) — NHWC with N = major,
C = minor — exactly the layout agr_upfirdn2d takes."""
import ctypes as C

import torch

from animatablegaussians_b200 import _lib
from animatablegaussians_b200 import styleunet_ops as _ops   # registers agr_upfirdn2d; host-tap cache


def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    if not input.is_cuda or not kernel.is_cuda:
        raise RuntimeError("input and kernel must be CUDA tensors")       # upfirdn2d.cpp:12-14
    if not input.is_contiguous():
        raise RuntimeError("input must be contiguous")
    if up_x != up_y or down_x != down_y:
        raise RuntimeError("upfirdn2d drop-in: equal x / y factors only (all the reference's call sites)")
    major, in_h, in_w, minor = input.shape
    kh, kw = kernel.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y      # upfirdn2d_kernel.cu:294-297
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    taps, _, _ = _ops._host_taps(kernel, True)                            # the op correlates with the flipped kernel (:157-163)
    out = torch.empty((major, out_h, out_w, minor), dtype=input.dtype, device=input.device)
    lib = _lib.load()
    with torch.cuda.device(input.device):
        st = lib.agr_upfirdn2d(_ops._code(input), C.c_void_p(input.data_ptr()), C.c_void_p(out.data_ptr()), major, in_h, in_w, minor,
                               out_h, out_w, C.cast(taps, C.POINTER(C.c_float)), kh, kw, int(up_x), int(down_x), int(pad_x0), int(pad_y0),
                               C.c_void_p(torch.cuda.current_stream(input.device).cuda_stream))
    if st != _lib.AGR_OK:
        raise RuntimeError("agr_upfirdn2d failed (status %d)" % st)
    return out
