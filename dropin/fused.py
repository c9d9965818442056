"""Drop-in for the reference's native module `fused` (network/styleunet/fused_bias_act.cpp:18-31, loaded by
network/styleunet/fused_act.py:30 `import fused`): same function, same argument order and checks, on libagr_b200.so
(agr_fused_bias_act, include/agr_styleunet.h).  Put this directory on sys.path in place of the reference's built
extension and the reference's fused_act.py runs unchanged."""
import ctypes as C

import torch

from animatablegaussians_b200 import _lib

_p = C.c_void_p
_lib.register_symbols({"agr_fused_bias_act": (C.c_int, [C.c_int32, _p, _p, _p, _p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                                        C.c_float, C.c_float, _p])})


def _check(t, name):
    # fused_bias_act.cpp:14-16 CHECK_CUDA / CHECK_CONTIGUOUS
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    _check(input, "input")
    _check(bias, "bias")
    code = {torch.float32: 0, torch.bfloat16: 1}.get(input.dtype)
    if code is None:
        raise RuntimeError("fused_bias_act: fp32 or bf16 tensors (got %s)" % input.dtype)
    x = input
    out = torch.empty_like(x)
    use_bias, use_ref = bias.numel() > 0, refer.numel() > 0
    step_b = 1
    for i in range(2, x.dim()):
        step_b *= x.size(i)
    size_b = x.size(1) if x.dim() > 1 else 1
    if use_bias and (bias.dtype != x.dtype or bias.numel() != size_b):
        raise RuntimeError("fused_bias_act: bias must have input.size(1) elements of the input dtype")
    if use_ref:
        _check(refer, "refer")
        if refer.dtype != x.dtype or refer.numel() != x.numel():
            raise RuntimeError("fused_bias_act: refer must match input")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.agr_fused_bias_act(code, _p(x.data_ptr()), _p(bias.data_ptr()) if use_bias else None,
                                    _p(refer.data_ptr()) if use_ref else None, _p(out.data_ptr()), x.numel(), step_b, size_b,
                                    int(act), int(grad), float(alpha), float(scale),
                                    _p(torch.cuda.current_stream(x.device).cuda_stream))
    if st != _lib.AGR_OK:
        raise RuntimeError("agr_fused_bias_act failed (status %d)" % st)
    return out
