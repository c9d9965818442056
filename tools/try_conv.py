"""Hardware check of include/agr_conv.h: forward / data gradient / weight gradient of every layer geometry of the path,
both paths (tcgen05 bf16, direct fp32 / bf16), against torch's fp32 convolution (the CHECKER here, not the product).
    timeout 300 python tools/try_conv.py [group]        group in {same, down, up, view, narrow, big, all}
Run under `timeout`: a wrong descriptor can hang the mbarrier pipeline."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animatablegaussians_b200 import styleunet_ops as ops  # noqa: E402

CL = torch.channels_last
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def ref_conv(x, w, k, stride, pad, transposed):
    if transposed:
        return F.conv_transpose2d(x, w.transpose(0, 1), stride=stride, padding=pad)   # w (Cout,Cin,k,k) -> (Cin,Cout,k,k)
    return F.conv2d(x, w, stride=stride, padding=pad)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))


def case(N, H, W, Cin, Cout, k, stride, pad, transposed, dtype, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(N * 7 + H + Cin * 3 + Cout + k)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).to(dtype).contiguous(memory_format=CL)
    w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).to(dtype).contiguous(memory_format=CL)
    geom = ops.conv_geom(x.shape, Cout, k, stride, pad, transposed)
    paths = [ops.conv_path(x, geom, i) for i in range(3)]
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    yr = ref_conv(xr, wr, k, stride, pad, transposed)
    y = ops.conv_forward(x, w, geom)
    dy = torch.randn(yr.shape, device="cuda", generator=g).to(dtype).contiguous(memory_format=CL)
    yr.backward(dy.float())
    dx = ops.conv_dgrad(dy, ops.weight_transpose(w), geom)
    dx2 = ops.conv_dgrad_w(dy, w, geom)            # tensor-core path: straight from the KRSC operand (MN-major B)
    if Cin % 128 == 0 and paths[1] == 1:           # ... and a slice of the weight's input channels (split contraction)
        gh = ops.conv_geom((N, Cin // 2, H, W), Cout, k, stride, pad, transposed)
        dxh = ops.conv_dgrad_w(dy, w, gh, Cin, Cin // 2)
    else:
        dxh = None
    dw = ops._as_kcrs(ops.conv_wgrad(x, dy, geom))
    torch.cuda.synchronize()
    e = (rel(y, yr), max(rel(dx, xr.grad), rel(dx2, xr.grad), 0.0 if dxh is None else rel(dxh, xr.grad[:, Cin // 2:])), rel(dw, wr.grad))
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    msg = "%-4s N%-2d %3dx%-3d %4d->%-4d k%d s%d p%d %s paths %s  fwd %.1e dgrad %.1e wgrad %.1e" % (
        "bf16" if dtype == torch.bfloat16 else "fp32", N, H, W, Cin, Cout, k, stride, pad, "T" if transposed else " ", paths, *e)
    if time_it:
        fl = ops._flops(geom)
        for name, fn in (("fwd", lambda: ops.conv_forward(x, w, geom)), ("dgrad", lambda: ops.conv_dgrad_w(dy, w, geom)),
                         ("wgrad", lambda: ops.conv_wgrad(x, dy, geom))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 100
            msg += "  %s %.3f ms %.0f TF/s" % (name, ms, fl / ms / 1e9)
    ok = all(v <= tol for v in e)
    print(msg + ("" if ok else "   <-- FAIL"), flush=True)
    return ok


GROUPS = {
    "same": [(1, 8, 16, 64, 128, 1, 1, 0, False), (1, 16, 16, 64, 64, 3, 1, 1, False), (2, 32, 32, 128, 128, 3, 1, 1, False),
             (1, 8, 8, 512, 512, 3, 1, 1, False), (1, 64, 64, 512, 256, 3, 1, 1, False), (3, 24, 40, 64, 192, 3, 1, 1, False)],
    "down": [(1, 33, 33, 128, 256, 3, 2, 0, False), (1, 17, 17, 512, 512, 3, 2, 0, False), (2, 65, 65, 64, 64, 3, 2, 0, False),
             (1, 129, 129, 256, 512, 3, 2, 0, False)],
    "up": [(1, 8, 8, 512, 512, 3, 2, 0, True), (1, 16, 16, 512, 512, 3, 2, 0, True), (2, 32, 32, 128, 64, 3, 2, 0, True),
           (1, 64, 64, 512, 256, 3, 2, 0, True)],
    "view": [(4, 64, 64, 64, 128, 4, 2, 1, False), (4, 128, 128, 1, 64, 4, 2, 1, False)],
    "narrow": [(1, 65, 65, 3, 128, 3, 2, 0, False), (1, 32, 32, 3, 128, 1, 1, 0, False), (2, 32, 32, 64, 12, 1, 1, 0, False),
               (1, 16, 16, 512, 32, 1, 1, 0, False)],
    "big": [(16, 512, 512, 64, 64, 3, 1, 1, False), (16, 256, 256, 128, 64, 3, 2, 0, True), (16, 256, 256, 256, 128, 3, 1, 1, False),
            (1, 129, 129, 256, 512, 3, 2, 0, False), (1, 64, 64, 1024, 512, 3, 1, 1, False), (16, 512, 512, 64, 12, 1, 1, 0, False),
            (32, 256, 256, 64, 128, 4, 2, 1, False), (1, 512, 512, 64, 64, 3, 1, 1, False), (3, 200, 328, 64, 64, 3, 1, 1, False),
            (2, 200, 168, 128, 64, 3, 2, 0, True), (1, 256, 256, 256, 128, 3, 1, 1, False), (1, 128, 128, 512, 256, 3, 1, 1, False)],
}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if len(sys.argv) > 2:     # generation of the forward-form tcgen05 kernel (include/agr_conv.h agr_conv2d_set_generation)
        from animatablegaussians_b200 import _lib
        print("generation", _lib.load().agr_conv2d_set_generation(int(sys.argv[2])), flush=True)
    ok = True
    for name, cases in GROUPS.items():
        if which not in ("all", name):
            continue
        print("== %s" % name, flush=True)
        for c in cases:
            if name == "big":
                ok &= case(*c, torch.bfloat16, time_it=True)
                continue
            ok &= case(*c, torch.bfloat16)
            if c[0] * c[1] * c[2] * c[3] * c[4] <= 2 * 64 * 64 * 512 * 256:
                ok &= case(*c, torch.float32)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)
