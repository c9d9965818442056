"""Single-kernel drivers for `ncu --set full` captures (one launch of each kernel of interest inside a
cudaProfilerStart/Stop bracket, after a warm-up).   python tools/ncu_targets.py conv64|conv64p|conv128|wgrad64|lbs|raster"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animatablegaussians_b200 import _lib, styleunet_ops as ops  # noqa: E402

CL = torch.channels_last
what = sys.argv[1]
torch.cuda.set_device(0)
lib = _lib.load()
start, stop = torch.cuda.cudart().cudaProfilerStart, torch.cuda.cudart().cudaProfilerStop


def conv_case(N, H, W, Cin, Cout, k=3, stride=1, pad=1, T=False):
    x = torch.randn(N, Cin, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=CL)
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).to(torch.bfloat16).contiguous(memory_format=CL)
    g = ops.conv_geom(x.shape, Cout, k, stride, pad, T)
    dy = torch.randn(N, Cout, g.OH, g.OW, device="cuda").to(torch.bfloat16).contiguous(memory_format=CL)
    return x, w, g, dy


if what in ("conv64", "conv64p", "conv128"):
    if what == "conv64p":
        os.environ["AGR_CONV_PERSISTENT"] = "1"
    lib.agr_conv2d_set_generation(0 if what == "conv64p" else 1)
    x, w, g, dy = conv_case(16, 512, 512, 64, 64) if what != "conv128" else conv_case(16, 256, 256, 128, 128)
    for _ in range(2):
        ops.conv_forward(x, w, g)
    torch.cuda.synchronize()
    start(); ops.conv_forward(x, w, g); torch.cuda.synchronize(); stop()
elif what == "wgrad64":
    x, w, g, dy = conv_case(16, 512, 512, 64, 64)
    for _ in range(2):
        ops.conv_wgrad(x, dy, g)
    torch.cuda.synchronize()
    start(); ops.conv_wgrad(x, dy, g); torch.cuda.synchronize(); stop()
elif what == "lbs":
    from animatablegaussians_b200 import lbs, synthetic as S
    g = S.make_gaussians(300000)
    wts, mats = S.make_skinning(g["cano"], J=55)
    T = lambda a: torch.from_numpy(a).cuda()
    x, q = T(g["xyz"]).requires_grad_(True), T(g["rotations"]).requires_grad_(True)
    wt, mt = T(wts), T(mats)
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize(); start()
        px, pq = lbs.transform_cano2live(wt, mt, x, q)
        (px.sum() + pq.sum()).backward()
    torch.cuda.synchronize(); stop()
elif what == "raster":
    from animatablegaussians_b200 import synthetic as S, camera as C, rasterizer as R
    V, P, img = 16, 300000, 1024
    g = S.make_gaussians(P)
    extrs, Ks = S.ring_cameras(V)
    T = lambda a: torch.from_numpy(a).cuda()
    x, o, s, q, c = (T(g[k]).requires_grad_(True) for k in ("xyz", "opacity", "scales", "rotations", "rgb"))
    bs = C.make_batched_settings(extrs, Ks, img, img, torch.zeros(3, device="cuda"), "cuda")
    up = [torch.randn(V, 3, img, img, device="cuda"), torch.randn(V, 1, img, img, device="cuda"), torch.randn(V, 1, img, img, device="cuda")]
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize(); start()
        col, rad, dep, alp = R.rasterize_gaussians_batched(x, None, None, c, o, s, q, None, bs)
        torch.autograd.backward([col, dep, alp], up)
    torch.cuda.synchronize(); stop()
print("done", what)
