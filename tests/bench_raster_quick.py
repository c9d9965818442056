"""Quick raster-only timing (development aid, not the bench contract): product vs reference kernels on the
300k / 1024^2 synthetic scene, one view at a time and batched."""
import sys, os, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from animatablegaussians_b200 import synthetic as S, camera as C, rasterizer as R
from tests import raster_harness as Hn

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2]

P, img, V = 300000, 1024, 16
g = S.make_gaussians(P)
extrs, Ks = S.ring_cameras(V)
dev = "cuda"
T = lambda a: torch.from_numpy(a).to(dev)
x, o, s, q, c = T(g["xyz"]), T(g["opacity"]), T(g["scales"]), T(g["rotations"]), T(g["rgb"])
bg = torch.zeros(3, device=dev)
up = [torch.randn(3, img, img, device=dev), torch.randn(1, img, img, device=dev), torch.randn(1, img, img, device=dev)]

rs = C.make_raster_settings(extrs[0], Ks[0], img, img, bg, dev)
def prod_fwd():
    return R.rasterize_gaussians(x, torch.zeros_like(x), torch.Tensor([]), c, o, s, q, torch.Tensor([]), rs)
def prod_fwdbwd():
    xx = x.clone().requires_grad_(True); oo = o.clone().requires_grad_(True); ss = s.clone().requires_grad_(True)
    qq = q.clone().requires_grad_(True); cc = c.clone().requires_grad_(True)
    col, rad, dep, alp = R.rasterize_gaussians(xx, torch.zeros_like(xx), torch.Tensor([]), cc, oo, ss, qq, torch.Tensor([]), rs)
    torch.autograd.backward([col, dep, alp], up)
print("product  fwd  ms/view: %.3f" % timeit(prod_fwd))
print("product  f+b  ms/view: %.3f" % timeit(prod_fwdbwd))

from oracle.ref_rasterizer import RefRasterizer
ref = RefRasterizer()
cb = C.camera_block(extrs[0], Ks[0], img, img)
vm, pm, cp = T(cb["viewmatrix"]), T(cb["projmatrix"]), T(cb["campos"])
def ref_fwd():
    return ref.forward(bg, x, c, o, s, q, 1.0, None, vm, pm, cb["tanfovx"], cb["tanfovy"], img, img, campos=cp)
def ref_fwdbwd():
    ref_fwd(); ref.backward(*up)
print("reference fwd ms/view: %.3f" % timeit(ref_fwd))
print("reference f+b ms/view: %.3f" % timeit(ref_fwdbwd))

bs = C.make_batched_settings(extrs, Ks, img, img, bg, dev)
upb = [torch.randn(V, 3, img, img, device=dev), torch.randn(V, 1, img, img, device=dev), torch.randn(V, 1, img, img, device=dev)]
def prod_b_fwd():
    return R.rasterize_gaussians_batched(x, None, None, c, o, s, q, None, bs)
def prod_b_fwdbwd():
    xx = x.clone().requires_grad_(True); oo = o.clone().requires_grad_(True); ss = s.clone().requires_grad_(True)
    qq = q.clone().requires_grad_(True); cc = c.clone().requires_grad_(True)
    col, rad, dep, alp = R.rasterize_gaussians_batched(xx, None, None, cc, oo, ss, qq, None, bs)
    torch.autograd.backward([col, dep, alp], upb)
t = timeit(prod_b_fwd, n=5); print("product batched(16) fwd ms/view: %.3f" % (t / V))
t = timeit(prod_b_fwdbwd, n=5); print("product batched(16) f+b ms/view: %.3f" % (t / V))
