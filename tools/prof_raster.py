"""Profiling driver: view-batched raster fwd+bwd on the 300k / 1024^2 / 16-view synthetic scene.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/raster_launches.csv python tools/prof_raster.py 2"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animatablegaussians_b200 import synthetic as S, camera as C, rasterizer as R

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
V = int(sys.argv[2]) if len(sys.argv) > 2 else 16
P, img = 300000, 1024
g = S.make_gaussians(P)
extrs, Ks = S.ring_cameras(V)
dev = "cuda"
T = lambda a: torch.from_numpy(a).to(dev)
x, o, s, q, c = (T(g[k]).requires_grad_(True) for k in ("xyz", "opacity", "scales", "rotations", "rgb"))
bg = torch.zeros(3, device=dev)
bs = C.make_batched_settings(extrs, Ks, img, img, bg, dev)
up = [torch.randn(V, 3, img, img, device=dev), torch.randn(V, 1, img, img, device=dev), torch.randn(V, 1, img, img, device=dev)]
for it in range(iters):
    col, rad, dep, alp = R.rasterize_gaussians_batched(x, None, None, c, o, s, q, None, bs)
    torch.autograd.backward([col, dep, alp], up)
torch.cuda.synchronize()
print("done", float(col.sum()))
