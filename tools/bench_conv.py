"""Device-time table of the convolution entry points (include/agr_conv.h) for the layer geometries of the benched step,
per generation of the forward-form tcgen05 kernel (agr_conv2d_set_generation).  No checker here (tools/try_conv.py is).
    timeout 600 python tools/bench_conv.py [gens, e.g. 1,3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animatablegaussians_b200 import _lib, styleunet_ops as ops  # noqa: E402

CL = torch.channels_last
SHAPES = [  # N, H, W, Cin, Cout, k, stride, pad, transposed
    (16, 512, 512, 64, 64, 3, 1, 1, False), (16, 256, 256, 128, 64, 3, 2, 0, True), (16, 256, 256, 128, 128, 3, 1, 1, False),
    (1, 512, 512, 64, 64, 3, 1, 1, False), (1, 256, 256, 128, 128, 3, 1, 1, False), (1, 256, 256, 256, 128, 3, 1, 1, False),
    (1, 128, 128, 256, 256, 3, 1, 1, False), (1, 128, 128, 512, 256, 3, 1, 1, False), (1, 64, 64, 512, 512, 3, 1, 1, False),
    (1, 64, 64, 1024, 512, 3, 1, 1, False), (1, 32, 32, 512, 512, 3, 1, 1, False), (1, 32, 32, 1024, 512, 3, 1, 1, False),
    (1, 16, 16, 512, 512, 3, 1, 1, False), (1, 8, 8, 512, 512, 3, 1, 1, False),
    (1, 64, 64, 512, 256, 3, 2, 0, True), (1, 128, 128, 256, 128, 3, 2, 0, True), (1, 256, 256, 128, 64, 3, 2, 0, True),
    (1, 257, 257, 128, 256, 3, 2, 0, False), (1, 129, 129, 256, 512, 3, 2, 0, False), (1, 65, 65, 512, 512, 3, 2, 0, False),
    (32, 256, 256, 64, 128, 4, 2, 1, False),
]


def dev_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    gens = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,2,3,0").split(",")]
    lib = _lib.load()
    wg = [tuple(int(u) for u in v.split(":")) for v in os.environ.get("WGRAD_SPLIT", "296:16").split(",")]   # max_ctas:min_boxes
    print("%-44s" % "geometry" + "".join("  fwd g%d  TF/s | dgrad g%d  TF/s |" % (g, g) for g in gens) + "".join("  wg@%d:%d TF/s |" % c for c in wg))
    tot = {g: [0.0, 0.0] for g in gens}
    totw = {}
    for (N, H, W, Cin, Cout, k, st, pad, T) in SHAPES:
        x = torch.randn(N, Cin, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=CL)
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).to(torch.bfloat16).contiguous(memory_format=CL)
        g = ops.conv_geom(x.shape, Cout, k, st, pad, T)
        dy = torch.randn(N, Cout, g.OH, g.OW, device="cuda").to(torch.bfloat16).contiguous(memory_format=CL)
        wt = ops.weight_transpose(w)
        fl = ops._flops(g)
        line = "%-44s" % ops._label("", g)
        for gen in gens:
            lib.agr_conv2d_set_generation(gen)
            a = dev_ms(lambda: ops.conv_forward(x, w, g))
            b = dev_ms(lambda: ops.conv_dgrad_w(dy, w, g))
            tot[gen][0] += a; tot[gen][1] += b
            line += " %7.3f %5.0f | %8.3f %5.0f |" % (a, fl / a / 1e9, b, fl / b / 1e9)
        for c_ in wg:
            lib.agr_conv2d_set_wgrad_split(*c_)
            c = dev_ms(lambda: ops.conv_wgrad(x, dy, g))
            totw[c_] = totw.get(c_, 0.0) + c
            line += " %7.3f %5.0f |" % (c, fl / c / 1e9)
        lib.agr_conv2d_set_wgrad_split(296, 16)
        print(line, flush=True)
        del x, w, dy, wt
    print("sum ms: " + "  ".join("g%d fwd %.3f dgrad %.3f" % (g, tot[g][0], tot[g][1]) for g in gens) + "  wgrad " + str({k: round(v, 3) for k, v in totw.items()}))
    lib.agr_conv2d_set_generation(0)


if __name__ == "__main__":
    main()
