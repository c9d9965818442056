"""Per-call inventory of every dense contraction in one eager train step of the bench workload: which path it took
(tcgen05 kernel, split contraction, or a library call through torch) with shapes and device time (CUDA events; eager, so
times include no launch gaps only when the stream is backlogged — compare shares, not absolutes).
    python tools/conv_inventory.py > gpurun_out/conv_inventory.txt"""
import collections
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from animatablegaussians_b200 import styleunet_ops as ops  # noqa: E402

records = []


def timed(tag, describe, fn):
    def wrapper(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        records.append((tag, describe(*a, **kw), e0, e1))
        return out
    return wrapper


def d_conv(x, w, *a, **kw):
    return "x%s w%s stride=%s %s" % (tuple(x.shape), tuple(w.shape), kw.get("stride", a[1] if len(a) > 1 else 1), x.dtype)


F.conv2d = timed("lib conv2d", d_conv, F.conv2d)
F.conv_transpose2d = timed("lib conv_transpose2d", d_conv, F.conv_transpose2d)
torch.ops.aten.convolution_backward = timed(   # the explicit dgrad / wgrad calls of _ConvAct / _SplitConvAct
    "lib convolution_backward", lambda dz, x, w, *a, **kw: "dz%s x%s w%s out_mask=%s" % (tuple(dz.shape), tuple(x.shape), tuple(w.shape), a[-1]),
    torch.ops.aten.convolution_backward)
ops._tc_conv = timed("tcgen05", lambda x, w, Cout, k, *a, **kw: "x%s Cout=%d k=%d" % (tuple(x.shape), Cout, k), ops._tc_conv)
ops._tc_conv_split = timed("tcgen05 split", lambda x, w, Cout, k, *a, **kw: "x%s Cout=%d k=%d" % (tuple(x.shape), Cout, k), ops._tc_conv_split)

torch.cuda.set_device(0)
torch.backends.cudnn.benchmark = True
wl = bench.ProductWorkload(0, 1, torch.device("cuda", 0))
wl.step(False)
torch.cuda.synchronize()
records.clear()
torch.cuda._sleep(400_000_000)   # backlog the stream so the event pairs bracket kernels, not host gaps
wl.step(False)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for tag, desc, e0, e1 in records:
    k = (tag, desc)
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + e0.elapsed_time(e1))
tot = sum(t for _, t in agg.values())
print("%-22s %-78s %4s %9s %6s" % ("path", "call", "n", "ms", "%"))
for (tag, desc), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-22s %-78s %4d %9.3f %6.1f" % (tag, desc, n, t, 100 * t / tot))
print("total %.3f ms over %d calls (autograd-internal cuDNN backward calls of the library convs are not visible here)" % (tot, len(records)))
