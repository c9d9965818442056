"""Profiling driver for the tcgen05 implicit-GEMM conv: a few representative layer shapes of the DualStyleUNet
(SURVEY.md Appendix A), timed with CUDA events; also the target of `ncu --set full -k regex:conv_tc`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animatablegaussians_b200 import styleunet_ops as ops

SHAPES = [  # (N, H, W, Cin, Cout, k)
    (1, 64, 64, 1024, 512, 3), (1, 64, 64, 512, 512, 3), (1, 128, 128, 512, 256, 3), (1, 128, 128, 256, 256, 3),
    (1, 256, 256, 256, 128, 3), (1, 256, 256, 128, 128, 3), (1, 512, 512, 64, 64, 3), (16, 256, 256, 256, 128, 3),
    (16, 512, 512, 64, 64, 3),
]
peak = 1719.8
for (N, H, W, Cin, Cout, k) in SHAPES:
    x = torch.randn(N, Cin, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.zeros(Cout, device="cuda")
    for _ in range(3):
        y = ops._tc_conv(x, w, Cout, k, b, None, None, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        y = ops._tc_conv(x, w, Cout, k, b, None, None, True)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps
    fl = 2.0 * N * H * W * Cin * Cout * k * k
    print("N=%2d %4dx%-4d %4d->%-4d k%d : %8.1f us  %7.1f TFLOP/s  (%.1f%% of measured bf16 peak %.0f)" % (N, H, W, Cin, Cout, k, us, fl / us / 1e6, 100 * fl / us / 1e6 / peak, peak))
