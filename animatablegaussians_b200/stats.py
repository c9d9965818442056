"""Launch counting and per-stage device timing (CUDA events on the launching stream) for bench.py.

The roofline numbers bench.py prints come from here:
  raster   : HBM-bound.  Algorithmic bytes per view (SURVEY.md §8d / BASELINE.md §3):
             B_raster = 184*P + 52*W*H   (fwd: 56 B/Gaussian in + 4 B radii + 24 B/pixel out;
                                          bwd: 56 B re-read + 28 B/pixel in + 68 B/Gaussian gradients out)
  lbs      : HBM-bound.  580*N bytes per pose fwd+bwd (276*N fwd, 304*N bwd).
  styleunet: tensor-bound. Dense-conv FLOPs with the exact per-pose prefix cache:
             3 * (3*585.8 + (v-1)*136.0) GFLOP for v local views (fwd + 2x bwd).
"""
import contextlib

import torch

ENABLED = False
_events = {}
_launches = 0
_work = {}


def reset():
    global ENABLED, _events, _launches, _work
    ENABLED, _events, _launches, _work = True, {}, 0, {}


def add_work(name, amount):
    """Algorithmic work (FLOPs or bytes) of the call being timed under stage `name`."""
    if ENABLED:
        _work[name] = _work.get(name, 0.0) + float(amount)


def count(n):
    global _launches
    if ENABLED:
        _launches += n


@contextlib.contextmanager
def stage(name, launches=0, label=None):
    """Times the enclosed C-ABI call(s) with CUDA events on the current stream; `launches` = number of OUR
    kernels the call launches (library kernels such as CUB are not counted).  `label`: optional finer key (a layer
    geometry) reported under snapshot()["detail"]."""
    if not ENABLED:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    yield
    e1.record()
    _events.setdefault(name, []).append((e0, e1, label))
    count(launches)


def snapshot():
    global ENABLED
    torch.cuda.synchronize()
    detail = {}
    for k, v in _events.items():
        for a, b, lab in v:
            if lab is not None:
                ms, n = detail.get(lab, (0.0, 0))
                detail[lab] = (ms + a.elapsed_time(b), n + 1)
    out = {"launches": _launches, "stages": {k: (sum(a.elapsed_time(b) for a, b, _ in v), len(v)) for k, v in _events.items()},
           "work": dict(_work), "detail": detail}
    ENABLED = False
    return out


def stage_ms(st, steps):
    return {k: round(ms / steps, 4) for k, (ms, n) in st["stages"].items()}


def roofline_tensor(st, steps, peaks):
    """Tensor-pipe roofline of the tcgen05 implicit-GEMM convolution launches of one step."""
    peak = peaks.get("bf16_tflops_sustained")
    which = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
    if not peak:
        peak, which = 1400.0, "fallback (B200_PROFILING.md sustained)"
    ms, n = st["stages"].get("styleunet_conv_tc", (0.0, 0))
    fl = st.get("work", {}).get("styleunet_conv_tc", 0.0)
    ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc2_kernel (CTA pairs) / conv_tc3_kernel (persistent) + conv_wgrad_tc_kernel (tcgen05.mma implicit-GEMM convolutions: forward, data gradient and weight gradient launches of every layer with Cin, Cout multiples of 64)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": which,
            "launches_per_step": n / max(steps, 1), "algorithmic_flop_per_step": fl / max(steps, 1), "seconds_per_step": ms * 1e-3 / max(steps, 1)}


def roofline(st, steps, views_local, P, W, H, peaks):
    hbm = peaks.get("hbm_gbs")
    which = "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    if not hbm:
        hbm, which = 6650.0, "fallback (B200_PROFILING.md)"
    s = st["stages"]
    t_raster = (s.get("raster_fwd", (0, 0))[0] + s.get("raster_bwd", (0, 0))[0]) / steps * 1e-3
    bytes_step = (184.0 * P + 52.0 * W * H) * views_local
    achieved = bytes_step / t_raster / 1e9 if t_raster > 0 else 0.0
    return {"bound": "hbm", "kernel": "rasterizer fwd+bwd launches of one step (preprocess, duplicate, CUB scan/sort, "
            "gather, blend fwd, blend bwd, preprocess bwd)", "achieved": achieved, "peak": hbm, "unit": "GB/s",
            "frac": achieved / hbm,
            # dram__bytes_read.sum + dram__bytes_write.sum of the two blend kernels (the stage's dominant launches) for a
            # 16-view step, from profiles/r01_ncu_full_blend_{fwd,bwd}.txt (ncu --set full, scene of tools/prof_raster.py)
            "traffic": (506.76e6 + 375.75e6 + 774.91e6 + 196.65e6) * views_local / 16.0, "peak_source": which,
            "algorithmic_bytes_per_step": bytes_step, "seconds_per_step": t_raster}
