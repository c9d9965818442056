#!/usr/bin/env python
"""bench.py — rendered views/sec (fwd+bwd) of the avatar hot path on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): one training step =
ONE pose x 16 ring cameras @1024x1024 of a synthetic 300k-Gaussian avatar:
    3x DualStyleUNet (bf16; position/other nets once, colour net = per-pose prefix + per-view tail)
    -> gather + activations -> fused LBS -> view-batched rasterizer (RGB+depth+alpha) -> loss
    -> backward of all of it -> [one NCCL all-reduce of the flat gradient bucket if N>1] -> fused Adam.
Views are sharded over the N ranks (rank r renders views r, r+N, ...): total work is fixed => "strong".

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = views/s with step inputs resident in HBM; `e2e` = same metric
with the step's inputs coming from pinned HOST buffers (H2D inside the timed region) and the loss read back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

P_GAUSS, IMG, N_VIEWS, J = 300000, 1024, 16, 55
METRIC = "rendered views/sec fwd+bwd @300k Gaussians, 1024x1024, 16 cams"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ product arm
class ProductWorkload:
    def __init__(self, rank, world, device, lr=1e-7, P=P_GAUSS, n_views=N_VIEWS, dtype=torch.bfloat16):
        from animatablegaussians_b200 import avatar, optim, synthetic as S, styleunet_ops as ops
        self.rank, self.world, self.dev = rank, world, device
        self.n_views = n_views
        torch.manual_seed(31359)
        ops.set_compute_dtype(dtype)
        canonical, jnt_mats = avatar.synthetic_canonical(P, size=IMG, J=J)
        self.net = avatar.AvatarNet({"with_viewdirs": True}, canonical=canonical, device=device).to(device)
        self.net.train()
        self.P = self.net.init_points.shape[0]
        # lr: the reference uses 5e-4 (main_avatar.py:45-51).  With the synthetic "sum of outputs" loss that rate walks the
        # random-init nets away from the emulated pre-trained state within a few steps (Gaussians grow until they fill
        # the screen), so the workload would not be stationary.  The Adam arithmetic does not depend on lr.
        self.opt = optim.FlatAdam(self.net.parameters(), lr=lr)
        # owner-computes (parallel.py): a rank holds no gradient for the networks it does not run; the all-reduce brings it
        self.opt.assume_all_active = bool(self.net.net_parallel)
        extrs, Ks = S.ring_cameras(n_views, img=IMG)
        self.views = list(range(rank, n_views, world))
        self.extrs, self.Ks = [extrs[v] for v in self.views], [Ks[v] for v in self.views]
        # step inputs: host (pinned) master copies + resident device copies
        self.h_mats = torch.from_numpy(jnt_mats).pin_memory()
        with torch.no_grad():
            pose = self.net.get_pose_map({"cano2live_jnt_mats_woRoot": self.h_mats.to(device)})
            avatar.emulate_pretrained_heads(self.net, pose[:3])
        self.h_pose = pose.cpu().pin_memory()
        self.d_mats, self.d_pose = self.h_mats.to(device), self.h_pose.to(device)
        self.h_loss = torch.zeros(1).pin_memory()
        # camera block of this rank's views: pinned host master + static device tensors (graph inputs)
        self.views_dev = self.net.prepare_views(self.extrs, self.Ks, IMG, IMG)
        st = self.views_dev["settings"]
        self.h_cam = torch.cat([st.viewmatrix.reshape(len(self.views), 16), st.projmatrix.reshape(len(self.views), 16),
                                st.campos], 1).cpu().pin_memory()
        # ground truth of this rank's views as the cameras store it: uint8 colour + boolean mask + boundary band
        # (main_avatar.py:193-222 reads color_img / mask_img / boundary_mask_img); synthesised from the initial render
        import torch.nn.functional as F
        with torch.no_grad():
            out = self.net.render_views({"smpl_pos_map": self.d_pose, "cano2live_jnt_mats": self.d_mats}, views=self.views_dev)
            m = (out["mask_maps"][..., 0] > 0.5)
            mf = m.float()[:, None]
            band = (F.max_pool2d(mf, 5, 1, 2) - (1.0 - F.max_pool2d(1.0 - mf, 5, 1, 2)))[:, 0] > 0
            gt = (out["rgb_maps"].clamp(0, 1) * 0.7 + 0.15).mul(255.0).round().to(torch.uint8)
        self.h_gt, self.h_mask, self.h_band = gt.cpu().pin_memory(), m.to(torch.uint8).cpu().pin_memory(), band.to(torch.uint8).cpu().pin_memory()
        self.d_gt, self.d_mask, self.d_band = gt.contiguous(), m.to(torch.uint8).contiguous(), band.to(torch.uint8).contiguous()
        self.bg = torch.zeros(3, device=device)
        del out
        self.h2d_bytes = (self.h_mats.numel() * 4 + self.h_pose.numel() * 4 + self.h_cam.numel() * 4 + self.h_gt.numel()
                          + self.h_mask.numel() + self.h_band.numel())
        self.d2h_bytes = 4
        self.graph = None

    def _body(self):
        from animatablegaussians_b200 import styleunet_ops as ops
        with ops.step_arena():
            return self._body_impl()

    def _loss(self, out):
        """Loss head of the trainer on the device (main_avatar.py:193-222: boundary compositing, L1 image, L1 mask; one fused
        kernel, include/agr_loss.h) + a small depth sum so that colour, depth AND alpha all receive gradients (SURVEY.md
        §8d config 4) + the offset regulariser.  Photometric terms are means over this rank's views (weighted by the rank's
        share), the view-independent regulariser is split over the ranks: the summed gradient is the 1-GPU gradient."""
        from animatablegaussians_b200 import loss as L
        photo, _, _ = L.photometric_loss(out["rgb_maps"], out["mask_maps"], self.d_gt, self.d_mask, self.d_band, self.bg, w_l1=1.0, w_mask=0.1)
        return photo * (len(self.views) / float(self.n_views)) + out["depth_maps"].sum() * (1e-3 / (IMG * IMG * self.n_views)) \
            + (0.005 / self.world) * torch.linalg.norm(out["offset"], dim=-1).mean()

    def _body_impl(self):
        items = {"smpl_pos_map": self.d_pose, "cano2live_jnt_mats": self.d_mats}
        loss = self._loss(self.net.render_views(items, return_depth=True, views=self.views_dev))
        loss.backward()
        if self.world > 1:
            self.opt.all_reduce()
        self.opt.step(grad_scale=1.0, zero_grad=True)
        return loss.detach().reshape(1)

    def capture(self):
        """Whole train step (3 U-Nets fwd, LBS, raster fwd, loss, full backward, all-reduce, Adam) as ONE CUDA graph:
        ~9k kernel launches per step replay without per-launch host cost.  Needs the sync-free rasterizer, whose
        fixed instance capacity is taken from one eager (synchronising) step."""
        from animatablegaussians_b200 import rasterizer
        self._body()  # eager: cuDNN autotune, tap caches, and the measured instance count
        torch.cuda.synchronize()
        need = max(rasterizer._capacity_hint.values())
        self.views_dev = dict(self.views_dev, settings=self.views_dev["settings"]._replace(capacity=int(need * 1.2)))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body()
        torch.cuda.synchronize()

    def check_overflow(self):
        from animatablegaussians_b200 import rasterizer
        st = rasterizer.last_device_status
        if st is not None and int(st[1].item()) != 0:
            raise RuntimeError("sync-free rasterizer: binning capacity exceeded (%d instances)" % int(st[0].item()))
        return None if st is None else int(st[0].item())

    def _e2e_setup(self):
        st = self.views_dev["settings"]
        V = len(self.views)
        self._host_inputs = [self.h_mats, self.h_pose, self.h_cam, self.h_gt, self.h_mask, self.h_band]
        self._staging = [torch.empty_like(h, device=self.dev) for h in self._host_inputs]

        def consume():   # staging -> the graph's static input tensors (device-to-device, ~90 MB at HBM rate)
            m, p, c, g, k, b = self._staging
            self.d_mats.copy_(m); self.d_pose.copy_(p)
            st.viewmatrix.copy_(c[:, :16].reshape(V, 4, 4)); st.projmatrix.copy_(c[:, 16:32].reshape(V, 4, 4)); st.campos.copy_(c[:, 32:35])
            self.d_gt.copy_(g); self.d_mask.copy_(k); self.d_band.copy_(b)
        self._consume = consume
        self.copy_stream = torch.cuda.Stream(self.dev)
        self._staged, self._staging_free = torch.cuda.Event(), torch.cuda.Event()
        self._staging_free.record(torch.cuda.current_stream())
        self._upload_pending = False

    def _upload(self):
        """This step's inputs (pose map, joint matrices, cameras, 16 ground-truth images + masks as bytes) from pinned host
        memory into device staging buffers on a copy stream: the upload of step i+1 overlaps the compute of step i."""
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._staging_free)
            for d, h in zip(self._staging, self._host_inputs):
                d.copy_(h, non_blocking=True)
            self._staged.record(self.copy_stream)
        self._upload_pending = True

    def step(self, e2e):
        if e2e:  # this step's inputs come from pinned host memory (one upload per step, pipelined one step ahead)
            if getattr(self, "copy_stream", None) is None:
                self._e2e_setup()
            if not self._upload_pending:
                self._upload()
            main = torch.cuda.current_stream()
            main.wait_event(self._staged)
            self._consume()
            self._staging_free.record(main)
            self._upload()
        if self.graph is not None:
            self.opt.refresh_hyper()       # lr schedule: the captured step re-reads the pinned {lr, grad_scale} pair
            self.graph.replay()
            loss = self.static_loss
        else:
            loss = self._body()
        if e2e:
            self.h_loss.copy_(loss, non_blocking=True)
        return loss


def device_time_ms(fn, steps, world):
    """barrier + sync, time `steps` calls with CUDA events on the current stream, sync; max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(ms.item())


def cpu_baseline_sample(n_views=2):
    """Oracle ("port") timed on the host cores: raster fwd+bwd + LBS of `n_views` of the 16 views."""
    from animatablegaussians_b200 import synthetic as S, camera as C
    from oracle.raster_oracle import RasterOracle
    from oracle import lbs_oracle
    g = S.make_gaussians(P_GAUSS)
    w, mats = S.make_skinning(g["cano"], J=J)
    extrs, Ks = S.ring_cameras(N_VIEWS, img=IMG)
    up = (np.ones((3, IMG, IMG), np.float32), np.ones((1, IMG, IMG), np.float32), np.ones((1, IMG, IMG), np.float32))
    o = RasterOracle()
    t0 = time.perf_counter()
    x = torch.from_numpy(g["xyz"]).requires_grad_(True)
    q = torch.from_numpy(g["rotations"]).requires_grad_(True)
    px, pq = lbs_oracle.transform_cano2live(torch.from_numpy(w), torch.from_numpy(mats), x, q)
    (px.sum() + pq.sum()).backward()
    for v in range(n_views):
        cb = C.camera_block(extrs[v], Ks[v], IMG, IMG)
        o.forward(np.zeros(3, np.float32), px.detach().numpy(), g["rgb"], g["opacity"], g["scales"], pq.detach().numpy(), 1.0, None,
                  cb["viewmatrix"], cb["projmatrix"], cb["tanfovx"], cb["tanfovy"], IMG, IMG)
        o.backward(*up)
    dt = time.perf_counter() - t0
    return {"value": n_views / dt, "unit": "views/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d of 16 views: CPU oracle raster fwd+bwd (oracle/raster_oracle.c, OpenMP) + LBS fwd+bwd "
                      "(oracle/lbs_oracle.py) at 300k/1024^2; StyleUNet NOT included (the reference has no CPU build of "
                      "its CUDA ops and its Python cannot travel to the GPU box)" % n_views}


def run_check(args, rank, world, device):
    """--check: ONE eager train step from the seeded initial state (eval mode: no view-direction noise, so every world size
    sees the same inputs), then the loss, the gradient bucket and the parameter delta of the Adam step as checksums + a
    strided sample.  With --check-against <file of the N=1 run> the N-GPU step must reproduce the 1-GPU step up to fp32
    reduction order: the view-sharded gradient, summed by the one all-reduce, IS the 16-view gradient."""
    import torch.distributed as dist
    wl = ProductWorkload(rank, world, device, lr=1e-3)
    wl.net.eval()
    before = wl.opt.flat_param.clone()
    from animatablegaussians_b200 import styleunet_ops as ops
    with ops.step_arena():
        items = {"smpl_pos_map": wl.d_pose, "cano2live_jnt_mats": wl.d_mats}
        loss = wl._loss(wl.net.render_views(items, return_depth=True, views=wl.views_dev))
        loss.backward()
        if world > 1:
            wl.opt.all_reduce()
        grad = wl.opt.flat_grad.clone()
        wl.opt.step(grad_scale=1.0, zero_grad=True)
    loss = loss.detach().reshape(1).double()
    if world > 1:
        dist.all_reduce(loss)
    torch.cuda.synchronize()
    delta = wl.opt.flat_param - before
    stride = 211
    res = {"n_gpus": world, "loss": float(loss.item()), "grad_l2": float(grad.double().norm()), "grad_sum": float(grad.double().sum()),
           "delta_l2": float(delta.double().norm()), "delta_sum": float(delta.double().sum()), "params": int(wl.opt.numel)}
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        path = os.path.join(ROOT, "gpurun_out", "check_n%d.pt" % world)
        torch.save({"grad": grad[::stride].cpu(), "delta": delta[::stride].cpu(), "res": res}, path)
        if args.check_against:
            ref = torch.load(args.check_against, weights_only=False)
            rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
            res["vs_n1"] = {"loss_rel": abs(res["loss"] - ref["res"]["loss"]) / max(abs(ref["res"]["loss"]), 1e-30),
                            "grad_rel_l2": rel(grad[::stride].cpu(), ref["grad"]),
                            # Adam's first step is -lr * g / (|g| + eps): entries with |g| ~ eps amplify rounding; compare
                            # where the 1-GPU gradient is clearly non-zero
                            "delta_rel_l2": rel(delta[::stride].cpu()[ref["grad"].abs() > 1e-6], ref["delta"][ref["grad"].abs() > 1e-6])}
            # tolerances: fp32 summation order (atomics in the blend backward and in the loss sums — the loss of a rank is a
            # float accumulated by thousands of block-level atomic adds: ~1e-5 relative —, the all-reduce tree) and, through
            # it, a few leaky-ReLU kink / alpha-threshold flips per million gradient entries
            res["vs_n1"]["ok"] = bool(res["vs_n1"]["loss_rel"] < 1e-4 and res["vs_n1"]["grad_rel_l2"] < 2e-3 and res["vs_n1"]["delta_rel_l2"] < 2e-2)
        print(json.dumps({"check": res}))
        sys.stdout.flush()
    if world > 1:
        _exit_multi_rank()
    return 0 if (rank != 0 or not args.check_against or res["vs_n1"]["ok"]) else 1


def run_precision_check(args, device):
    """--check-bf16: the benched configuration (bf16 StyleUNet, 512 -> 1024 maps, 16 views, 300k Gaussians) against the same step
    in fp32 (the parity mode, pinned to the reference by tests/) from the same seeded state: relative L2 of the loss, the
    rendered maps and the gradient bucket per network.  One eager step each, eval mode (no view-direction noise)."""
    from animatablegaussians_b200 import styleunet_ops as ops
    res, keep = {}, {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        wl = ProductWorkload(0, 1, device, lr=1e-3, dtype=dt)
        wl.net.eval()
        with ops.step_arena():
            items = {"smpl_pos_map": wl.d_pose, "cano2live_jnt_mats": wl.d_mats}
            out = wl.net.render_views(items, return_depth=True, views=wl.views_dev)
            loss = wl._loss(out)
            loss.backward()
        wl.opt.gather_grads()
        torch.cuda.synchronize()
        segs = {}
        for pname, p in wl.net.named_parameters():
            if p.requires_grad and p.grad is not None:
                segs.setdefault(pname.split(".")[0], []).append(p.grad.detach().flatten().float())
        keep[name] = {"loss": float(loss), "rgb": out["rgb_maps"].detach().float().cpu(), "mask": out["mask_maps"].detach().float().cpu(),
                      "depth": out["depth_maps"].detach().float().cpu(), "grads": {k: torch.cat(v).cpu() for k, v in segs.items()}}
        del wl, out, loss
        torch.cuda.empty_cache()
    ops.set_compute_dtype(torch.float32)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    a, b = keep["bf16"], keep["fp32"]
    res = {"loss_fp32": b["loss"], "loss_bf16": a["loss"], "loss_rel": abs(a["loss"] - b["loss"]) / abs(b["loss"]),
           "rgb_rel_l2": rel(a["rgb"], b["rgb"]), "mask_rel_l2": rel(a["mask"], b["mask"]), "depth_rel_l2": rel(a["depth"], b["depth"]),
           "grad_rel_l2": {k: rel(a["grads"][k], b["grads"][k]) for k in b["grads"] if k in a["grads"]},
           "config": "BASELINE configs[3]: 300k Gaussians, 16 views @1024^2, maps 1024^2; bf16 StyleUNet vs fp32 StyleUNet, same state"}
    print(json.dumps({"check_bf16": res}))
    return 0


def _time_ms(fn, steps, warmup):
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_aux_config(args, device):
    """BASELINE.json configs other than the headline (1-based): 1 smplx LBS of 10k vertices, 2 single-view forward raster of
    300k Gaussians, 3 full forward (StyleUNet -> LBS -> raster) of 4 views in fp32.  One GPU, one JSON line each."""
    from animatablegaussians_b200 import avatar, camera, synthetic as S, styleunet_ops as ops
    sampler = ClockSampler(device.index or 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs") or 6650.0
    base = {"n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic"}
    if args.config == 2:
        import diff_gaussian_rasterization_depth_alpha as D   # the drop-in surface, as gaussian_renderer.py:14 imports it
        g = S.make_gaussians(P_GAUSS)
        extrs, Ks = S.ring_cameras(N_VIEWS, img=IMG)
        host = {k: torch.from_numpy(g[k]).pin_memory() for k in ("xyz", "opacity", "scales", "rotations", "rgb")}
        dev = {k: v.to(device) for k, v in host.items()}
        rs = camera.make_raster_settings(extrs[0], Ks[0], IMG, IMG, torch.zeros(3, device=device), device)
        rast = D.GaussianRasterizer(rs)
        h_out = torch.zeros(1).pin_memory()

        def step(e2e=False):
            if e2e:
                for k in dev:
                    dev[k].copy_(host[k], non_blocking=True)
            with torch.no_grad():
                color, radii, depth, alpha = rast(means3D=dev["xyz"], means2D=torch.zeros_like(dev["xyz"]), opacities=dev["opacity"],
                                                  colors_precomp=dev["rgb"], scales=dev["scales"], rotations=dev["rotations"])
                if e2e:
                    h_out.copy_(color.sum().reshape(1), non_blocking=True)
        sampler.start()
        ms = _time_ms(step, args.steps, args.warmup)
        ms_e2e = _time_ms(lambda: step(True), args.steps, 3)
        clocks = sampler.stop()
        bytes_fwd = 60.0 * P_GAUSS + 24.0 * IMG * IMG     # SURVEY §8d forward terms: 56 B/Gaussian in + 4 B radii + 24 B/pixel out
        out = dict(base, metric="rendered views/sec, forward raster only @300k Gaussians, 1024x1024, 1 view per call", value=1e3 / ms,
                   unit="views/s", ms_per_step=ms, dtype="fp32",
                   config={"workload": "BASELINE configs[1]: single-view forward raster of 300k posed Gaussians through "
                                       "diff_gaussian_rasterization_depth_alpha.GaussianRasterizer (colours precomputed, one view per call, "
                                       "blocking instance-count read as in the reference API)", "gaussians": P_GAUSS, "image": [IMG, IMG]},
                   e2e={"value": 1e3 / ms_e2e, "unit": "views/s", "h2d_bytes_per_step": sum(v.numel() * 4 for v in host.values()), "d2h_bytes_per_step": 4},
                   gpu_launches=None, clocks=clocks,
                   roofline={"bound": "hbm", "kernel": "forward rasterizer launches of one view", "achieved": bytes_fwd / (ms * 1e-3) / 1e9, "peak": hbm,
                             "unit": "GB/s", "frac": bytes_fwd / (ms * 1e-3) / 1e9 / hbm, "traffic": None, "algorithmic_bytes_per_step": bytes_fwd})
    elif args.config == 3:
        ops.set_compute_dtype(torch.float32)
        torch.manual_seed(31359)
        V = 4
        canonical, mats = avatar.synthetic_canonical(P_GAUSS, size=IMG, J=J)
        net = avatar.AvatarNet({"with_viewdirs": True}, canonical=canonical, device=device).to(device).eval()
        extrs, Ks = S.ring_cameras(N_VIEWS, img=IMG)
        h_mats = torch.from_numpy(mats).pin_memory()
        d_mats = h_mats.to(device)
        with torch.no_grad():
            pose = net.get_pose_map({"cano2live_jnt_mats_woRoot": d_mats})
            avatar.emulate_pretrained_heads(net, pose[:3])
        h_pose = pose.cpu().pin_memory()
        views = net.prepare_views(extrs[:V], Ks[:V], IMG, IMG)
        h_out = torch.zeros(1).pin_memory()

        def step(e2e=False):
            if e2e:
                d_mats.copy_(h_mats, non_blocking=True)
                pose.copy_(h_pose, non_blocking=True)
            with torch.no_grad():
                out = net.render_views({"smpl_pos_map": pose, "cano2live_jnt_mats": d_mats}, views=views)
                if e2e:
                    h_out.copy_(out["rgb_maps"].sum().reshape(1), non_blocking=True)
        sampler.start()
        ms = _time_ms(step, args.steps, args.warmup)
        ms_e2e = _time_ms(lambda: step(True), args.steps, 3)
        clocks = sampler.stop()
        out = dict(base, metric="rendered views/sec, full forward (StyleUNet -> LBS -> raster), 4 views @1024x1024, fp32", value=V * 1e3 / ms,
                   unit="views/s", ms_per_step=ms, dtype="fp32 (exact: CUDA-core convolutions, no TF32)",
                   config={"workload": "BASELINE configs[2]: AvatarNet.render_views, 4 views of one pose, eval mode, fp32 compute",
                           "gaussians": int(net.init_points.shape[0]), "views_per_step": V, "image": [IMG, IMG]},
                   e2e={"value": V * 1e3 / ms_e2e, "unit": "views/s", "h2d_bytes_per_step": h_mats.numel() * 4 + h_pose.numel() * 4, "d2h_bytes_per_step": 4},
                   gpu_launches=None, clocks=clocks,
                   roofline={"bound": "tensor", "kernel": "fp32 parity path: CUDA-core implicit-GEMM convolutions (conv_direct_kernel)", "achieved": (3 * 585.8e9 + (V - 1) * 136.0e9) / (ms * 1e-3) / 1e12,
                             "peak": 80.0, "unit": "TFLOP/s", "frac": (3 * 585.8e9 + (V - 1) * 136.0e9) / (ms * 1e-3) / 1e12 / 80.0, "traffic": None,
                             "peak_source": "nominal fp32 FMA rate of B200 (148 SMs x 128 lanes x 2 x 1.965 GHz ~ 74-80 TFLOP/s): this path does not use tensor cores"})
        ops.set_compute_dtype(torch.float32)
    else:   # config 1
        from animatablegaussians_b200 import smpl_lbs
        rng = np.random.default_rng(0)
        Vn, Jn, NB = 10000, 55, 20
        t = lambda a: torch.from_numpy(np.asarray(a, np.float32)).to(device)
        parents = np.arange(-1, Jn - 1); parents[0] = -1
        model = dict(betas=t(rng.normal(0, 1, (1, NB))), pose=t(rng.normal(0, 0.2, (1, Jn * 3))), v_template=t(rng.normal(0, 0.5, (Vn, 3))),
                     shapedirs=t(rng.normal(0, 0.01, (Vn, 3, NB))), posedirs=t(rng.normal(0, 0.01, ((Jn - 1) * 9, Vn * 3))),
                     J_regressor=t(np.abs(rng.normal(0, 1, (Jn, Vn))) / Vn), parents=torch.from_numpy(parents).to(device),
                     lbs_weights=t(rng.dirichlet(np.ones(Jn) * 0.1, Vn)))
        sampler.start()
        ms = _time_ms(lambda: smpl_lbs.lbs(**model, return_affine_mat=True), args.steps, args.warmup)
        clocks = sampler.stop()
        out = dict(base, metric="smplx.lbs.lbs calls/sec, 10k vertices, 55 joints", value=1e3 / ms, unit="calls/s", ms_per_step=ms, dtype="fp32",
                   config={"workload": "BASELINE configs[0]: SMPL-X linear blend skinning of 10k vertices (plumbing check; the reference runs it on the CPU)"},
                   e2e={"value": 1e3 / ms, "unit": "calls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=None, clocks=clocks,
                   roofline={"bound": "hbm", "kernel": "agr_lbs_points + agr_smpl_joint_chain (latency-bound at this size)", "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None, "traffic": None})
    print(json.dumps(out))
    return 0


def _exit_multi_rank(wl=None):
    """Clean shutdown of a multi-rank run: drop the captured CUDA graph FIRST (it holds the communicator's streams and
    kernels — destroying the process group underneath a live graph is what hung on torch 2.11 / NCCL 2.28), drain the
    device, then destroy the process group.  A watchdog turns a teardown that still hangs into a plain exit after 20 s:
    every collective of the run has completed by then (device_time_ms ends with a barrier + synchronize), the result line is
    already printed, and the driver must never wait on a wedged rank."""
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()

    def bail():
        time.sleep(20.0)
        os._exit(0)
    threading.Thread(target=bail, daemon=True).start()
    try:
        if wl is not None:
            wl.graph = None
            wl.static_loss = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    sys.exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs, 1-based: 4 = the headline train step (default), 5 = 500k Gaussians x 24 views; 1-3: auxiliary lines")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"],
                    help="StyleUNet compute dtype of the product arm: bf16 (tcgen05, the headline) or fp32 (CUDA-core parity kernels) — the "
                         "fp32 line separates the view-batch restructuring from the precision / tensor-core share of the speed-up")
    ap.add_argument("--check", action="store_true", help="one eager step: loss / gradient / parameter-delta checksums (see run_check)")
    ap.add_argument("--check-bf16", action="store_true", help="one eager step in bf16 and in fp32 from the same state: relative L2 of loss / maps / gradients (see run_precision_check)")
    ap.add_argument("--check-against", default=None, help="gpurun_out/check_n1.pt of the 1-GPU --check run to compare with")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        from oracle import reference_arm
        return reference_arm.main(args, rank, world, local)

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cudnn.benchmark = True
    from animatablegaussians_b200 import _lib, stats

    if args.check:
        sys.exit(run_check(args, rank, world, device))
    if args.check_bf16:
        sys.exit(run_precision_check(args, device))
    if args.config in (1, 2, 3):
        if rank == 0:
            run_aux_config(args, device)
        if world > 1:
            _exit_multi_rank()
        return
    n_views, n_gauss = (24, 500000) if args.config == 5 else (N_VIEWS, P_GAUSS)
    wl = ProductWorkload(rank, world, device, P=n_gauss, n_views=n_views, dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    if not args.no_graph:
        wl.capture()
    for _ in range(args.warmup):
        wl.step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = device_time_ms(lambda: wl.step(False), args.steps, world)
    for _ in range(2):
        wl.step(True)
    ms_e2e = device_time_ms(lambda: wl.step(True), args.steps, world)
    instances = wl.check_overflow()
    # per-stage device times + launch count: a separate EAGER pass (event pairs cannot be recorded inside a graph
    # replay and would perturb the timed region anyway)
    graph, wl.graph = wl.graph, None
    wl.net.concurrent_nets = False   # one stream: the stage event pairs must not overlap each other
    for sub in (wl.net.color_net, wl.net.position_net, wl.net.other_net):
        sub.concurrent_decoders = False
    wl.step(False)
    stats.reset()

    def staged_step():
        # Eager launching is host-bound (~1.6k launches per step): without a backlog on the stream an event pair
        # around one launch also measures the host time between the two records.  A spin kernel ahead of the step
        # lets the host run ahead, so the pairs bracket back-to-back kernel executions only.
        try:
            torch.cuda._sleep(240_000_000)   # ~120 ms at 1.97 GHz
        except Exception:
            pass
        wl.step(False)

    device_time_ms(staged_step, args.steps, world)
    st = stats.snapshot()
    wl.graph = graph
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        _exit_multi_rank(wl)

    if os.environ.get("AGR_STAGE_DETAIL"):   # per-layer-geometry device times of the staged pass (stderr)
        for lab, (t, n) in sorted(st["detail"].items(), key=lambda kv: -kv[1][0]):
            print("%-46s n/step %5.1f  ms/step %8.4f" % (lab, n / args.steps, t / args.steps), file=sys.stderr)
    # where the step's kernel time sits with respect to the view shard: convolutions of batch-1 tensors (position / other nets,
    # colour prefix) are REPLICATED on every rank, those of the local view batch are SHARDED (labels carry the batch size)
    conv_rep = sum(t for lab, (t, n) in st["detail"].items() if " N1 " in lab) / args.steps
    conv_shard = sum(t for lab, (t, n) in st["detail"].items() if " N1 " not in lab) / args.steps
    ms_step = ms / args.steps
    value = n_views / (ms_step * 1e-3)
    e2e_value = n_views / (ms_e2e / args.steps * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roof = stats.roofline(st, args.steps, len(wl.views), wl.P, IMG, IMG, peaks)
    roof_tc = stats.roofline_tensor(st, args.steps, peaks)
    out = {
        "metric": METRIC if args.config == 4 else "rendered views/sec fwd+bwd @500k Gaussians, 1024x1024, 24 cams", "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("bf16 (StyleUNet) + fp32 (LBS, rasterizer)" if args.dtype == "bf16" else "fp32 (StyleUNet on the CUDA-core parity kernels, LBS, rasterizer)"),
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[%d]: train step fwd+bwd+Adam (loss head on the device: L1 image + L1 mask vs uint8 ground truth, "
                               "depth term, offset regulariser), synthetic %dk-Gaussian capsule avatar, 1 pose x %d views @1024x1024, "
                               "%s StyleUNet, view-sharded over %d GPU(s)" % (args.config - 1, wl.P // 1000, n_views, args.dtype, world),
                   "gaussians": wl.P, "views_per_step": n_views, "views_per_rank": len(wl.views), "image": [IMG, IMG],
                   "parallelism": ("view-shard x%d + 1 all-reduce" % world) + (", per-pose networks on owner ranks (broadcast / reduce of their outputs)" if wl.net.net_parallel else ""), "cuda_graph": not args.no_graph,
                   "tile_instances_per_step": instances,
                   "l2": "step working set (activations, maps, instance streams: several GB) exceeds the 126 MB L2; no explicit flush",
                   "library_ops": list(__import__("animatablegaussians_b200.styleunet_ops", fromlist=["x"]).LIBRARY_OPS)},
        "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes},
        "gpu_launches": st["launches"], "clocks": clocks, "roofline": roof, "roofline_tensor": roof_tc, "stage_ms_per_step": stats.stage_ms(st, args.steps),
        "shard_split_ms_per_step": {"conv_replicated": conv_rep, "conv_sharded": conv_shard,
                                    "raster_lbs_loss_sharded": sum(stats.stage_ms(st, args.steps).get(k, 0.0) for k in ("raster_fwd", "raster_bwd", "lbs", "loss_head", "avatar_gather")),
                                    "allreduce": stats.stage_ms(st, args.steps).get("allreduce", 0.0), "adam_replicated": stats.stage_ms(st, args.steps).get("adam", 0.0),
                                    "grad_bucket_bytes": int(wl.opt.numel) * 4,
                                    "note": "eager per-call event pairs (rank 0); the act / FIR / weight stages are not split by batch"},
    }
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only
        try:
            out["cpu_baseline"] = cpu_baseline_sample()
        except Exception as e:  # the oracle is the checker; its absence must not hide the GPU number
            out["cpu_baseline"] = {"value": None, "unit": "views/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(out))
    if world > 1:
        _exit_multi_rank(wl)


if __name__ == "__main__":
    main()
