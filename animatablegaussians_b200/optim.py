"""Flat-bucket Adam: all trainable parameters of the avatar live in ONE fp32 buffer (and their gradients in
another), so that (a) the view-sharded step needs exactly one NCCL all-reduce and (b) the optimizer is one
fused streaming kernel (include/agr_optim.h).  Mirrors torch.optim.Adam(lr) + step() + zero_grad() of the
reference trainer (main_avatar.py:49-51,255-256).

Gradients reach the bucket lazily: after `step()` / `zero_grad()` every `p.grad` is None (torch's set_to_none default),
so autograd hands each parameter its freshly produced gradient tensor without launching an accumulation kernel
(r01 profile: ~800 `p.grad += g` launches per step otherwise); `flat_grad` / `all_reduce()` / `step()` first gather
those tensors into the bucket with one multi-tensor copy and re-point `p.grad` at the bucket views."""
import ctypes as C

import torch

from . import _lib, stats

_p = C.c_void_p
_lib.register_symbols({
    "agr_adam_step": (C.c_int, [C.c_int64, _p, _p, _p, _p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                C.c_float, C.c_int32, _p]),
    "agr_adam_step_graph": (C.c_int, [C.c_int64, _p, _p, _p, _p, C.c_float, C.c_float, C.c_float, C.c_float, _p,
                                      C.c_float, C.c_int32, _p]),
    "agr_adam_step_segments": (C.c_int, [C.c_int64, _p, _p, _p, _p, C.c_int32, _p, _p, _p, _p, _p, C.c_float, C.c_float, C.c_float,
                                         C.c_int32, _p]),
})


def cosine_lr(lr_init, iter_idx, iter_num, alpha=0.05):
    """The trainer's schedule (main_avatar.py:61-68): cosine from lr_init down to alpha * lr_init over iter_num."""
    import math
    progress = iter_idx / iter_num
    return lr_init * ((math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha)


CHUNK = 256   # elements per chunk of the segmented kernel (include/agr_optim.h)


class FlatAdam:
    """torch.optim.Adam semantics over one flat bucket:
      * a parameter whose `.grad` is None at step time is skipped entirely (no moment decay, no step increment) — the
        trainer toggles requires_grad per iteration (main_avatar.py:184-189) and pretraining never reaches some nets;
      * every parameter has its own step counter (bias correction), kept on the device;
      * lr and grad_scale are read from device memory, refreshed from a pinned host pair by an 8-byte copy that is part of
        the step — a captured CUDA graph therefore follows `param_groups[0]['lr']` (update_lr, main_avatar.py:61-68):
        call refresh_hyper() before each replay."""

    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        self._all_params = list(params)           # positions = torch.optim.Adam's parameter indices (checkpoints)
        self.params = [p for p in self._all_params if p.requires_grad]
        dev = self.params[0].device
        # every parameter = one segment, padded to whole 256-element chunks (16-byte aligned views, one segment per chunk)
        offs, total, chunk_seg = [], 0, []
        for i, p in enumerate(self.params):
            offs.append(total)
            nchunk = (p.numel() + CHUNK - 1) // CHUNK
            chunk_seg += [i] * nchunk
            total += nchunk * CHUNK
        self.numel = sum(p.numel() for p in self.params)
        S = len(self.params)
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self._flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32, device=dev)
        self._seg_step = torch.zeros(S, dtype=torch.int32, device=dev)
        self._seg_corr = torch.zeros(S, 2, dtype=torch.float32, device=dev)
        pin = dev.type == "cuda"
        # True: parameters without a gradient on THIS rank still step — for the owner-computes multi-GPU scheme
        # (parallel.py), where the all-reduce delivers the owner's gradient into the (locally zero) bucket segment
        self.assume_all_active = False
        self._h_active = torch.ones(S, dtype=torch.int32).pin_memory() if pin else torch.ones(S, dtype=torch.int32)
        self._seg_active = torch.ones(S, dtype=torch.int32, device=dev)
        self._h_hyper = torch.tensor([lr, 1.0], dtype=torch.float32)
        if pin:
            self._h_hyper = self._h_hyper.pin_memory()
        self._d_hyper = self._h_hyper.to(dev)
        self._last_hyper = (float(lr), 1.0)
        self._grad_views = []
        for p, o in zip(self.params, offs):
            k = p.numel()
            self.flat_param[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + k].view_as(p.data)
            self._grad_views.append(self._flat_grad[o:o + k].view_as(p.data))
            p.grad = None
        self._bucket_clean = True   # bucket is all-zero: gathered gradients can be copied instead of added
        # one group, torch.optim layout: the trainer's update_lr() writes param_groups[0]['lr'] (main_avatar.py:61-68)
        self.param_groups = [dict(params=self._all_params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)]
        self._offsets = offs

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = v

    @property
    def betas(self):
        return self.param_groups[0]["betas"]

    @property
    def eps(self):
        return self.param_groups[0]["eps"]

    @property
    def t(self):
        """Largest per-parameter step (host read; for logging / tests)."""
        return int(self._seg_step.max().item())

    def refresh_hyper(self, grad_scale=None):
        """Write lr (param_groups[0]['lr']) [and grad_scale] into the pinned pair the step's 8-byte H2D copy reads.  Eager
        steps do this themselves; call it before replaying a graph captured around step().  When a value CHANGES the call
        first waits for the device to drain: an earlier replay still in flight must read the old pair."""
        lr = float(self.lr)
        gs = self._last_hyper[1] if grad_scale is None else float(grad_scale)
        if (lr, gs) != self._last_hyper:
            if self.flat_param.is_cuda and not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(self.flat_param.device).synchronize()
            self._h_hyper[0] = lr
            self._h_hyper[1] = gs
            self._last_hyper = (lr, gs)

    def state_dict(self):
        """torch.optim.Adam's checkpoint layout (what the trainer stores in optm.pt, main_avatar.py:790-795): per-parameter
        'step' / 'exp_avg' / 'exp_avg_sq' keyed by the parameter's index in the list given to the constructor; parameters
        that never received a gradient have no entry, as in torch."""
        state = {}
        steps = self._seg_step.cpu().tolist()
        index = {id(p): i for i, p in enumerate(self._all_params)}
        for p, o, t in zip(self.params, self._offsets, steps):
            if t <= 0:
                continue
            k = p.numel()
            state[index[id(p)]] = {"step": torch.tensor(float(t)),
                                   "exp_avg": self.exp_avg[o:o + k].view_as(p).clone(),
                                   "exp_avg_sq": self.exp_avg_sq[o:o + k].view_as(p).clone()}
        g = self.param_groups[0]
        group = {"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self._all_params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts a torch.optim.Adam state_dict of the same parameter list (e.g. the authors' optm.pt) or our own."""
        groups = sd["param_groups"]
        n = sum(len(g["params"]) for g in groups)
        if n != len(self._all_params):
            raise ValueError("optimizer checkpoint holds %d parameters, this model has %d" % (n, len(self._all_params)))
        g0 = groups[0]
        self.param_groups[0].update(lr=g0["lr"], betas=tuple(g0["betas"]), eps=g0["eps"])
        index = {id(p): i for i, p in enumerate(self._all_params)}
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        steps = []
        for p, o in zip(self.params, self._offsets):
            st = sd["state"].get(index[id(p)])
            if st is None:
                steps.append(0)
                continue
            k = p.numel()
            if st["exp_avg"].numel() != k:
                raise ValueError("optimizer checkpoint: parameter %d has %d elements, expected %d" % (index[id(p)], st["exp_avg"].numel(), k))
            self.exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(int(float(st["step"])))
        self._seg_step.copy_(torch.tensor(steps, dtype=torch.int32))

    def gather_grads(self):
        """Move the gradients autograd left in `p.grad` into the bucket (multi-tensor copy when the bucket is known to
        be zero, multi-tensor add otherwise), re-point `p.grad` at the bucket views and note which parameters have a
        gradient at all (the others are skipped by the step, like torch.optim.Adam).  Idempotent."""
        views, grads, act = [], [], []
        for p, v in zip(self.params, self._grad_views):
            g = p.grad
            act.append(0 if (g is None and not self.assume_all_active) else 1)
            if g is None or g.data_ptr() == v.data_ptr():
                continue
            views.append(v)
            grads.append(g.detach().to(torch.float32).view_as(v) if g.dtype != torch.float32 or g.shape != v.shape else g.detach())
        if views:
            with torch.no_grad():
                if self._bucket_clean:
                    torch._foreach_copy_(views, grads)
                else:
                    torch._foreach_add_(views, grads)
            self._bucket_clean = False
        self._h_active.numpy()[:] = act
        for p, v in zip(self.params, self._grad_views):
            if p.grad is not None:
                p.grad = v

    @property
    def flat_grad(self):
        """The gradient bucket, with every pending `p.grad` gathered."""
        self.gather_grads()
        return self._flat_grad

    def all_reduce(self, group=None):
        """The ONE collective of the view-sharded step: sum of the flat gradient bucket."""
        import torch.distributed as dist
        with stats.stage("allreduce"):
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)

    def step(self, grad_scale=1.0, zero_grad=True, graph_safe=True):
        """One Adam step.  Safe to capture in a CUDA graph (step counters, lr and grad_scale live on the device; the
        `graph_safe` argument is kept for callers of the earlier API and ignored)."""
        lib = _lib.load()
        dev = self.flat_param.device
        self.gather_grads()
        self.refresh_hyper(grad_scale)
        self._d_hyper.copy_(self._h_hyper, non_blocking=True)          # 8 bytes; a memcpy node when captured
        self._seg_active.copy_(self._h_active, non_blocking=True)
        g = self.param_groups[0]
        with torch.cuda.device(dev), stats.stage("adam", launches=2):
            st = lib.agr_adam_step_segments(self.flat_param.numel(), _p(self.flat_param.data_ptr()), _p(self._flat_grad.data_ptr()),
                                            _p(self.exp_avg.data_ptr()), _p(self.exp_avg_sq.data_ptr()), len(self.params),
                                            _p(self._chunk_seg.data_ptr()), _p(self._seg_active.data_ptr()), _p(self._seg_step.data_ptr()),
                                            _p(self._seg_corr.data_ptr()), _p(self._d_hyper.data_ptr()), g["betas"][0], g["betas"][1],
                                            g["eps"], int(zero_grad), _p(torch.cuda.current_stream(dev).cuda_stream))
        if st != _lib.AGR_OK:
            raise RuntimeError("agr_adam_step_segments failed: %d" % st)
        if zero_grad:   # the kernel cleared the active segments (inactive ones hold no gradient); next backward hands over
            self._bucket_clean = True                      # fresh tensors (set_to_none semantics)
            for p in self.params:
                p.grad = None

    def zero_grad(self):
        self._flat_grad.zero_()
        self._bucket_clean = True
        for p in self.params:
            p.grad = None
