"""Multi-GPU plumbing of the view-sharded step beyond the one gradient all-reduce: OWNER-COMPUTES for the
view-independent networks.

Views are sharded over the ranks; the three per-pose networks (position net, "other" net, colour prefix) do not depend on the
view, so in the plain scheme every rank repeats them — the part of the step that does not shrink with N.  Here each of them
runs on ONE owner rank; its outputs (per-Gaussian attributes: a few MB; the colour prefix state: ~70 MB) are broadcast to the
ranks that render, and in the backward pass the gradients the ranks computed for those outputs are summed onto the owner
(`reduce`), which back-propagates through its network.  Parameters stay replicated: the owner's parameter gradients reach
everybody through the same single all-reduce of the flat bucket as before (the other ranks contribute zeros).

torch.distributed (NCCL) does the transport — the exchanged tensors are small and sit between two long compute phases; there is
no compute to fuse them with."""
import torch
import torch.distributed as dist


def active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _flatten(t):
    """Memory-order flattening without a copy for contiguous / channels_last tensors."""
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous():
        return t.permute(0, 2, 3, 1).reshape(-1)
    return t.contiguous().reshape(-1)


def _unflatten(flat, meta):
    shape, cl = meta
    if cl:
        n, c, h, w = shape
        return flat.view(n, h, w, c).permute(0, 3, 1, 2)
    return flat.view(shape)


def describe(tensors):
    """Metadata the receivers need: ((shape, is_channels_last), ...), dtype."""
    metas = tuple((tuple(t.shape), bool(t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()))
                  for t in tensors)
    dtypes = {t.dtype for t in tensors}
    assert len(dtypes) == 1, "one dtype per exchange"
    return metas, next(iter(dtypes))


class _OwnerBroadcast(torch.autograd.Function):
    """forward: the owner's tensors -> every rank (one broadcast of one flat buffer); backward: sum of the ranks' gradients ->
    the owner (one reduce).  `dummy` (1 element, requires_grad) keeps the node in the graph on the ranks that pass no tensor,
    so that every rank issues the backward collective."""

    @staticmethod
    def forward(ctx, owner, metas, dtype, dummy, *tensors):
        rank = dist.get_rank()
        sizes = [int(torch.Size(m[0]).numel()) for m in metas]
        if rank == owner:
            assert len(tensors) == len(metas)
            flat = torch.cat([_flatten(t.detach()) for t in tensors]) if len(tensors) > 1 else _flatten(tensors[0].detach()).clone()
        else:
            flat = torch.empty(sum(sizes), dtype=dtype, device=dummy.device)
        dist.broadcast(flat, src=owner)
        ctx.owner, ctx.metas, ctx.sizes, ctx.is_owner, ctx.dtype, ctx.device = owner, metas, sizes, rank == owner, dtype, flat.device
        outs, o = [], 0
        for m, n in zip(metas, sizes):
            outs.append(_unflatten(flat[o:o + n], m))
            o += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        parts = []
        for g, m, n in zip(grads, ctx.metas, ctx.sizes):
            parts.append(torch.zeros(n, dtype=ctx.dtype, device=ctx.device) if g is None
                         else _flatten(g.to(ctx.dtype) if m[1] is False else g.to(ctx.dtype).contiguous(memory_format=torch.channels_last)))
        flat = torch.cat(parts) if len(parts) > 1 else parts[0].clone()
        dist.reduce(flat, dst=ctx.owner, op=dist.ReduceOp.SUM)
        zero = torch.zeros(1, dtype=torch.float32, device=flat.device)
        if not ctx.is_owner:
            return (None, None, None, zero)
        outs, o = [], 0
        for m, n in zip(ctx.metas, ctx.sizes):
            outs.append(_unflatten(flat[o:o + n], m))
            o += n
        return (None, None, None, zero, *outs)


def owner_broadcast(owner, metas, dtype, dummy, tensors):
    """-> tuple of tensors (one per meta), identical on every rank, differentiable back to the owner's `tensors`."""
    return _OwnerBroadcast.apply(owner, metas, dtype, dummy, *tensors)


def share_meta(owner, tensors):
    """One-time (eager, blocking) exchange of the shapes / dtype the receivers need; call outside graph capture."""
    obj = [describe(tensors) if dist.get_rank() == owner else None]
    dist.broadcast_object_list(obj, src=owner)
    return obj[0]
