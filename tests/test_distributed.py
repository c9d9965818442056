"""N>1 host logic on CPU (gloo, world_size 2): the view shard covers every view exactly once and the ONE collective of
the step — all-reduce(sum) of the flat gradient bucket — yields the gradient of the un-sharded step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from animatablegaussians_b200 import optim
    torch.manual_seed(0)  # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    opt = optim.FlatAdam(net.parameters(), lr=1e-3)
    n_views = 16
    views = list(range(rank, n_views, world))       # bench.py's shard rule
    g = torch.Generator().manual_seed(1)
    xs = torch.randn(n_views, 4, 7, generator=g)      # one "view" = one mini-batch
    loss = sum(net(xs[v]).pow(2).sum() for v in views)
    loss.backward()
    fresh = [p.grad for p in net.parameters()]
    bucket = opt.flat_grad                               # gathers autograd's tensors into the bucket ...
    lo, hi = bucket.data_ptr(), bucket.data_ptr() + bucket.numel() * 4
    assert all(lo <= p.grad.data_ptr() < hi for p in net.parameters())   # ... and re-points p.grad at it
    assert all(torch.equal(f, p.grad) for f, p in zip(fresh, net.parameters()))
    opt.all_reduce()
    ret[rank] = (views, opt.flat_grad.clone(), opt.flat_param.clone())
    dist.destroy_process_group()


def test_view_shard_and_single_allreduce():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    views0, g0, p0 = ret[0]
    views1, g1, p1 = ret[1]
    assert sorted(views0 + views1) == list(range(16)) and not set(views0) & set(views1)
    assert torch.equal(g0, g1) and torch.equal(p0, p1)            # replicas stay bit-identical
    # un-sharded gradient
    from animatablegaussians_b200 import optim
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    opt = optim.FlatAdam(net.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    xs = torch.randn(16, 4, 7, generator=g)
    sum(net(xs[v]).pow(2).sum() for v in range(16)).backward()
    assert torch.allclose(opt.flat_grad, g0, rtol=1e-5, atol=1e-6)


def test_lazy_gradient_gather_semantics():
    """FlatAdam: p.grad is None after zero_grad (autograd then hands over its tensor: no accumulate launch); the bucket
    equals torch's own accumulation whether backward runs once, twice before a gather, or on both sides of one."""
    from animatablegaussians_b200 import optim
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    ref = [p.detach().clone().requires_grad_(True) for p in net.parameters()]
    opt = optim.FlatAdam(net.parameters(), lr=1e-3)
    assert all(p.grad is None for p in net.parameters())
    x1, x2, x3 = torch.randn(4, 6), torch.randn(4, 6), torch.randn(4, 6)

    def f(ps, x):
        return (torch.tanh(x @ ps[0].t() + ps[1]) @ ps[2].t() + ps[3]).pow(2).sum()

    ps = list(net.parameters())
    f(ps, x1).backward(); f(ps, x2).backward()          # two backwards before the first gather
    g12 = opt.flat_grad.clone()
    f(ps, x3).backward()                                 # accumulates into the bucket views in place
    g123 = opt.flat_grad.clone()
    (f(ref, x1) + f(ref, x2)).backward()
    want12 = torch.cat([r.grad.reshape(-1) for r in ref])
    f(ref, x3).backward()
    want123 = torch.cat([r.grad.reshape(-1) for r in ref])
    def flat(bucket):   # the bucket pads every tensor to 4 elements: pick the parameters' own ranges
        return torch.cat([bucket[v.storage_offset():v.storage_offset() + v.numel()] for v in opt._grad_views])

    assert torch.allclose(flat(g12), want12, rtol=1e-6, atol=1e-7) and torch.allclose(flat(g123), want123, rtol=1e-6, atol=1e-7)
    opt.zero_grad()
    assert all(p.grad is None for p in ps) and float(opt.flat_grad.abs().max()) == 0.0
    f(ps, x2).backward()
    for r in ref:
        r.grad = None
    f(ref, x2).backward()
    assert torch.allclose(flat(opt.flat_grad), torch.cat([r.grad.reshape(-1) for r in ref]), rtol=1e-6, atol=1e-7)


def _owner_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from animatablegaussians_b200 import parallel
    owner = 1
    torch.manual_seed(3)
    w = torch.randn(5, 3, requires_grad=True)                         # the owner's "network"
    x4 = torch.randn(1, 4, 3, 2).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dummy = torch.zeros(1, requires_grad=True)
    tensors = [w * 2.0, x4 * 1.0, w * 3.0] if rank == owner else []
    metas, dtype = parallel.share_meta(owner, tensors)
    a, b, _unused = parallel.owner_broadcast(owner, metas, dtype, dummy, tensors)      # the third output receives no gradient
    # every rank uses the outputs with its own (rank-dependent) weights, like each rank rendering its own views
    loss = (a * (rank + 1.0)).sum() + (b * b * (rank + 2.0)).sum()
    loss.backward()
    ret[rank] = (a.detach().clone(), b.detach().clone(), b.is_contiguous(memory_format=torch.channels_last),
                 None if w.grad is None else w.grad.clone(), None if x4.grad is None else x4.grad.clone())
    dist.destroy_process_group()


def test_owner_computes_broadcast_and_reduce():
    """parallel.owner_broadcast (gloo, world_size 2): every rank sees the owner's tensors (channels_last preserved); the
    owner's inputs receive the SUM of the ranks' gradients, the other ranks' copies of the network receive nothing."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_owner_worker, args=(world, port, ret), nprocs=world, join=True)
    a0, b0, cl0, gw0, gx0 = ret[0]
    a1, b1, cl1, gw1, gx1 = ret[1]
    torch.manual_seed(3)
    w = torch.randn(5, 3)
    x4 = torch.randn(1, 4, 3, 2)
    assert torch.equal(a0, a1) and torch.equal(a0, w * 2.0) and torch.equal(b0, b1) and torch.equal(b0, x4) and cl0 and cl1
    assert gw0 is None and gx0 is None                                          # rank 0 does not own the network
    assert torch.allclose(gw1, torch.full((5, 3), 2.0 * (1.0 + 2.0)))            # d/dw of sum_r (2w)(r+1)
    assert torch.allclose(gx1, 2.0 * x4 * (2.0 + 3.0))                            # d/dx of sum_r x^2 (r+2)
