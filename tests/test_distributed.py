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
    assert all(p.grad.data_ptr() >= opt.flat_grad.data_ptr() for p in net.parameters())  # grads live in the bucket
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
