"""FlatAdam (fused, segmented Adam over the flat bucket) against torch.optim.Adam: parameters without a gradient are
skipped (no moment decay, no step), per-parameter step counters, and a captured CUDA graph follows the learning-rate
schedule the trainer writes into param_groups[0]['lr'] (main_avatar.py:61-68, 184-189)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nets():
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(37, 300), torch.nn.Tanh(), torch.nn.Linear(300, 5), torch.nn.Linear(5, 3)).cuda()
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    return a, b


def test_matches_torch_adam_with_frozen_and_late_parameters(built_lib):
    from animatablegaussians_b200 import optim
    a, b = _nets()
    ref = torch.optim.Adam(a.parameters(), lr=5e-4)
    opt = optim.FlatAdam(b.parameters(), lr=5e-4)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(12):
        lr = optim.cosine_lr(5e-4, it, 12)
        for grp in ref.param_groups:
            grp["lr"] = lr
        for grp in opt.param_groups:          # the trainer's update_lr idiom
            grp["lr"] = lr
        x = torch.randn(16, 37, device="cuda", generator=g)
        for net in (a, b):
            # the last layer joins after 4 iterations (pretraining never reaches it), the first is frozen on odd iterations
            h = net[2](net[1](net[0](x)))
            loss = h.pow(2).sum() if it < 4 else net[3](h).pow(2).sum()
            net[0].weight.requires_grad_(it % 2 == 0)
            net[0].bias.requires_grad_(it % 2 == 0)
            loss.backward()
            net[0].weight.requires_grad_(True); net[0].bias.requires_grad_(True)
        ref.step(); ref.zero_grad()
        opt.step()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=1e-7), it
    sd, sr = opt.state_dict()["state"], ref.state_dict()["state"]
    assert sorted(sd) == sorted(sr)
    for i in sr:
        assert float(sd[i]["step"]) == float(sr[i]["step"]), i
        assert torch.allclose(sd[i]["exp_avg"], sr[i]["exp_avg"], rtol=2e-4, atol=1e-7)
        assert torch.allclose(sd[i]["exp_avg_sq"], sr[i]["exp_avg_sq"], rtol=2e-4, atol=1e-9)


def test_captured_graph_follows_the_lr_schedule(built_lib):
    from animatablegaussians_b200 import optim
    a, b = _nets()
    ref = torch.optim.Adam(a.parameters(), lr=1e-3)
    opt = optim.FlatAdam(b.parameters(), lr=1e-3)
    x = torch.randn(16, 37, device="cuda")

    def body(net, o):
        net(x).pow(2).sum().backward()
        o.step()

    body(b, opt); body(a, ref); ref.zero_grad()          # warm-up step (eager) on both
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body(b, opt); body(a, ref); ref.zero_grad()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body(b, opt)
    body(a, ref); ref.zero_grad()                        # the captured step itself does not execute: replay below
    graph.replay()
    torch.cuda.synchronize()
    for it in range(5):
        lr = 1e-3 * (0.5 ** (it + 1))
        for grp in ref.param_groups:
            grp["lr"] = lr
        opt.param_groups[0]["lr"] = lr
        opt.refresh_hyper()
        graph.replay()
        body(a, ref); ref.zero_grad()
        torch.cuda.synchronize()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=1e-7), (it, float((p - q).abs().max()), opt._seg_step.tolist(), opt._d_hyper.tolist())
    assert opt.t == 8
