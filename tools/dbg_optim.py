import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animatablegaussians_b200 import optim
torch.manual_seed(0)
mk = lambda: torch.nn.Sequential(torch.nn.Linear(37, 300), torch.nn.Tanh(), torch.nn.Linear(300, 5), torch.nn.Linear(5, 3)).cuda()
a, b = mk(), mk()
b.load_state_dict(a.state_dict())
ref = torch.optim.Adam(a.parameters(), lr=1e-3)
opt = optim.FlatAdam(b.parameters(), lr=1e-3)
x = torch.randn(16, 37, device="cuda")
def body(net, o):
    net(x).pow(2).sum().backward()
    o.step()
def diff(tag):
    torch.cuda.synchronize()
    print(tag, max(float((p - q).abs().max()) for p, q in zip(a.parameters(), b.parameters())), opt._seg_step.tolist(), opt._d_hyper.tolist())
body(b, opt); body(a, ref); ref.zero_grad(); diff("eager")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    body(b, opt); body(a, ref); ref.zero_grad()
torch.cuda.current_stream().wait_stream(s); diff("side")
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    body(b, opt)
diff("after capture (b not stepped)")
body(a, ref); ref.zero_grad()
graph.replay(); diff("replay1")
for it in range(3):
    lr = 1e-3 * (0.5 ** (it + 1))
    for grp in ref.param_groups: grp["lr"] = lr
    opt.param_groups[0]["lr"] = lr
    opt.refresh_hyper()
    graph.replay()
    body(a, ref); ref.zero_grad()
    diff("loop %d lr %g" % (it, lr))
