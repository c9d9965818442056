import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from tests import util
from tests.golden import make_styleunet_golden as G
name = sys.argv[1] if len(sys.argv) > 1 else "small"
cfg, use_view, stride = G.CASES[name]
z = np.load("tests/golden/styleunet_%s.npz" % name)
def run(oracle):
    if oracle:
        from oracle import styleunet_oracle as m
        m.set_compute_dtype(torch.float32)
        net = m.DualStyleUNet(**cfg)
    else:
        from animatablegaussians_b200 import styleunet as m, styleunet_ops as ops
        ops.set_compute_dtype(torch.float32)
        net = m.DualStyleUNet(**cfg)
    G.fill_state(net); net = net.cuda()
    cond, style, vf1, vf2, up = (t.cuda() if t is not None else None for t in G.inputs(cfg, use_view))
    cond.requires_grad_(True)
    out, _ = net([style], cond, randomize_noise=False, view_feature1=vf1, view_feature2=vf2)
    (out * up).sum().backward()
    res = {"out": util.rel_err(out.detach().cpu().numpy()[..., ::stride, ::stride], z["out"]),
           "grad_cond": util.rel_err(cond.grad.cpu().numpy(), z["grad_cond"])}
    named = dict(net.named_parameters())
    for k in G.GRAD_KEYS:
        if "grad:" + k in z.files:
            res[k] = util.rel_err(named[k].grad.cpu().numpy(), z["grad:" + k])
    return res
for o in (True, False):
    print("oracle " if o else "product", {k: "%.1e" % v for k, v in run(o).items()})
