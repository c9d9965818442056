import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from animatablegaussians_b200 import avatar, styleunet_ops as ops, lbs
torch.cuda.set_device(0)
wl = bench.ProductWorkload(0, 1, torch.device("cuda", 0))
net = wl.net
with torch.no_grad():
    pose = wl.d_pose[:3]
    for name, style in (("position_net", net.position_style), ("other_net", net.other_style), ("color_net", net.color_style)):
        m, _ = getattr(net, name)([style], pose[None], randomize_noise=False)
        g = net._gather(m)
        print(name, "map std %.3f absmax %.3f  gathered std %.3f absmax %.3f nan %d" % (float(m.std()), float(m.abs().max()), float(g.std()), float(g.abs().max()), int(torch.isnan(m).sum())))
    op, sc, rot = net.get_others(pose)
    pts = net.get_positions(pose)
    print("opacity mean %.3f  scale median %.5f max %.5f  q99.9 %.5f" % (float(op.mean()), float(sc.median()), float(sc.max()), float(sc.flatten().kthvalue(int(sc.numel()*0.999)).values)))
    print("offset std %.4f max %.4f" % (float((pts - net.init_points).std()), float((pts - net.init_points).abs().max())))
    pos, r2 = lbs.transform_cano2live(net.lbs, wl.d_mats, pts, rot)
    print("posed rot norm: mean %.3f min %.3f max %.3f" % (float(r2.norm(dim=1).mean()), float(r2.norm(dim=1).min()), float(r2.norm(dim=1).max())))
