"""Debug print: LPIPS ours vs the reference module on the GPU, value / per layer / gradient, fp32 and bf16."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "AnimatableGaussians"))
from network.lpips import LPIPS as Ref
from animatablegaussians_b200 import lpips, styleunet_ops as ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
ref = Ref(net="vgg", pnet_rand=True, verbose=False).cuda()
ours = lpips.LPIPS(net="vgg", pnet_rand=True).cuda(); ours.load_state_dict(ref.state_dict(), strict=True)
for dtype, H, W, N in ((torch.float32, 64, 64, 1), (torch.float32, 96, 80, 2), (torch.bfloat16, 128, 128, 1)):
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand(N, 3, H, W, device="cuda", generator=g)
    gt = (img + 0.25 * torch.randn(N, 3, H, W, device="cuda", generator=g)).clamp(0, 1)
    a = img.clone().requires_grad_(True)
    v_ref, per_ref = ref.forward(a, gt, retPerLayer=True, normalize=True); v_ref.mean().backward()
    ops.set_compute_dtype(dtype)
    b = img.clone().requires_grad_(True)
    v, per = ours.forward(b, gt, retPerLayer=True, normalize=True); v.mean().backward()
    # trunk features side by side
    with torch.no_grad():
        x = ops.to_compute(torch.cat([ours.scaling_layer(2 * img - 1), ours.scaling_layer(2 * gt - 1)], 0))
        fo = ours.net(x)
        fr0 = ref.net.forward(ref.scaling_layer(2 * img - 1)); fr1 = ref.net.forward(ref.scaling_layer(2 * gt - 1))
        ferr = [float((fo[k][:N].float() - fr0[k]).abs().max() / fr0[k].abs().max()) for k in range(5)]
    ops.set_compute_dtype(torch.float32)
    print(dtype, H, W, N, "val", v.flatten().tolist(), "ref", v_ref.flatten().tolist())
    print("  per ours", [p.flatten().tolist() for p in per]); print("  per ref ", [p.flatten().tolist() for p in per_ref])
    print("  feature max-rel err per slice", ferr, " grad rel L2 %.3e" % float((b.grad - a.grad).norm() / a.grad.norm()))
