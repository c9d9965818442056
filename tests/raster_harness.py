"""GPU-side helpers shared by the -m gpu parity tests: run a scene through (a) the product's C ABI via the
reference-shaped Python surface and (b) the UNMODIFIED reference kernels (oracle/_ref)."""
import numpy as np
import torch

from animatablegaussians_b200 import rasterizer as R


def to_dev(a, dev="cuda"):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def settings_for(sc, dev="cuda", debug=False):
    return R.GaussianRasterizationSettings(
        image_height=sc["H"], image_width=sc["W"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], bg=to_dev(sc["bg"], dev),
        scale_modifier=1.0, viewmatrix=to_dev(sc["viewmatrix"], dev), projmatrix=to_dev(sc["projmatrix"], dev),
        sh_degree=sc["sh_degree"], campos=to_dev(sc["campos"], dev), prefiltered=False, debug=debug)


def run_product(sc, grads=None, dev="cuda"):
    """Through GaussianRasterizer.forward + autograd (the call the reference's render3 makes)."""
    req = grads is not None
    def leaf(a):
        if a is None:
            return None
        t = to_dev(a, dev)
        t.requires_grad_(req)
        return t
    means3D, opac = leaf(sc["xyz"]), leaf(sc["opacity"])
    rgb, sh, scales, rot, cov = leaf(sc["rgb"]), leaf(sc["sh"]), leaf(sc["scales"]), leaf(sc["rotations"]), leaf(sc["cov3D"])
    means2D = torch.zeros_like(means3D, requires_grad=req)
    rast = R.GaussianRasterizer(settings_for(sc, dev))
    color, radii, depth, alpha = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=sh, colors_precomp=rgb,
                                      scales=scales, rotations=rot, cov3D_precomp=cov)
    out = dict(color=color.detach().cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.detach().cpu().numpy(),
               alpha=alpha.detach().cpu().numpy())
    if req:
        gc, gd, ga = (to_dev(g, dev) for g in grads)
        torch.autograd.backward([color, depth, alpha], [gc, gd, ga])
        g = dict(means3D=means3D.grad, means2D=means2D.grad, opacity=opac.grad)
        if rgb is not None: g["colors"] = rgb.grad
        if sh is not None: g["sh"] = sh.grad
        if scales is not None: g["scales"] = scales.grad; g["rotations"] = rot.grad
        if cov is not None: g["cov3D"] = cov.grad
        out["grads"] = {k: v.detach().cpu().numpy() for k, v in g.items()}
    return out


def run_reference(sc, grads=None, dev="cuda"):
    from oracle.ref_rasterizer import RefRasterizer
    ref = RefRasterizer()
    d = lambda k: to_dev(sc[k], dev)
    torch.cuda.synchronize()
    color, radii, depth, alpha = ref.forward(d("bg"), d("xyz"), d("rgb"), d("opacity"), d("scales"), d("rotations"), 1.0,
                                             d("cov3D"), d("viewmatrix"), d("projmatrix"), sc["tanfovx"], sc["tanfovy"],
                                             sc["H"], sc["W"], sh=d("sh"), degree=sc["sh_degree"], campos=d("campos"))
    torch.cuda.synchronize()
    out = dict(color=color.cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.cpu().numpy(), alpha=alpha.cpu().numpy(),
               R=ref.num_rendered)
    if grads is not None:
        g = ref.backward(*(to_dev(x, dev) for x in grads))
        torch.cuda.synchronize()
        out["grads"] = {k: v.cpu().numpy() for k, v in g.items()}
    return out


GRAD_KEYS = ("means3D", "means2D", "colors", "opacity", "scales", "rotations", "cov3D", "sh")


def compare(tag, got, want, tol, assert_close, check_radii=True):
    report = {}
    if check_radii:
        assert np.array_equal(got["radii"], want["radii"]), tag + ": radii differ"
    for k in ("color", "depth", "alpha"):
        report[k] = assert_close(tag + ":" + k, got[k], want[k], tol)
    if "grads" in got and "grads" in want:
        for k in GRAD_KEYS:
            if k in got["grads"] and k in want["grads"] and got["grads"][k].size and want["grads"][k].size:
                a, b = got["grads"][k], want["grads"][k]
                report["grad_" + k] = assert_close(tag + ":grad_" + k, a.reshape(b.shape), b, tol)
    return report
