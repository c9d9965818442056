"""TEST INFRASTRUCTURE — float64 PyTorch-autograd restatement of the reference rasterizer's FORWARD
semantics (RAST/cuda_rasterizer/forward.cu:74-381) for tiny scenes.  Its gradients come from autograd,
not from the reference's hand-derived backward, so agreement with oracle/raster_oracle.c pins the
backward formulas (backward.cu:144-601) independently.  The reference's gradient conventions are kept:
  * the min(0.99, o*G) clamp is passed straight through (backward.cu:530,582),
  * rejected pairs (power>0, alpha<1/255), the T<1e-4 termination and culling carry no gradient,
  * the 1.3*tanfov clamp of t.x/t.y blocks the gradient when active (backward.cu:175-176),
  * the quaternion is used un-normalised (forward.cu:127).
Pure PyTorch loops over Gaussians: small cases only."""
import torch


def render(means3D, colors, opacity, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, bg):
    dt = torch.float64
    P = means3D.shape[0]
    view = torch.as_tensor(viewmatrix, dtype=dt).reshape(4, 4)   # memory order of the reference = standard^T
    proj = torch.as_tensor(projmatrix, dtype=dt).reshape(4, 4)
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    t = ph @ view            # row-vector times transposed matrix == transformPoint4x3
    p_hom = ph @ proj
    p_w = 1.0 / (p_hom[:, 3] + 0.0000001)
    p_proj = p_hom[:, :3] * p_w[:, None]
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)

    r, x, y, z = rotations.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
    Sigma = R @ torch.diag_embed(scales * scales) @ R.transpose(1, 2)

    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = t[:, 0] / t[:, 2], t[:, 1] / t[:, 2]
    # clamp: value clamped, gradient blocked entirely when active (x_grad_mul / y_grad_mul)
    tx = torch.where((txtz < -limx) | (txtz > limx), (txtz.clamp(-limx, limx) * t[:, 2]).detach(), t[:, 0])
    ty = torch.where((tytz < -limy) | (tytz > limy), (tytz.clamp(-limy, limy) * t[:, 2]).detach(), t[:, 1])
    tz = t[:, 2]
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], 1).reshape(P, 2, 3)
    Rw2c = view[:3, :3].T     # standard world->camera rotation
    A = J @ Rw2c
    cov2 = A @ Sigma @ A.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((p_proj[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((p_proj[:, 1] + 1.0) * H - 1.0) * 0.5
    visible = (t[:, 2] > 0.2).detach()

    order = sorted([i for i in range(P) if visible[i]], key=lambda i: (float(t[i, 2].detach().float()), i))
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    T = torch.ones(H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    C = torch.zeros(3, H, W, dtype=dt)
    Dm = torch.zeros(H, W, dtype=dt)
    Am = torch.zeros(H, W, dtype=dt)
    for i in order:
        # tile rectangle test of the reference (per 16x16 tile): pixels outside the Gaussian's tiles never see it
        rad = float(radius[i])
        x0 = max(0, int((float(px[i]) - rad) / 16)) * 16
        y0 = max(0, int((float(py[i]) - rad) / 16)) * 16
        x1 = max(0, int((float(px[i]) + rad + 15) / 16)) * 16
        y1 = max(0, int((float(py[i]) + rad + 15) / 16)) * 16
        in_rect = (xs >= x0) & (xs < x1) & (ys >= y0) & (ys < y1)
        dx, dy = px[i] - xs, py[i] - ys
        power = -0.5 * (conic[i, 0] * dx * dx + conic[i, 2] * dy * dy) - conic[i, 1] * dx * dy
        G = torch.exp(power)
        raw = opacity[i, 0] * G
        alpha = raw + (torch.clamp(raw, max=0.99) - raw).detach()   # straight-through clamp
        ok = in_rect & (~done) & (power <= 0) & (alpha >= 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = ok & (test_T < 0.0001)
        done = done | stop.detach()
        ok = (ok & ~stop).detach()
        w = torch.where(ok, alpha * T, torch.zeros_like(T))
        C = C + colors[i][:, None, None] * w[None]
        Dm = Dm + t[i, 2] * w
        Am = Am + w
        T = torch.where(ok, test_T, T)
    bg = torch.as_tensor(bg, dtype=dt)
    return C + T[None] * bg[:, None, None], Dm[None], Am[None]
