"""On-disk formats either side of the hot path — host-side, no GPU work:

* posed-Gaussian PLY in the 3D-Gaussian-Splatting viewer layout, as the reference exports it
  (`gaussians/obj_io.py:24-99`): x y z, zero normals, SH DC of the BGR-swapped colour, 45 zero `f_rest_*`, logit
  opacity, log scales, raw quaternion; binary little-endian float32 `vertex` element (what `plyfile` writes on x86).
* trainer checkpoints `net.pt` / `optm.pt` (`main_avatar.py:778-813`).

The reference goes through the `plyfile` package, which this image does not have: the writer / reader below speak the
PLY container directly (parity with plyfile output is unpinned; the layout is covered by a known-answer test)."""
import os

import numpy as np
import torch

SH_C0 = 0.28209479177387814   # utils/sh_utils.py: RGB2SH(rgb) = (rgb - 0.5) / C0


def gaussian_ply_fields(sh_degree=3):
    n_rest = 3 * ((sh_degree + 1) ** 2 - 1)
    return (["x", "y", "z", "nx", "ny", "nz"] + ["f_dc_%d" % i for i in range(3)] + ["f_rest_%d" % i for i in range(n_rest)]
            + ["opacity"] + ["scale_%d" % i for i in range(3)] + ["rot_%d" % i for i in range(4)])


def _np(t):
    return t.detach().float().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t, np.float32)


def save_gaussians_as_ply(path, gaussian_vals, sh_degree=3):
    """`gaussian_vals`: positions (P,3), colors (P,3) in [0,1], opacity (P,1) in (0,1), scales (P,3) > 0, rotations (P,4)."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    xyz = _np(gaussian_vals["positions"])
    P = xyz.shape[0]
    rgb = _np(gaussian_vals["colors"])[:, [2, 1, 0]]
    opacity = _np(gaussian_vals["opacity"]).reshape(P, 1).astype(np.float64)
    n_rest = 3 * ((sh_degree + 1) ** 2 - 1)
    cols = np.concatenate([xyz, np.zeros_like(xyz), (rgb - 0.5) / SH_C0, np.zeros((P, n_rest), np.float32),
                           np.log(opacity / (1.0 - opacity)).astype(np.float32), np.log(_np(gaussian_vals["scales"])),
                           _np(gaussian_vals["rotations"])], axis=1).astype("<f4")
    fields = gaussian_ply_fields(sh_degree)
    assert cols.shape[1] == len(fields)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join("property float %s\n" % f for f in fields) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(cols).tobytes())


def read_ply_vertices(path):
    """Minimal PLY reader for a single float32 `vertex` element (binary little-endian or ascii) -> {name: (P,) array}."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("%s: not a PLY file" % path)
        fmt, count, names, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("%s: truncated PLY header" % path)
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] not in ("float", "float32"):
                    raise ValueError("%s: vertex property %s has type %s (only float32 supported)" % (path, tok[-1], tok[1]))
                names.append(tok[-1])
            elif tok[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            data = np.frombuffer(f.read(4 * count * len(names)), dtype="<f4")
        elif fmt == "ascii":
            data = np.array(f.read().split()[:count * len(names)], dtype=np.float32)
        else:
            raise ValueError("%s: unsupported PLY format %r" % (path, fmt))
    if data.size != count * len(names):
        raise ValueError("%s: truncated PLY body" % path)
    data = data.reshape(count, len(names))
    return {n: data[:, i] for i, n in enumerate(names)}


def load_gaussians_from_ply(path, device="cpu"):
    """Inverse of save_gaussians_as_ply with the reference's conventions (`gaussians/obj_io.py:45-99`): colours back
    to RGB order, sigmoid opacity, exp scales, normalised rotations, plus the (P,3,15) `features_extr`."""
    v = read_ply_vertices(path)

    def stack(prefix):
        names = sorted((n for n in v if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
        return np.stack([v[n] for n in names], 1) if names else np.zeros((len(v["x"]), 0), np.float32)

    xyz = np.stack([v["x"], v["y"], v["z"]], 1)
    f_dc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], 1)
    rest = stack("f_rest_")
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=device)
    return {
        "positions": t(xyz),
        "colors": t((f_dc * SH_C0 + 0.5)[:, [2, 1, 0]]),
        "opacity": torch.sigmoid(t(v["opacity"][:, None])),
        "scales": torch.exp(t(stack("scale_"))),
        "rotations": torch.nn.functional.normalize(t(stack("rot_"))),
        "features_extr": t(rest.reshape(rest.shape[0], 3, -1)) if rest.shape[1] else t(np.zeros((xyz.shape[0], 3, 0))),
    }


def save_ckpt(path, avatar_net, optimizer=None, epoch_idx=0, iter_idx=0):
    """`net.pt` {'epoch_idx','iter_idx','avatar_net'} and `optm.pt` {'avatar_net': optimizer state} (main_avatar.py:778-795)."""
    os.makedirs(path, exist_ok=True)
    torch.save({"epoch_idx": epoch_idx, "iter_idx": iter_idx, "avatar_net": avatar_net.state_dict()}, os.path.join(path, "net.pt"))
    if optimizer is not None:
        torch.save({"avatar_net": optimizer.state_dict()}, os.path.join(path, "optm.pt"))


def load_ckpt(path, avatar_net, optimizer=None, map_location="cpu"):
    """Returns (epoch_idx, iter_idx) (main_avatar.py:797-813).  `optimizer` may be a FlatAdam or a torch optimizer."""
    net = torch.load(os.path.join(path, "net.pt"), map_location=map_location, weights_only=False)
    if "avatar_net" in net:
        avatar_net.load_state_dict(net["avatar_net"])
    op = os.path.join(path, "optm.pt")
    if optimizer is not None and os.path.exists(op):
        sd = torch.load(op, map_location=map_location, weights_only=False)
        if "avatar_net" in sd:
            optimizer.load_state_dict(sd["avatar_net"])
    return net.get("epoch_idx", 0), net.get("iter_idx", 0)
