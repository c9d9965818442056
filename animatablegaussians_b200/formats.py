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


# ------------------------------------------------------------------------------------------ OpenEXR position maps
# The SMPL position maps the datasets store (dataset/dataset_mv_rgb.py:146-151: cv.imread(.../smpl_pos_map/%08d.exr,
# cv.IMREAD_UNCHANGED), written by gen_data/gen_pos_maps.py with cv.imwrite) are single-part scan-line OpenEXR files with
# FLOAT channels and ZIP compression (what OpenCV writes).  The reader below decodes that container directly — no OpenEXR
# library, no OpenCV — and returns what cv.imread(..., IMREAD_UNCHANGED) returns: (H, W, C) float32 with the channels in
# B, G, R[, A] order.  Pinned against OpenCV's own decoder (tests/test_formats.py: a committed cv2-written fixture, and a
# live comparison when cv2 is importable).
_EXR_MAGIC = 20000630
_EXR_LINES_PER_BLOCK = {0: 1, 1: 1, 2: 1, 3: 16}      # NONE, RLE, ZIPS, ZIP


def _exr_header(buf):
    import struct
    magic, version = struct.unpack_from("<II", buf, 0)
    if magic != _EXR_MAGIC:
        raise ValueError("not an OpenEXR file")
    if version & 0x200 or version & 0x800 or version & 0x1000:
        raise ValueError("tiled / deep / multi-part OpenEXR files are not position maps; only scan-line images are read")
    attrs, i = {}, 8
    while True:
        j = buf.index(b"\0", i)
        name = buf[i:j].decode()
        if not name:
            return attrs, j + 1
        k = buf.index(b"\0", j + 1)
        typ = buf[j + 1:k].decode()
        size = struct.unpack_from("<i", buf, k + 1)[0]
        attrs[name] = (typ, buf[k + 5:k + 5 + size])
        i = k + 5 + size


def _exr_unzip(raw, expect):
    import zlib
    if len(raw) == expect:                 # the writer stores a block raw when compression does not shrink it
        return np.frombuffer(raw, np.uint8)
    t = np.frombuffer(zlib.decompress(raw), np.uint8)
    if t.size != expect:
        raise ValueError("OpenEXR block decompressed to %d bytes, expected %d" % (t.size, expect))
    t = (np.cumsum(t.astype(np.int64) - 128) + 128).astype(np.uint8)       # undo the byte-delta predictor: t[i] += t[i-1] - 128
    half = (expect + 1) // 2
    out = np.empty(expect, np.uint8)                                       # undo the even/odd byte split
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out


def _exr_unrle(raw, expect):
    if len(raw) == expect:
        return np.frombuffer(raw, np.uint8)
    out, i = bytearray(), 0
    while i < len(raw):
        n = raw[i] - 256 if raw[i] > 127 else raw[i]
        i += 1
        if n < 0:
            out += raw[i:i - n]
            i -= n
        else:
            out += bytes([raw[i]]) * (n + 1)
            i += 1
    t = np.frombuffer(bytes(out), np.uint8)
    t = (np.cumsum(t.astype(np.int64) - 128) + 128).astype(np.uint8)
    half = (expect + 1) // 2
    res = np.empty(expect, np.uint8)
    res[0::2] = t[:half]
    res[1::2] = t[half:]
    return res


def read_exr(path):
    """-> (H, W, C) float32 like cv.imread(path, cv.IMREAD_UNCHANGED): channels B, G, R[, A] for colour images, the file's
    (alphabetical) channel order otherwise; (H, W) for a single channel.  HALF channels are widened to float32."""
    import struct
    buf = open(path, "rb").read()
    attrs, pos = _exr_header(buf)
    comp = attrs["compression"][1][0]
    if comp not in _EXR_LINES_PER_BLOCK:
        raise ValueError("OpenEXR compression %d is not supported (NONE, RLE, ZIPS, ZIP are)" % comp)
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    W, H = x1 - x0 + 1, y1 - y0 + 1
    chans, c = [], attrs["channels"][1]
    i = 0
    while c[i] != 0:
        j = c.index(b"\0", i)
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", c, j + 1)
        if xs != 1 or ys != 1 or ptype not in (1, 2):
            raise ValueError("OpenEXR channel %r: only unsampled HALF / FLOAT channels are read" % c[i:j])
        chans.append((c[i:j].decode(), ptype))
        i = j + 17
    bpp = [2 if t == 1 else 4 for _, t in chans]
    line_bytes = W * sum(bpp)
    lines = _EXR_LINES_PER_BLOCK[comp]
    n_blocks = (H + lines - 1) // lines
    offsets = struct.unpack_from("<%dQ" % n_blocks, buf, pos)
    planes = [np.empty((H, W), np.float32) for _ in chans]
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        rows = min(lines, y1 + 1 - y)
        raw = buf[off + 8:off + 8 + size]
        data = raw if comp == 0 else (_exr_unrle(raw, rows * line_bytes) if comp == 1 else _exr_unzip(raw, rows * line_bytes))
        data = np.frombuffer(bytes(data), np.uint8).reshape(rows, line_bytes)
        o = 0
        for ci, ((_, ptype), b) in enumerate(zip(chans, bpp)):
            seg = np.ascontiguousarray(data[:, o:o + W * b])
            planes[ci][y - y0:y - y0 + rows] = seg.view("<f2" if ptype == 1 else "<f4").astype(np.float32).reshape(rows, W)
            o += W * b
    names = [n for n, _ in chans]
    order = [names.index(n) for n in ("B", "G", "R", "A") if n in names] if {"B", "G", "R"} <= set(names) else list(range(len(names)))
    if len(order) == 1:
        return planes[order[0]]
    return np.stack([planes[k] for k in order], axis=-1)


def load_smpl_pos_map(path, device="cpu"):
    """The per-pose input of the three U-Nets as the dataset builds it (dataset_mv_rgb.py:146-151): the (H, 2H, 3) position
    map split into its front | back halves and stacked on the channel axis -> (6, H, H) float32."""
    m = read_exr(path)
    h = m.shape[1] // 2
    pos_map = np.concatenate([m[:, :h], m[:, h:]], axis=2).transpose(2, 0, 1)
    return torch.from_numpy(np.ascontiguousarray(pos_map, dtype=np.float32)).to(device)
