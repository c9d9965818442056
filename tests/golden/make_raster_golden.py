"""Generates tests/golden/raster_*.npz by running the UNMODIFIED reference CUDA kernels (oracle/_ref,
built from /root/reference by oracle/build_ref.py) on a B200:
    gpurun -- 'python tests/golden/make_raster_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/.  Scenes are regenerated from (kwargs, seed) by
tests/util.random_scene, so only outputs + upstream gradients are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402
from tests import raster_harness as Hn  # noqa: E402

CASES = {
    "basic": dict(P=400, W=64, H=48, seed=201),
    "sh2": dict(P=300, W=48, H=48, seed=202, sh_degree=2),
    "cov": dict(P=300, W=40, H=56, seed=203, cov_precomp=True),
    "dense": dict(P=1500, W=64, H=64, seed=204, spread=0.4),
}

if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, kw in CASES.items():
        sc = util.random_scene(**kw)
        g = util.upstream_grads(sc["W"], sc["H"], kw["seed"])
        ref = Hn.run_reference(sc, g)
        d = dict(scene_keys=np.array(list(kw.keys())), g_color=g[0], g_depth=g[1], g_alpha=g[2], color=ref["color"],
                 depth=ref["depth"], alpha=ref["alpha"], radii=ref["radii"], num_rendered=np.int64(ref["R"]))
        for k, v in kw.items():
            d[k] = np.array(v)
        for k, v in ref["grads"].items():
            if k in ("conic", "depth"):
                continue
            d["grad_" + k] = v.astype(np.float32)
        np.savez_compressed(os.path.join(out_dir, "raster_%s.npz" % name), **d)
        print("wrote", name, "R =", ref["R"])
