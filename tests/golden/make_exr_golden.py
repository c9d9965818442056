"""Writes tests/golden/exr_*.exr with OpenCV — the library the reference's dataset reads and writes its position maps with
(dataset/dataset_mv_rgb.py:147, gen_data/gen_pos_maps.py) — plus the arrays OpenCV reads back from them (exr_golden.npz).
    python tests/golden/make_exr_golden.py"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"   # as main_avatar.py:4 does
import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(5)
out = {}
yy, xx = np.meshgrid(np.linspace(-1, 1, 20), np.linspace(-1, 1, 48), indexing="ij")
pos = np.stack([xx * 0.4, yy * 0.9, 0.1 * np.sin(3 * xx)], -1).astype(np.float32)
pos[rng.random(pos.shape[:2]) < 0.3] = 0.0                      # empty texels, like a real position map
cases = {
    "zip_f32": (pos, []),                                                                        # OpenCV's defaults: FLOAT, ZIP
    "zips_f32": (pos, [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_ZIPS]),
    "none_f32": (pos, [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_NO]),
    "rle_f32": (pos, [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_RLE]),
    "zip_f16": (pos, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF]),
    "zip_noise_f32": (rng.normal(size=(37, 21, 3)).astype(np.float32), []),                      # incompressible: raw blocks, ragged last block
    "zip_gray_f32": (rng.normal(size=(18, 9)).astype(np.float32), []),
}
for name, (img, flags) in cases.items():
    path = os.path.join(HERE, "exr_%s.exr" % name)
    assert cv2.imwrite(path, img, flags)
    out[name] = cv2.imread(path, cv2.IMREAD_UNCHANGED)
np.savez_compressed(os.path.join(HERE, "exr_golden.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
