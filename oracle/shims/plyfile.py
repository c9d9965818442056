"""TEST / BENCH INFRASTRUCTURE — import stand-in for the third-party package `plyfile`, which the reference imports at
module scope (gaussians/gaussian_model.py:17) but never calls on the avatar path.  Absent here (no network)."""


class PlyData:
    @staticmethod
    def read(*a, **k):
        raise RuntimeError("plyfile shim: not available in this environment")


class PlyElement:
    @staticmethod
    def describe(*a, **k):
        raise RuntimeError("plyfile shim: not available in this environment")
