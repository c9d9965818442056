"""TEST / BENCH INFRASTRUCTURE — stand-in for the third-party package pytorch3d == 0.7.4 (reference requirements.txt:9),
which is neither vendored in /root/reference nor installable here (no network).  Only the three functions the
reference's avatar path calls exist (network/avatar.py:87,89; gaussians/gaussian_model.py:170); they restate the
published pytorch3d 0.7.x algorithms.  PARITY UNPINNED: nothing reference-held checks these restatements.
Put on sys.path ONLY by oracle/ref_stock.py when it runs the reference's own Python; the product never imports it."""
from . import ops, transforms  # noqa: F401
