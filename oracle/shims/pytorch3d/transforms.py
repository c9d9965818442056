"""pytorch3d.transforms.{quaternion_to_matrix, matrix_to_quaternion} (0.7.x rotation_conversions.py), restated in
oracle/lbs_oracle.py — see its header for the provenance note."""
from oracle.lbs_oracle import matrix_to_quaternion, quaternion_to_matrix  # noqa: F401
