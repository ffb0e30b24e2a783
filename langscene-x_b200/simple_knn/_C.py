"""Stand-in for the reference's pybind module `simple_knn._C` (simple-knn/ext.cpp:15-17)."""
from lsx_b200.ops import distCUDA2  # noqa: F401
