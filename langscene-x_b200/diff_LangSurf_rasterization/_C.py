"""Stand-in for the reference's pybind module `diff_LangSurf_rasterization._C`
(diff-langsurf-rasterizer/ext.cpp:15-19): same three functions, same argument order and return tuples,
implemented on liblsx_b200.so."""
from lsx_b200.ops import mark_visible, rasterize_gaussians, rasterize_gaussians_backward  # noqa: F401
