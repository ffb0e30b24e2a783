"""lsx_b200 — B200-native (sm_100a) implementation of LangScene-X's rasterizer hot path.

Sub-modules
  _lib       ctypes binding of liblsx_b200.so (C ABI in include/lsx_rasterizer.h)
  ops        tensor-level functions with the reference's `_C` signatures
  multiview  view-sharded multi-GPU gradient accumulation (one process per GPU, NCCL all-reduce)
  synthetic  seeded synthetic scenes/cameras used by tests and bench.py
"""
__version__ = "0.1.0"
