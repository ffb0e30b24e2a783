#!/usr/bin/env python
"""Generates tests/golden/depth_normal.npz by running the REFERENCE's own depth -> normal code on CPU.

Imports field_construction/utils/graphics_utils.py from /root/reference by file path (pure torch, device agnostic) and
calls normal_from_depth_image exactly as render_normal does (field_construction/gaussian_renderer/__init__.py:28-40), with
autograd for dL/ddepth.  Run in the build container (the reference tree does not exist on the GPU box)."""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_graphics_utils", os.path.join(REF, "field_construction", "utils", "graphics_utils.py"))
gu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gu)

out = {}
for name, (H, W, fx, fy, cx, cy, seed) in {"a": (37, 53, 61.0, 59.5, 26.5, 18.5, 0), "b": (64, 48, 40.0, 40.0, 23.3, 31.9, 1)}.items():
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = (3.0 + 0.02 * xs - 0.015 * ys + 0.3 * torch.rand(H, W, generator=g)).requires_grad_(True)  # tilted plane + noise
    K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
    E = torch.eye(4)
    normal = gu.normal_from_depth_image(depth, K, E).permute(2, 0, 1)        # what render_normal returns
    alpha = torch.rand(1, H, W, generator=g)
    upstream = torch.randn(3, H, W, generator=g)
    (normal * alpha).backward(upstream)                                       # call site: * rendered_alpha.detach()
    out.update({f"{name}_depth": depth.detach().numpy(), f"{name}_intr": np.array([fx, fy, cx, cy], np.float32),
                f"{name}_alpha": alpha.numpy(), f"{name}_upstream": upstream.numpy(),
                f"{name}_normal": normal.detach().numpy(), f"{name}_normal_alpha": (normal * alpha).detach().numpy(),
                f"{name}_g_depth": depth.grad.numpy()})
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "depth_normal.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), {k: v.shape for k, v in out.items()})
