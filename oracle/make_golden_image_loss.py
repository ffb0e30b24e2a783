#!/usr/bin/env python
"""Generates tests/golden/image_loss.npz by running the REFERENCE's own l1_loss / ssim (+ autograd) on CPU.
field_construction/utils/loss_utils.py is pure torch; it is imported by file path from /root/reference."""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_loss_utils", os.path.join(REF, "field_construction", "utils", "loss_utils.py"))
lu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(lu)

out = {}
for name, (H, W, seed, noise) in {"a": (45, 61, 0, 0.15), "b": (32, 80, 1, 0.02), "c": (11, 9, 2, 0.5)}.items():
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    gt = torch.stack([0.5 + 0.4 * torch.sin(6 * xs + c) * torch.cos(4 * ys - c) for c in range(3)]).clamp(0, 1)
    img = (gt + noise * torch.randn(3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    lam = 0.2
    s = lu.ssim(img, gt)
    l1 = lu.l1_loss(img, gt)
    loss = (1.0 - lam) * l1 + lam * (1.0 - s)          # gaussian_field.py:246 with opt.lambda_dssim = 0.2
    loss.backward()
    out.update({f"{name}_img": img.detach().numpy(), f"{name}_gt": gt.numpy(), f"{name}_lambda": np.float32(lam),
                f"{name}_ssim": np.float32(s.item()), f"{name}_l1": np.float32(l1.item()), f"{name}_loss": np.float32(loss.item()),
                f"{name}_g_img": img.grad.numpy()})
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "image_loss.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
