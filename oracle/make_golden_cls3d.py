#!/usr/bin/env python
"""Generates tests/golden/cls3d.npz by executing the REFERENCE's own loss_cls_3d
(field_construction/utils/loss_utils.py:158-186, pure torch) on CPU, called like field_construction/gaussian_field.py:461-465:
features = xyz.detach(), predictions = the feature parameter; gradients from autograd through that code.
The function draws its rows with torch.randperm on the global CPU generator; the script seeds it, replays the same draws to
record the selected rows, and asserts that the neighbour sets the reference's cdist / topk found equal the exact
(float64) nearest neighbours, i.e. that the vectors are not in the regime where cdist's matmul rounding decides."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(REF, "field_construction"))
loss_utils = importlib.import_module("utils.loss_utils")

CASES = {  # name: (N, C, k, lambda, max_points, sample_size, seed)
    "a": (3000, 3, 5, 2.0, 200000, 800, 11),      # the call pattern of gaussian_field.py:462-464 (k = 5, 800 samples)
    "b": (1200, 3, 3, 1.0, 700, 100, 12),         # down-sampling branch
    "c": (500, 1, 8, 0.5, 200000, 50, 13),        # 1-d predictions, k = 8
    "d": (400, 3, 5, 2.0, 200000, 800, 14),       # fewer points than samples; constant predictions (max == min: no normalisation)
}
out = {}
for name, (N, C, k, lam, max_points, sample_size, seed) in CASES.items():
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(N, 3, generator=g) * 2 - 1
    pred = torch.randn(N, C, generator=g) * 0.7
    if name == "a":          # tied extrema: the gradient through min / max is split evenly (torch's full-reduction backward)
        pred[5, 0] = pred[7, 1] = pred.min() - 0.1
        pred[9, 2] = pred[11, 0] = pred.max() + 0.1
    if name == "d":
        pred[:] = 0.3
    pred_in = (pred[:, 0] if C == 1 else pred).clone().requires_grad_(True)
    torch.manual_seed(seed)
    loss = loss_utils.loss_cls_3d(xyz, pred_in, k, lam, max_points, sample_size)
    (loss * 1.3).backward()
    # replay the draws (loss_utils.py:161,172)
    torch.manual_seed(seed)
    down = torch.randperm(N)[:max_points] if N > max_points else None
    n = N if down is None else max_points
    samples = torch.randperm(n)[:sample_size]
    pts = xyz if down is None else xyz[down]
    # exact neighbours (float64) vs what cdist + topk select
    d2 = ((pts[samples].double()[:, None, :] - pts.double()[None, :, :]) ** 2).sum(-1)
    exact = torch.sort(d2, dim=1, stable=True).indices[:, :k]
    ref_nbr = torch.cdist(pts[samples], pts).topk(k, largest=False).indices
    assert torch.equal(exact.sort(dim=1).values, ref_nbr.sort(dim=1).values), name
    gap = (torch.sort(d2, dim=1).values[:, k] - torch.sort(d2, dim=1).values[:, k - 1]).min()
    out.update({f"{name}_xyz": xyz.numpy(), f"{name}_pred": pred_in.detach().numpy(), f"{name}_loss": loss.detach().numpy(),
                f"{name}_g_pred": pred_in.grad.numpy(), f"{name}_samples": samples.numpy(), f"{name}_nbr": exact.numpy(),
                f"{name}_down": (down.numpy() if down is not None else np.zeros(0, np.int64)),
                f"{name}_args": np.array([k, lam, max_points, sample_size, seed, 1.3], np.float64)})
    print(name, "loss", float(loss), "min gap between the k-th and (k+1)-th squared distance", float(gap))
dst = os.path.join(HERE, "..", "tests", "golden", "cls3d.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), len(out), "arrays", os.path.getsize(dst) // 1024, "KiB")
