#!/usr/bin/env python
"""Generates tests/golden/pose_l1.npz by executing the REFERENCE's own pure-torch helpers on CPU:
  field_construction/utils/pose_utils.py   get_camera_from_tensor (:60-87), quadmultiply (:89-107), composed exactly like the
                                           camera_pose branch of render() (field_construction/gaussian_renderer/__init__.py:79-87,
                                           re-typed here because render() needs the rasterizer; `.cuda()` -> CPU)
  field_construction/utils/loss_utils.py   l1_loss (:20-21), called like field_construction/gaussian_field.py:450-451
Gradients come from autograd through that code."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(REF, "field_construction"))     # the reference imports `utils.*` relative to this directory
import types  # noqa: E402
_stepfun = types.ModuleType("utils.stepfun")     # utils/stepfun.py is absent from the reference tree; only the camera-path
_stepfun.sample = _stepfun.sample_np = None      # generators (pose_utils.py:346,567) use it, not the functions executed here
sys.modules["utils.stepfun"] = _stepfun
pose_utils = importlib.import_module("utils.pose_utils")
loss_utils = importlib.import_module("utils.loss_utils")

out = {}
for name, (P, seed) in {"a": (301, 0), "b": (64, 1)}.items():
    g = torch.Generator().manual_seed(seed)
    pose = torch.cat([torch.randn(4, generator=g) * 1.4, torch.randn(3, generator=g)]).requires_grad_(True)   # quaternion NOT unit
    xyz = (torch.randn(P, 3, generator=g) * 2).requires_grad_(True)
    rot = (torch.randn(P, 4, generator=g) * 1.2).requires_grad_(True)
    rel_w2c = pose_utils.get_camera_from_tensor(pose)                                    # render(): :79-87
    gaussians_xyz = xyz.clone()
    gaussians_rot = rot.clone()
    xyz_ones = torch.ones(gaussians_xyz.shape[0], 1).float()
    xyz_homo = torch.cat((gaussians_xyz, xyz_ones), dim=1)
    gaussians_xyz_trans = (rel_w2c @ xyz_homo.T).T[:, :3]
    gaussians_rot_trans = pose_utils.quadmultiply(pose[:4], gaussians_rot)
    up_x, up_r = torch.randn(P, 3, generator=g), torch.randn(P, 4, generator=g)
    ((gaussians_xyz_trans * up_x).sum() + (gaussians_rot_trans * up_r).sum()).backward()
    out.update({f"{name}_pose": pose.detach().numpy(), f"{name}_xyz": xyz.detach().numpy(), f"{name}_rot": rot.detach().numpy(),
                f"{name}_means3D": gaussians_xyz_trans.detach().numpy(), f"{name}_rotations": gaussians_rot_trans.detach().numpy(),
                f"{name}_up_means3D": up_x.numpy(), f"{name}_up_rotations": up_r.numpy(),
                f"{name}_g_pose": pose.grad.numpy(), f"{name}_g_xyz": xyz.grad.numpy(), f"{name}_g_rot": rot.grad.numpy()})
for name, (C, H, W, seed) in {"l3": (3, 37, 53, 2), "l16": (16, 24, 40, 3)}.items():
    g = torch.Generator().manual_seed(seed)
    lf = torch.randn(C, H, W, generator=g).requires_grad_(True)
    gt = torch.randn(C, H, W, generator=g)
    gt[:, ::5, ::3] = lf.detach()[:, ::5, ::3]                      # exact ties: sign(0) = 0
    seg = torch.randint(-1, 4, (H, W), generator=g)
    mask = seg != -1                                                # Camera.get_language_feature (cameras.py:143-145)
    loss = loss_utils.l1_loss(lf * mask, gt * mask)                 # gaussian_field.py:450-451
    (loss * 1.7).backward()
    out.update({f"{name}_lf": lf.detach().numpy(), f"{name}_gt": gt.numpy(), f"{name}_mask": mask.numpy(),
                f"{name}_loss": loss.detach().numpy(), f"{name}_g_lf": lf.grad.numpy()})
dst = os.path.join(HERE, "..", "tests", "golden", "pose_l1.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), len(out), "arrays", os.path.getsize(dst) // 1024, "KiB")
