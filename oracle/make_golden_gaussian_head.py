#!/usr/bin/env python
"""Generates tests/golden/gaussian_head.npz by executing the REFERENCE's own GaussianModel methods on CPU.

field_construction/scene/gaussian_model.py is imported from /root/reference with its absent / GPU-only dependencies stubbed
(pytorch3d.transforms.quaternion_to_matrix <- the restatement in oracle/gaussian_head_oracle.py; simple_knn, plyfile: unused
here).  get_scaling / get_rotation / get_opacity / get_normal are the reference's code; the six all_map lines of render()
(field_construction/gaussian_renderer/__init__.py:188-196) are restated inline because render() cannot run without the
rasterizer.  Gradients come from autograd through that code."""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gaussian_head_oracle import quaternion_to_matrix  # noqa: E402

REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


class _Any:
    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


_stub("pytorch3d")
_stub("pytorch3d.transforms", quaternion_to_matrix=quaternion_to_matrix)
_stub("simple_knn")
_stub("simple_knn._C", distCUDA2=None)
for n in ("plyfile", "open3d"):
    try:
        __import__(n)
    except Exception:
        _stub(n, PlyData=_Any(), PlyElement=_Any())
GaussianModel = importlib.import_module("field_construction.scene.gaussian_model").GaussianModel

out = {}
for name, (P, seed, yaw) in {"a": (257, 0, 0.3), "b": (64, 1, -1.1)}.items():
    g = torch.Generator().manual_seed(seed)
    gm = object.__new__(GaussianModel)
    gm.setup_functions()
    gm._xyz = (torch.randn(P, 3, generator=g) * 2).requires_grad_(True)
    gm._scaling = (torch.randn(P, 3, generator=g) - 3).requires_grad_(True)
    gm._rotation = (torch.randn(P, 4, generator=g) * 1.7).requires_grad_(True)      # NOT normalised
    gm._opacity = torch.randn(P, 1, generator=g).requires_grad_(True)
    c, s = np.cos(yaw), np.sin(yaw)
    w2c = torch.eye(4)
    w2c[:3, :3] = torch.tensor([[c, 0, s], [0.1 * s, 1, -0.1 * c], [-s, 0.1, c]], dtype=torch.float32)
    w2c[:3, :3] = torch.linalg.qr(w2c[:3, :3])[0]
    w2c[:3, 3] = torch.tensor([0.3, -0.2, 4.0])
    cam = types.SimpleNamespace(world_view_transform=w2c.t().contiguous(), camera_center=torch.linalg.inv(w2c)[:3, 3].contiguous())
    scales, rots, opac = gm.get_scaling, gm.get_rotation, gm.get_opacity
    global_normal = gm.get_normal(cam)
    local_normal = global_normal @ cam.world_view_transform[:3, :3]                         # render(): :188-196
    pts_in_cam = gm.get_xyz @ cam.world_view_transform[:3, :3] + cam.world_view_transform[3, :3]
    local_distance = (local_normal * pts_in_cam).sum(-1).abs()
    all_map = torch.zeros((P, 5))
    all_map[:, :3] = local_normal
    all_map[:, 3] = 1.0
    all_map[:, 4] = local_distance
    ups = [torch.randn(t.shape, generator=g) for t in (scales, rots, opac, all_map)]
    (scales * ups[0]).sum().add((rots * ups[1]).sum()).add((opac * ups[2]).sum()).add((all_map * ups[3]).sum()).backward()
    out.update({f"{name}_xyz": gm._xyz.detach().numpy(), f"{name}_scaling": gm._scaling.detach().numpy(),
                f"{name}_rotation": gm._rotation.detach().numpy(), f"{name}_opacity": gm._opacity.detach().numpy(),
                f"{name}_view": cam.world_view_transform.numpy(), f"{name}_campos": cam.camera_center.numpy(),
                f"{name}_scales": scales.detach().numpy(), f"{name}_rotations": rots.detach().numpy(),
                f"{name}_opac": opac.detach().numpy(), f"{name}_all_map": all_map.detach().numpy(),
                f"{name}_up_scales": ups[0].numpy(), f"{name}_up_rotations": ups[1].numpy(), f"{name}_up_opac": ups[2].numpy(),
                f"{name}_up_all_map": ups[3].numpy(),
                f"{name}_g_xyz": gm._xyz.grad.numpy(), f"{name}_g_scaling": gm._scaling.grad.numpy(),
                f"{name}_g_rotation": gm._rotation.grad.numpy(), f"{name}_g_opacity": gm._opacity.grad.numpy()})
dst = os.path.join(HERE, "..", "tests", "golden", "gaussian_head.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), len(out), "arrays")
