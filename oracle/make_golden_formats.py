#!/usr/bin/env python
"""Generates tests/golden/formats.npz and tests/golden/checkpoint_ref.pth by executing the REFERENCE's own readers / capture on CPU:
  read_camera_npz                 field_construction/scene/dataset_readers.py:234-296   on synthetic camera/*.npz files
  Camera.get_language_feature     field_construction/scene/cameras.py:137-151           on synthetic <name>_f.npy / _s.npy files
  GaussianModel.capture           field_construction/scene/gaussian_model.py:90-137     with a real torch.optim.Adam state
The synthetic input files are stored inside the golden archive so that the tests can re-create them.
(plyfile is absent here, so save_ply / load_ply cannot be executed: the PLY layout is pinned by a restated byte-level KAT in
tests/test_io_formats.py and by the round trip.)"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "field_construction"))


class _Any:
    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


for n in ("plyfile", "open3d", "simple_knn", "simple_knn._C", "pytorch3d", "pytorch3d.transforms"):
    m = types.ModuleType(n)
    m.__dict__.update(PlyData=_Any(), PlyElement=_Any(), distCUDA2=None, quaternion_to_matrix=None)
    sys.modules[n] = m
st = types.ModuleType("utils.stepfun")
st.sample = st.sample_np = None
sys.modules["utils.stepfun"] = st
dr = importlib.import_module("field_construction.scene.dataset_readers")
cams = importlib.import_module("field_construction.scene.cameras")
GaussianModel = importlib.import_module("field_construction.scene.gaussian_model").GaussianModel

out = {}
rng = np.random.default_rng(0)
with tempfile.TemporaryDirectory() as d:
    cam_dir = os.path.join(d, "camera")
    os.makedirs(cam_dir)
    file_names = ["00003.npz", "00000.npz", "frame_00007.npz"]
    for i, fn in enumerate(file_names):
        A = rng.normal(size=(3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        pose = np.eye(4)
        pose[:3, :3], pose[:3, 3] = Q, rng.normal(size=3)
        K = np.array([[500.0 + i, 0, 360.0], [0, 510.0 + i, 240.0], [0, 0, 1]])
        np.savez(os.path.join(cam_dir, fn), pose=pose, intrinsics=K)
        out[f"cam_in_{i}_pose"], out[f"cam_in_{i}_K"] = pose, K
    out["cam_file_names"] = np.array(file_names)
    images, cameras = dr.read_camera_npz(cam_dir)
    ids = sorted(images)
    out["cam_ids"] = np.array(ids)
    out["cam_names"] = np.array([images[i].name for i in ids])
    out["cam_qvec"] = np.stack([images[i].qvec for i in ids])
    out["cam_tvec"] = np.stack([images[i].tvec for i in ids])
    out["cam_params"] = np.stack([cameras[i].params for i in ids])
    out["cam_wh"] = np.array([[cameras[i].width, cameras[i].height] for i in ids])

    lf_dir = os.path.join(d, "lang")
    os.makedirs(lf_dir)
    fmap = rng.normal(size=(3, 30, 45)).astype(np.float32)
    seg = rng.integers(-1, 5, size=(48, 72)).astype(np.int32)
    np.save(os.path.join(lf_dir, "00003_f.npy"), fmap)
    np.save(os.path.join(lf_dir, "00003_s.npy"), seg)
    cam = object.__new__(cams.Camera)
    torch.nn.Module.__init__(cam)
    cam.image_name, cam.data_device, cam.image_height, cam.image_width = "00003", "cpu", 48, 72
    feat, mask, seg_out = cam.get_language_feature(lf_dir)
    out.update({"lf_in_f": fmap, "lf_in_s": seg, "lf_feat": feat.numpy(), "lf_mask": mask.numpy(), "lf_seg": seg_out.numpy()})

# ---- checkpoint tuple ----
P, F = 40, 3
g = torch.Generator().manual_seed(3)
r = lambda *s: torch.randn(*s, generator=g)
gm = object.__new__(GaussianModel)
gm.setup_functions()
gm.active_sh_degree = 2
gm._xyz, gm._knn_f = torch.nn.Parameter(r(P, 3)), torch.nn.Parameter(r(P, 6))
gm._features_dc, gm._features_rest = torch.nn.Parameter(r(P, 1, 3)), torch.nn.Parameter(r(P, 15, 3))
gm._opacity, gm._scaling, gm._rotation = torch.nn.Parameter(r(P, 1)), torch.nn.Parameter(r(P, 3)), torch.nn.Parameter(r(P, 4))
gm._language_feature, gm._instance_feature = torch.nn.Parameter(r(P, F)), torch.nn.Parameter(r(P, 3))
gm.max_radii2D, gm.max_weight = torch.rand(P, generator=g) * 30, torch.zeros(P)
gm.xyz_gradient_accum, gm.xyz_gradient_accum_abs = torch.rand(P, 1, generator=g), torch.rand(P, 1, generator=g)
gm.denom = torch.randint(0, 5, (P, 1), generator=g).float()
gm.denom_abs = gm.denom.clone()
gm.spatial_lr_scale = 2.5
gm.P = torch.nn.Parameter(r(4, 7))
names = ("xyz", "knn_f", "f_dc", "f_rest", "opacity", "scaling", "rotation", "language_feature", "instance_feature")
attr = {"xyz": "_xyz", "knn_f": "_knn_f", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation", "language_feature": "_language_feature", "instance_feature": "_instance_feature"}
gm.optimizer = torch.optim.Adam([{"params": [getattr(gm, attr[n])], "lr": 1e-3 * (i + 1), "name": n} for i, n in enumerate(names)],
                                lr=0.0, eps=1e-15)
gm.cam_optimizer = torch.optim.Adam([{"params": [gm.P], "lr": 1e-4, "name": "pose"}], lr=0.0, eps=1e-15)
for _ in range(3):
    for n in names:
        p = getattr(gm, attr[n])
        p.grad = torch.randn(p.shape, generator=g) * 0.01
    gm.P.grad = torch.randn(gm.P.shape, generator=g) * 0.01
    gm.optimizer.step()
    gm.cam_optimizer.step()
ckpt = (gm.capture(include_feature=True), 1234)
torch.save(ckpt, os.path.join(HERE, "..", "tests", "golden", "checkpoint_ref.pth"))
for n in names:
    p = getattr(gm, attr[n])
    stt = gm.optimizer.state[p]
    out[f"ck_{n}"] = p.detach().reshape(P, -1).numpy().copy()
    out[f"ck_m_{n}"], out[f"ck_v_{n}"] = stt["exp_avg"].reshape(P, -1).numpy().copy(), stt["exp_avg_sq"].reshape(P, -1).numpy().copy()
out["ck_grad_accum"], out["ck_denom"] = gm.xyz_gradient_accum.reshape(-1).numpy(), gm.denom.reshape(-1).numpy()
out["ck_max_radii2D"] = gm.max_radii2D.numpy()
dst = os.path.join(HERE, "..", "tests", "golden", "formats.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), len(out), "arrays", os.path.getsize(dst) // 1024, "KiB;",
      os.path.getsize(os.path.join(HERE, "..", "tests", "golden", "checkpoint_ref.pth")) // 1024, "KiB checkpoint")
