#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/*.so) on a B200.

TEST INFRASTRUCTURE ONLY.  Run on a GPU box:  python oracle/make_golden.py   (writes gpurun_out/golden/,
copy the files into tests/golden/ and commit them).  Each file holds the seeded inputs AND every output /
bit-exact intermediate of the reference for one small case, so that the CPU oracle can be pinned against
the real reference in the GPU-less test tier (tests/test_oracle_golden.py) and the CUDA path can be checked
against the same vectors (tests/test_parity_gpu.py) even where oracle/_ref is unavailable.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import harness as hz  # noqa: E402
from lsx_b200.synthetic import make_camera, make_scene, make_all_map, make_upstream_grads  # noqa: E402

CASES = {
    # name: dict(P, W, H, F, seed, yaw, sh_degree, render_geo, include_feature, use_sh, cov_precomp, bg)
    "full_f3": dict(P=1500, W=96, H=84, F=3, seed=11, yaw=0.0, deg=3, geo=True, feat=True, sh=True, cov=False, bg=(0.0, 0.0, 0.0)),
    "full_f16_yaw": dict(P=1200, W=80, H=64, F=16, seed=12, yaw=9.0, deg=3, geo=True, feat=True, sh=True, cov=False, bg=(0.2, 0.5, 0.1)),
    "rgb_only_deg1": dict(P=1500, W=96, H=80, F=3, seed=13, yaw=-6.0, deg=1, geo=False, feat=False, sh=True, cov=False, bg=(1.0, 1.0, 1.0)),
    "precomp": dict(P=1000, W=64, H=64, F=3, seed=14, yaw=0.0, deg=0, geo=True, feat=False, sh=False, cov=True, bg=(0.0, 0.0, 0.0)),
}


def cov_from_scene(scene):
    """Sigma = (S R)^T (S R) packed upper-triangular, computed in float64 on the host (input to the precomp case)."""
    s = scene.scales.double().cpu()
    q = scene.rotations.double().cpu()
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
    L = R @ torch.diag_embed(s)
    Sig = L @ L.transpose(1, 2)
    return torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], dim=1).float()


def run_case(name, c, outdir):
    dev = torch.device("cuda:0")
    scene = make_scene(c["P"], c["W"], c["H"], F=c["F"], seed=c["seed"]).to(dev)
    cam = make_camera(c["W"], c["H"], yaw_deg=c["yaw"]).to(dev)
    bg = torch.tensor(c["bg"], dtype=torch.float32, device=dev)
    grads = make_upstream_grads(c["W"], c["H"], c["F"], seed=c["seed"] + 100, device=dev)
    cov = cov_from_scene(scene).to(dev) if c["cov"] else None
    all_map = make_all_map(scene, cam) if c["geo"] else None
    fargs = hz.native_forward_args(scene, cam, bg, c["F"], sh_degree=c["deg"], render_geo=c["geo"],
                                   include_feature=c["feat"], use_sh=c["sh"], all_map=all_map, cov3D_precomp=cov)
    ref = hz.ref_rast_for(c["F"])
    fwd, bwd = hz.run_native(ref, fargs, grads)
    torch.cuda.synchronize()
    R = fwd["num_rendered"]
    buf = hz.parse_ref_buffers(fwd["geom"], fwd["binning"], fwd["img"], c["P"], R, c["W"], c["H"])
    vis = (fwd["radii"] > 0)
    out = {}
    # inputs
    names = ["bg", "means3D", "colors_precomp", "language_feature", "instance_feature", "opacities", "scales", "rotations",
             "scale_modifier", "cov3D_precomp", "all_map_in", "viewmatrix", "projmatrix", "tanfovx", "tanfovy", "H", "W", "sh",
             "sh_degree", "campos", "prefiltered", "render_geo", "debug", "include_feature"]
    for n, v in zip(names, fargs):
        out["in_" + n] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    for k, v in grads.items():
        out["gin_" + k] = v.cpu().numpy()
    # reference outputs
    out["num_rendered"] = np.asarray(R)
    for k in ["color", "language_feature", "instance_feature", "radii", "out_observe", "all_map", "plane_depth"]:
        out["ref_" + k] = fwd[k].cpu().numpy()
    for k in ["final_T", "n_contrib", "ranges", "tiles_touched"]:
        out["ref_" + k] = buf[k].cpu().numpy()
    if R > 0:
        out["ref_keys"] = buf["keys"].cpu().numpy()
        out["ref_point_list"] = buf["point_list"].cpu().numpy()
    m = vis.cpu().numpy()
    for k in ["depths", "means2D", "conic_opacity", "rgb", "cov3D", "clamped"]:
        a = buf[k].cpu().numpy().copy()
        a[~m] = 0  # rows of culled Gaussians are uninitialised memory in the reference
        out["ref_" + k] = a
    for k, v in bwd.items():
        out["refgrad_" + k] = v.cpu().numpy()
    path = os.path.join(outdir, f"rast_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "R =", R, "P_vis =", int(vis.sum()), "->", path, os.path.getsize(path) // 1024, "KiB")


def run_knn(outdir):
    dev = torch.device("cuda:0")
    knn = hz.load_ref("ref_knn")
    g = torch.Generator().manual_seed(21)
    cases = {
        "uniform": torch.rand(3000, 3, generator=g) * 4 - 2,
        "clustered": torch.cat([torch.randn(1500, 3, generator=g) * 0.05 + 1.0, torch.randn(1500, 3, generator=g) * 2.0,
                                torch.zeros(4, 3), torch.ones(3, 3) * 0.25]),  # exact duplicates included
        "tiny3": torch.rand(3, 3, generator=g),
        "plane": torch.cat([torch.rand(2500, 2, generator=g), torch.zeros(2500, 1)], dim=1),
    }
    out = {}
    for k, pts in cases.items():
        pts = pts.float().contiguous()
        d = knn.distCUDA2(pts.to(dev))
        torch.cuda.synchronize()
        out["pts_" + k] = pts.numpy()
        out["ref_" + k] = d.cpu().numpy()
    path = os.path.join(outdir, "knn.npz")
    np.savez_compressed(path, **out)
    print("knn ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    outdir = os.path.join(REPO, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for n, c in CASES.items():
        run_case(n, c, outdir)
    run_knn(outdir)
