"""CPU tier: the oracle (oracle/lsx_oracle.c) is pinned against outputs of the UNMODIFIED reference CUDA code
recorded on a B200 (tests/golden/*.npz, made by oracle/make_golden.py).

The oracle cannot be bit-exact against CUDA (hardware ex2, FMA contraction in the covariance chain), so:
  * integer / index outputs: mismatch COUNTS must stay below a small fraction,
  * float outputs: tolerance stated per check (tensor-scale relative error).
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RAST_FILES = sorted(glob.glob(os.path.join(GOLDEN, "rast_*.npz")))


def _opt(a):
    return None if a.size == 0 else a


def relerr(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / max(float(np.abs(b).max()), 1e-30)


def run_oracle(z):
    geo, feat = bool(z["in_render_geo"]), bool(z["in_include_feature"])
    f = orc.rasterize_forward(
        z["in_means3D"], z["in_opacities"], z["in_viewmatrix"], z["in_projmatrix"], z["in_campos"], int(z["in_W"]),
        int(z["in_H"]), float(z["in_tanfovx"]), float(z["in_tanfovy"]), z["in_bg"], shs=_opt(z["in_sh"]),
        sh_degree=int(z["in_sh_degree"]), colors_precomp=_opt(z["in_colors_precomp"]), scales=_opt(z["in_scales"]),
        rotations=_opt(z["in_rotations"]), cov3D_precomp=_opt(z["in_cov3D_precomp"]),
        scale_modifier=float(z["in_scale_modifier"]), language_feature=z["in_language_feature"] if feat else None,
        instance_feature=z["in_instance_feature"] if feat else None, all_map=_opt(z["in_all_map_in"]),
        include_feature=feat, render_geo=geo)
    b = orc.rasterize_backward(f, z["gin_color"], z["gin_language_feature"] if feat else None,
                               z["gin_instance_feature"] if feat else None, z["gin_all_map"], z["gin_plane_depth"])
    return f, b


def test_golden_files_present():
    assert len(RAST_FILES) >= 4, "tests/golden/rast_*.npz missing (generate with oracle/make_golden.py on a GPU box)"
    assert os.path.exists(os.path.join(GOLDEN, "knn.npz"))


@pytest.mark.parametrize("path", RAST_FILES, ids=[os.path.basename(p)[5:-4] for p in RAST_FILES])
def test_oracle_matches_reference(path):
    z = np.load(path)
    f, b = run_oracle(z)
    P = z["in_means3D"].shape[0]
    # --- binning: integer outputs -----------------------------------------------------------------------
    radii_bad = int((f["radii"] != z["ref_radii"]).sum())
    assert radii_bad <= max(1, P // 500), f"{radii_bad} radii differ"
    if radii_bad == 0 and int(z["num_rendered"]) == f["num_rendered"]:
        assert (f["tiles_touched"] == z["ref_tiles_touched"].astype(np.uint32)).all()
        # sorted keys: tile ids must match exactly; depth bits may differ in the last ulps
        rk, ok = z["ref_keys"].astype(np.uint64), f["keys"]
        assert ((rk >> np.uint64(32)) == (ok >> np.uint64(32))).all()
        assert (f["ranges"] == z["ref_ranges"].astype(np.uint32)).all()
        pl_bad = int((f["point_list"] != z["ref_point_list"].astype(np.uint32)).sum())
        assert pl_bad <= max(2, f["num_rendered"] // 1000), f"{pl_bad} point-list entries differ (depth ties/ulps)"
        nc_bad = int((f["n_contrib"] != z["ref_n_contrib"].astype(np.uint32)).sum())
        assert nc_bad <= max(2, f["n_contrib"].size // 200), f"{nc_bad} n_contrib differ (threshold flips)"
    vis = z["ref_radii"] > 0
    # --- per-Gaussian floats: 1e-5 of tensor scale --------------------------------------------------------
    for k in ["depths", "means2D", "conic_opacity", "cov3D"]:
        if k == "cov3D" and z["in_cov3D_precomp"].size:
            continue
        assert relerr(f[k][vis], z["ref_" + k][vis]) < 1e-5, k
    if z["in_sh"].size:
        assert relerr(f["rgb"][vis], z["ref_rgb"][vis]) < 1e-5
        assert int((f["clamped"][vis] != z["ref_clamped"][vis]).sum()) <= 1
    # --- images: compare pixels whose blend list agrees (threshold flips change a pixel completely) --------
    same = (f["n_contrib"] == z["ref_n_contrib"].astype(np.uint32)).reshape(int(z["in_H"]), int(z["in_W"]))
    assert same.mean() > 0.99
    for k in ["color", "language_feature", "instance_feature", "all_map"]:
        ref = z["ref_" + k]
        if ref.ndim == 3:
            assert relerr(f[k][:, same], ref[:, same]) < 1e-4, k  # glibc expf vs CUDA ex2.approx path
    if bool(z["in_render_geo"]):
        # plane depth is ill-conditioned where the blended normal is orthogonal to the ray: compare where it is not
        am = z["ref_all_map"]
        ok = same & (np.abs(z["ref_plane_depth"][0]) < 50) & (am[3] > 0.5)
        assert relerr(f["plane_depth"][0][ok], z["ref_plane_depth"][0][ok]) < 1e-3
    # --- gradients: 2e-3 of tensor scale (thresholds flip for a few pixels; reference sums with fp32 atomics) -----
    for k in ["means2D", "means2D_abs", "colors", "opacity", "means3D", "cov3D", "sh", "scales", "rotations", "all_map",
              "language_feature", "instance_feature"]:
        ref = z["refgrad_" + k]
        if ref.size <= 1 or not np.abs(ref).max() > 0:
            continue
        assert relerr(b[k].reshape(ref.shape), ref) < 2e-3, k


def test_knn_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "knn.npz"))
    for k in [n[4:] for n in z.files if n.startswith("pts_")]:
        got, ref = orc.knn_mean_dist2(z["pts_" + k]), z["ref_" + k]
        assert got.shape == ref.shape
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"knn case {k}: not bit-exact"
