"""numpy front-end of the CPU oracle (oracle/lsx_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of lsx_oracle.c.  Never imported by langscene-x_b200/.
`rasterize_forward` / `rasterize_backward` mirror the reference's `_C.rasterize_gaussians[_backward]`
data flow (rasterize_points.cu:35-259) on numpy arrays; `knn_mean_dist2` mirrors distCUDA2.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lsx_oracle.c")
LIB = os.path.join(HERE, "build", "liblsx_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or (os.path.exists(SRC) and os.path.getmtime(SRC) > os.path.getmtime(LIB)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        # -ffp-contract=off: FMAs only where the source says fmaf (they mirror the GPU build's placement)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c99",
                               "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_binning.restype = ctypes.c_int64
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _cf(x):
    return ctypes.c_float(float(x))


def rasterize_forward(means3D, opacities, view, proj, campos, W, H, tanfovx, tanfovy, bg, shs=None, sh_degree=3,
                      colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
                      language_feature=None, instance_feature=None, all_map=None, include_feature=True,
                      render_geo=True):
    """Returns a dict with every output and intermediate of the reference's forward pass."""
    L = lib()
    means3D, opacities, view, proj, campos, bg = map(_f, (means3D, opacities, view, proj, campos, bg))
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(_f, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    lang, inst, all_map = _f(language_feature), _f(instance_feature), _f(all_map)
    P = means3D.shape[0]
    M = shs.shape[1] if shs is not None else 0
    F = lang.shape[1] if (include_feature and lang is not None) else 0
    Fi = inst.shape[1] if (include_feature and inst is not None) else 0
    T = ((W + 15) // 16) * ((H + 15) // 16)
    o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
             cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
             conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
             clamped=np.zeros((P, 3), np.uint8))
    L.orc_preprocess(P, int(sh_degree), M, _p(means3D), _p(scales), _cf(scale_modifier), _p(rotations), _p(opacities),
                     _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(view), _p(proj), _p(campos), W, H, _cf(tanfovx),
                     _cf(tanfovy), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
                     _p(o["conic_opacity"]), _p(o["tiles_touched"]), _p(o["clamped"]))
    R = int(o["tiles_touched"].astype(np.int64).sum())
    o["num_rendered"] = R
    o["keys"] = np.zeros(R, np.uint64)
    o["point_list"] = np.zeros(R, np.uint32)
    o["ranges"] = np.zeros((T, 2), np.uint32)
    R2 = L.orc_binning(P, W, H, _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["keys"]) if R else None,
                       _p(o["point_list"]) if R else None, _p(o["ranges"]))
    assert R2 == R
    colors = colors_precomp if colors_precomp is not None else o["rgb"]
    o["colors"] = colors
    o["color"] = np.zeros((3, H, W), np.float32)
    o["language_feature"] = np.zeros((F, H, W), np.float32) if include_feature else np.zeros(1, np.float32)
    o["instance_feature"] = np.zeros((Fi, H, W), np.float32) if include_feature else np.zeros(1, np.float32)
    o["out_observe"] = np.zeros(P, np.int32)
    o["all_map"] = np.zeros((5, H, W), np.float32)
    o["plane_depth"] = np.zeros((1, H, W), np.float32)
    o["final_T"] = np.zeros(H * W, np.float32)
    o["n_contrib"] = np.zeros(H * W, np.uint32)
    L.orc_render_forward(W, H, _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]), _p(o["conic_opacity"]), _p(colors),
                         _p(lang), _p(inst), _p(all_map), F, Fi, int(include_feature), int(render_geo), _cf(tanfovx),
                         _cf(tanfovy), _p(bg), _p(o["color"]), _p(o["language_feature"]), _p(o["instance_feature"]),
                         _p(o["out_observe"]), _p(o["all_map"]), _p(o["plane_depth"]), _p(o["final_T"]),
                         _p(o["n_contrib"]))
    o["_inputs"] = dict(means3D=means3D, opacities=opacities, view=view, proj=proj, campos=campos, bg=bg, shs=shs,
                        sh_degree=int(sh_degree), scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
                        scale_modifier=float(scale_modifier), lang=lang, inst=inst, all_map=all_map, F=F, Fi=Fi, M=M,
                        W=W, H=H, tanfovx=float(tanfovx), tanfovy=float(tanfovy), include_feature=bool(include_feature),
                        render_geo=bool(render_geo))
    return o


def rasterize_backward(fwd, dL_color, dL_language_feature=None, dL_instance_feature=None, dL_all_map=None,
                       dL_plane_depth=None):
    """Gradients in the reference's return order names (rasterize_points.cu:258)."""
    L = lib()
    i = fwd["_inputs"]
    P, W, H, F, Fi, M = i["means3D"].shape[0], i["W"], i["H"], i["F"], i["Fi"], i["M"]
    g = dict(means2D=np.zeros((P, 3), np.float32), means2D_abs=np.zeros((P, 3), np.float32),
             conic=np.zeros((P, 4), np.float32), opacity=np.zeros((P, 1), np.float32), colors=np.zeros((P, 3), np.float32),
             language_feature=np.zeros((P, max(F, 1)), np.float32), instance_feature=np.zeros((P, max(Fi, 1)), np.float32),
             all_map=np.zeros((P, 5), np.float32), means3D=np.zeros((P, 3), np.float32), cov3D=np.zeros((P, 6), np.float32),
             sh=np.zeros((P, M, 3), np.float32), scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32))
    dL_color, dLF, dLFi, dLA, dLD = map(_f, (dL_color, dL_language_feature, dL_instance_feature, dL_all_map, dL_plane_depth))
    L.orc_render_backward(W, H, _p(fwd["ranges"]), _p(fwd["point_list"]), _p(fwd["means2D"]), _p(fwd["conic_opacity"]),
                          _p(fwd["colors"]), _p(i["lang"]), _p(i["inst"]), _p(i["all_map"]), F, Fi,
                          int(i["include_feature"]), int(i["render_geo"]), _cf(i["tanfovx"]), _cf(i["tanfovy"]), _p(i["bg"]),
                          _p(fwd["all_map"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL_color), _p(dLF), _p(dLFi),
                          _p(dLA), _p(dLD), _p(g["means2D"]), _p(g["means2D_abs"]), _p(g["conic"]), _p(g["opacity"]),
                          _p(g["colors"]), _p(g["language_feature"]), _p(g["instance_feature"]), _p(g["all_map"]))
    cov = i["cov3D_precomp"] if i["cov3D_precomp"] is not None else fwd["cov3D"]
    L.orc_preprocess_backward(P, i["sh_degree"], M, _p(i["means3D"]), _p(fwd["radii"]), _p(i["shs"]), _p(fwd["clamped"]),
                              _p(i["scales"]), _p(i["rotations"]), _cf(i["scale_modifier"]), _p(cov), _p(i["view"]),
                              _p(i["proj"]), _p(i["campos"]), W, H, _cf(i["tanfovx"]), _cf(i["tanfovy"]), _p(g["means2D"]),
                              _p(g["conic"]), _p(g["colors"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["sh"]),
                              _p(g["scales"]), _p(g["rotations"]))
    if not i["include_feature"]:
        g["language_feature"] = np.zeros(1, np.float32)
        g["instance_feature"] = np.zeros(1, np.float32)
    return g


def mark_visible(means3D, view):
    means3D, view = _f(means3D), _f(view)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().orc_mark_visible(means3D.shape[0], _p(means3D), _p(view), _p(out))
    return out.astype(bool)


def knn_mean_dist2(points):
    points = _f(points)
    out = np.zeros(points.shape[0], np.float32)
    lib().orc_knn_mean_dist2(points.shape[0], _p(points), _p(out))
    return out
