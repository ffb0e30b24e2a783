"""Shared helpers for the GPU parity tests, smoke() and bench.py's reference legs.

* loads the UNMODIFIED reference (oracle/_ref/*.so, built by oracle/build_ref.py) — test infrastructure,
  never imported by the product package;
* runs the reference and the new operator on identical inputs;
* parses the reference's private scratch buffers and ours (lsx_scratch_layout_query)
  so that radii / tile keys / sorted lists / tile ranges / n_contrib can be compared bit for bit.
"""
import ctypes
import importlib.util
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "langscene-x_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

REF_DIR = os.path.join(REPO, "oracle", "_ref")
_ref_cache = {}


def load_ref(name):
    """name in {ref_rast_f3, ref_rast_f16, ref_knn}; returns the pybind module or None if not built."""
    if name in _ref_cache:
        return _ref_cache[name]
    path = os.path.join(REF_DIR, name + ".so")
    mod = None
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _ref_cache[name] = mod
    return mod


def ref_rast_for(F):
    return load_ref("ref_rast_f16" if F == 16 else "ref_rast_f3")


def _al(x, a=128):
    return (x + a - 1) // a * a


def _view(buf, off, count, dtype):
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return buf[off:off + nbytes].view(dtype)


def parse_ref_buffers(geom, binning, img, P, R, W, H):
    """Byte layouts of GeometryState / BinningState / ImageState (rasterizer_impl.cu:155-194)."""
    out = {}
    N = W * H
    T = ((W + 15) // 16) * ((H + 15) // 16)
    # image
    o_n = _al(4 * N)
    o_r = _al(o_n + 4 * N)
    out["final_T"] = _view(img, 0, N, torch.float32)
    out["n_contrib"] = _view(img, o_n, N, torch.int32)
    out["ranges"] = _view(img, o_r, 2 * T, torch.int32).view(T, 2)
    # binning
    if R > 0:
        o1 = _al(4 * R)
        o2 = _al(o1 + 4 * R)
        out["point_list"] = _view(binning, 0, R, torch.int32)
        out["keys"] = _view(binning, o2, R, torch.int64)
    # geometry
    o = 0
    out["depths"] = _view(geom, o, P, torch.float32); o = _al(o + 4 * P)
    out["clamped"] = _view(geom, o, 3 * P, torch.uint8).view(P, 3); o = _al(o + 3 * P)
    o = _al(o + 4 * P)  # internal_radii
    out["means2D"] = _view(geom, o, 2 * P, torch.float32).view(P, 2); o = _al(o + 8 * P)
    out["cov3D"] = _view(geom, o, 6 * P, torch.float32).view(P, 6); o = _al(o + 24 * P)
    out["conic_opacity"] = _view(geom, o, 4 * P, torch.float32).view(P, 4); o = _al(o + 16 * P)
    out["rgb"] = _view(geom, o, 3 * P, torch.float32).view(P, 3); o = _al(o + 12 * P)
    out["tiles_touched"] = _view(geom, o, P, torch.int32)
    return out


def parse_new_buffers(geom, binning, img, P, R, W, H, n_blend):
    from lsx_b200 import _lib
    lib = _lib.load()
    lay = _lib.ScratchLayout()
    # the binning scratch is carved for a list CAPACITY >= R (R rounded up to 64, or the speculative hint of the call)
    cap = int(lib.lsx_binning_capacity(int(binning.numel()), W, H)) if R > 0 else 0
    assert cap >= R, (cap, R)
    _lib.check(lib.lsx_scratch_layout_query(P, W, H, cap, n_blend, ctypes.byref(lay)), "layout")
    N = W * H
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out = {}
    out["final_T"] = _view(img, lay.final_T, N, torch.float32)
    out["n_contrib"] = _view(img, lay.n_contrib, N, torch.int32)
    out["ranges"] = _view(img, lay.ranges, 2 * T, torch.int32).view(T, 2)
    out["depths"] = _view(geom, lay.depths, P, torch.float32)
    cl = _view(geom, lay.clamped, P, torch.uint8)
    out["clamped"] = torch.stack([(cl >> k) & 1 for k in range(3)], dim=1)
    out["means2D"] = _view(geom, lay.means2D, 2 * P, torch.float32).view(P, 2)
    out["cov3D"] = _view(geom, lay.cov3D, 6 * P, torch.float32).view(P, 6)
    out["conic_opacity"] = _view(geom, lay.conic_opacity, 4 * P, torch.float32).view(P, 4)
    out["rgb"] = _view(geom, lay.rgb, 3 * P, torch.float32).view(P, 3)
    out["tiles_touched"] = _view(geom, lay.tiles_touched, P, torch.int32)
    if R > 0:
        out["point_list"] = _view(binning, lay.point_list, R, torch.int32)
        out["masks"] = _view(binning, lay.masks, R, torch.uint8)
        out["blk_list"] = _view(binning, lay.blk_list, 8 * cap, torch.int32).view(8, cap)
        out["blk_cnt"] = _view(binning, lay.blk_cnt, 8 * T, torch.int32).view(T, 8)
        out["k_contrib"] = _view(img, lay.k_contrib, N, torch.int32)
        keys = torch.empty(R, dtype=torch.int64, device=geom.device)
        _lib.check(lib.lsx_debug_sorted_keys(P, W, H, R, n_blend, geom.data_ptr(), binning.data_ptr(), int(binning.numel()),
                                             img.data_ptr(), keys.data_ptr(), torch.cuda.current_stream().cuda_stream), "debug keys")
        out["keys"] = keys
    return out


def native_forward_args(scene, cam, bg, F, sh_degree=3, render_geo=True, include_feature=True, use_sh=True,
                        all_map=None, scale_modifier=1.0, cov3D_precomp=None):
    """The 24 positional arguments of `_C.rasterize_gaussians` (rasterize_points.cu:35-60)."""
    dev = scene.means3D.device
    empty = torch.Tensor([])
    if use_sh:
        colors, sh = empty, scene.shs
    else:
        colors, sh = torch.sigmoid(scene.shs[:, 0, :]).contiguous(), empty
    lang = scene.language_feature if include_feature else torch.zeros((1,), device=dev)
    inst = scene.instance_feature if include_feature else torch.zeros((1,), device=dev)
    if render_geo and all_map is None:
        from lsx_b200.synthetic import make_all_map
        all_map = make_all_map(scene, cam)
    if not render_geo:
        all_map = empty
    if cov3D_precomp is None:
        scales, rots, cov = scene.scales, scene.rotations, empty
    else:
        scales, rots, cov = empty, empty, cov3D_precomp
    return [bg, scene.means3D, colors, lang, inst, scene.opacities, scales, rots, float(scale_modifier), cov, all_map,
            cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.H, cam.W, sh, sh_degree, cam.campos,
            False, bool(render_geo), False, bool(include_feature)]


FWD_NAMES = ["num_rendered", "color", "language_feature", "instance_feature", "radii", "out_observe", "all_map",
             "plane_depth", "geom", "binning", "img"]
BWD_NAMES = ["means2D", "means2D_abs", "colors", "language_feature", "instance_feature", "opacity", "means3D", "cov3D",
             "sh", "scales", "rotations", "all_map"]


def native_backward_args(fargs, fwd, grads):
    """The 31 positional arguments of `_C.rasterize_gaussians_backward` (rasterize_points.cu:145-177)."""
    (bg, means3D, colors, lang, inst, opac, scales, rots, smod, cov, all_map, view, proj, tfx, tfy, H, W, sh, deg,
     campos, _pref, render_geo, debug, include_feature) = fargs
    dev = means3D.device
    one = torch.zeros((1,), device=dev)
    g_lang = grads["language_feature"] if include_feature else one
    g_inst = grads["instance_feature"] if include_feature else one
    return [bg, fwd["all_map"], means3D, fwd["radii"], colors, lang, inst, all_map, scales, rots, smod, cov, view, proj,
            tfx, tfy, grads["color"], g_lang, g_inst, grads["all_map"], grads["plane_depth"], sh, deg, campos,
            fwd["geom"], fwd["num_rendered"], fwd["binning"], fwd["img"], render_geo, debug, include_feature]


def run_native(mod, fargs, grads=None):
    """mod = reference pybind module or lsx_b200.ops (same signatures).  Returns (fwd dict, bwd dict or None)."""
    fwd = dict(zip(FWD_NAMES, mod.rasterize_gaussians(*fargs)))
    bwd = None
    if grads is not None:
        bwd = dict(zip(BWD_NAMES, mod.rasterize_gaussians_backward(*native_backward_args(fargs, fwd, grads))))
    return fwd, bwd


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): tensor-scale relative error."""
    a, b = a.double(), b.double()
    denom = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom
