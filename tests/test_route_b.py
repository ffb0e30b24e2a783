"""INTEGRATION.md Route B: the reference's C++ glue re-bound on the C ABI (langscene-x_b200/integration/rasterize_points_lsx.cpp).
CPU tier: it compiles against torch's headers, links with liblsx_b200.so and exports the reference's four pybind names.
GPU tier: driven with the reference's positional arguments it returns what the ctypes route (lsx_b200.ops) returns."""
import importlib.util

import pytest
import torch

import harness as hz


def _load():
    from lsx_b200 import _build
    path = _build.build_route_b()
    spec = importlib.util.spec_from_file_location("lsx_route_b", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_route_b_glue_compiles_and_exports_the_reference_names():
    mod = _load()
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible", "distCUDA2"):   # RAST/ext.cpp:15-19, KNN/ext.cpp:15-17
        assert callable(getattr(mod, name)), name
    with pytest.raises(RuntimeError, match="num_points, 3"):                  # the reference's own argument check (rasterize_points.cu:62-64)
        e = torch.Tensor([])
        mod.rasterize_gaussians(e, torch.zeros(4, 2), e, e, e, e, e, e, 1.0, e, e, e, e, 1.0, 1.0, 8, 8, e, 0, e, False, True, False, True)


@pytest.mark.gpu
def test_route_b_equals_the_ctypes_route():
    from lsx_b200 import ops
    from lsx_b200.synthetic import make_camera, make_scene, make_upstream_grads
    mod = _load()
    dev = "cuda:0"
    P, W, H, F = 9_000, 176, 112, 16
    scene = make_scene(P, W, H, F=F, seed=3).to(dev)
    cam = make_camera(W, H, yaw_deg=-3.0).to(dev)
    grads = make_upstream_grads(W, H, F, seed=4, device=dev)
    fargs = hz.native_forward_args(scene, cam, torch.tensor([0.2, 0.1, 0.0], device=dev), F)
    a_f, a_b = hz.run_native(ops, fargs, grads)
    b_f, b_b = hz.run_native(mod, fargs, grads)
    assert a_f["num_rendered"] == b_f["num_rendered"] > 0
    for k in ("color", "language_feature", "instance_feature", "radii", "out_observe", "all_map", "plane_depth"):
        assert torch.equal(a_f[k], b_f[k]), k                                 # same library, same kernels: bit-identical forward
    for k in hz.BWD_NAMES:
        assert a_b[k].shape == b_b[k].shape and hz.rel_err(b_b[k], a_b[k]) < 1e-4, k
    assert torch.equal(ops.mark_visible(scene.means3D, cam.viewmatrix, cam.projmatrix),
                       mod.mark_visible(scene.means3D, cam.viewmatrix, cam.projmatrix))
    pts = scene.means3D[:5000].contiguous()
    assert torch.equal(ops.distCUDA2(pts), mod.distCUDA2(pts))
    e = torch.Tensor([])                                                      # P = 0: the reference returns its zero outputs
    z = mod.rasterize_gaussians(torch.zeros(3, device=dev), torch.zeros(0, 3, device=dev), e, torch.zeros(0, F, device=dev),
                                torch.zeros(0, 3, device=dev), torch.zeros(0, 1, device=dev), torch.zeros(0, 3, device=dev),
                                torch.zeros(0, 4, device=dev), 1.0, e, torch.zeros(0, 5, device=dev), cam.viewmatrix, cam.projmatrix,
                                cam.tanfovx, cam.tanfovy, H, W, torch.zeros(0, 16, 3, device=dev), 3, cam.campos, False, True, False, True)
    assert z[0] == 0 and float(z[1].abs().max()) == 0.0
