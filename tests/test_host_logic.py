"""CPU tier: the C-ABI library loads and exports every symbol the header declares, the ctypes argument blocks
match the C structs byte for byte, the drop-in Python surface validates like the reference, and the product
refuses to run without CUDA (no CPU fallback).  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

import harness as hz

HEADER = os.path.join(hz.REPO, "include", "lsx_rasterizer.h")


def test_library_exports_every_declared_symbol():
    from lsx_b200 import _lib
    lib = _lib.load()
    declared = set(re.findall(r"LSX_API\s+[\w\s\*]+?\b(lsx_\w+)\s*\(", open(HEADER).read()))
    assert declared, "no LSX_API declarations found in the header"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lsx_abi_version() == _lib.ABI_VERSION
    assert _lib.kernel_launch_count() == 0  # nothing ran


def test_ctypes_structs_match_c_layout():
    """Compile a tiny C program against the header and compare sizeof / offsetof with ctypes."""
    from lsx_b200 import _lib
    structs = {"lsx_forward_args": _lib.ForwardArgs, "lsx_backward_args": _lib.BackwardArgs,
               "lsx_scratch_layout": _lib.ScratchLayout, "lsx_densify_plan_args": _lib.DensifyPlanArgs,
               "lsx_densify_plan_result": _lib.DensifyPlanResult, "lsx_densify_apply_args": _lib.DensifyApplyArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-std=c99", "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_scratch_layout_query_is_pure_host_and_aligned():
    from lsx_b200 import _lib
    lib = _lib.load()
    lay = _lib.ScratchLayout()
    assert lib.lsx_scratch_layout_query(1000, 1920, 1080, 5000, 27, ctypes.byref(lay)) == 0
    offs = [lay.depths, lay.clamped, lay.means2D, lay.cov3D, lay.conic_opacity, lay.rgb, lay.tiles_touched, lay.records,
            lay.final_T, lay.n_contrib, lay.ranges, lay.point_list]
    assert all(o % 256 == 0 for o in offs)
    assert lay.record_stride == 40 and lay.record_stride % 8 == 0       # 8 head + 28 channels -> 160-B records
    assert lay.geom_bytes > lay.records + 1000 * lay.record_stride * 4
    assert lib.lsx_scratch_layout_query(10, 64, 64, 0, 100, ctypes.byref(lay)) != 0  # too many channels
    assert "bad arguments" in _lib.last_error()


def test_settings_tuple_and_validation_errors():
    from diff_LangSurf_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix", "sh_degree",
        "campos", "prefiltered", "render_geo", "debug", "include_feature")
    eye = torch.eye(4)
    s = GaussianRasterizationSettings(32, 32, 0.5, 0.5, torch.zeros(3), 1.0, eye, eye, 0, torch.zeros(3), False, False, False, False)
    r = GaussianRasterizer(s)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=x, means2D_abs=x, opacities=torch.ones(4, 1), scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=x, means2D_abs=x, opacities=torch.ones(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=x,
          scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, means2D_abs=x, opacities=torch.ones(4, 1), colors_precomp=x, scales=x)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, means2D_abs=x, opacities=torch.ones(4, 1), colors_precomp=x, scales=x,
          rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    """CPU tensors are an error, never a silent slow path."""
    from diff_LangSurf_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from simple_knn._C import distCUDA2
    eye = torch.eye(4)
    s = GaussianRasterizationSettings(32, 32, 0.5, 0.5, torch.zeros(3), 1.0, eye, eye, 0, torch.zeros(3), False, False, False, False)
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        GaussianRasterizer(s)(means3D=x, means2D=x, means2D_abs=x, opacities=torch.ones(4, 1), colors_precomp=x, scales=x,
                              rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        GaussianRasterizer(s)(means3D=torch.zeros(4, 2), means2D=x, means2D_abs=x, opacities=torch.ones(4, 1), colors_precomp=x,
                              scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        distCUDA2(torch.rand(10, 3))


def test_missing_library_fails_loudly():
    code = ("import sys; sys.path.insert(0, %r); import os; os.environ['LSX_B200_LIB']='/nonexistent/liblsx_b200.so';"
            "from lsx_b200 import _lib\ntry:\n    _lib.load()\nexcept ImportError as e:\n    print('IMPORTERROR', 'no CPU fallback' in str(e))"
            % os.path.join(hz.REPO, "langscene-x_b200"))
    out = subprocess.check_output([sys.executable, "-c", code], text=True)
    assert "IMPORTERROR True" in out


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under langscene-x_b200/ may reference it."""
    pkg = os.path.join(hz.REPO, "langscene-x_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "lsx_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_synthetic_generator_is_deterministic_and_matches_spec():
    from lsx_b200.synthetic import make_all_map, make_camera, make_cameras, make_scene
    a, b = make_scene(5000, 640, 360, F=16, seed=3), make_scene(5000, 640, 360, F=16, seed=3)
    for f in a.__dataclass_fields__:
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert a.shs.shape == (5000, 16, 3) and a.language_feature.shape == (5000, 16) and a.opacities.shape == (5000, 1)
    assert torch.allclose(a.rotations.norm(dim=1), torch.ones(5000), atol=1e-5)
    near = (a.means3D[:, 2] < 0.2).float().mean().item()
    assert 0.03 < near < 0.07                                  # ~5 % of the points exercise the near cull
    cams = make_cameras(640, 360, 5)
    assert len(cams) == 5 and abs(cams[0].tanfovy - cams[0].tanfovx * 360 / 640) < 1e-9
    c0 = make_camera(640, 360)
    assert torch.allclose(c0.viewmatrix, torch.eye(4)) and torch.allclose(c0.campos, torch.zeros(3), atol=1e-6)
    assert torch.allclose(c0.projmatrix, c0.viewmatrix @ c0.projmatrix)   # identity view: full proj == P^T
    am = make_all_map(a, cams[2])
    assert am.shape == (5000, 5) and torch.all(am[:, 3] == 1) and torch.all(am[:, 4] >= 0)
    assert torch.allclose(am[:, :3].norm(dim=1), torch.ones(5000), atol=1e-5)


def test_oracle_small_properties():
    """Oracle self-checks that need no golden data: sortedness, range partition, alpha bookkeeping, KNN vs numpy."""
    import numpy as np
    from oracle import oracle as orc
    from lsx_b200.synthetic import make_all_map, make_camera, make_scene
    sc, cam = make_scene(800, 80, 60, F=3, seed=5), make_camera(80, 60, yaw_deg=4.0)
    n = lambda t: t.numpy()
    o = orc.rasterize_forward(n(sc.means3D), n(sc.opacities), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos), 80, 60,
                              cam.tanfovx, cam.tanfovy, [0.1, 0.2, 0.3], shs=n(sc.shs), scales=n(sc.scales),
                              rotations=n(sc.rotations), language_feature=n(sc.language_feature),
                              instance_feature=n(sc.instance_feature), all_map=n(make_all_map(sc, cam)))
    k = o["keys"]
    assert (k[1:] >= k[:-1]).all() and int((o["ranges"][:, 1] - o["ranges"][:, 0]).sum()) == o["num_rendered"]
    assert np.abs(o["all_map"][3].reshape(-1) + o["final_T"] - 1).max() < 1e-4
    assert (o["out_observe"] >= 0).all() and o["out_observe"].sum() > 0
    assert np.array_equal(orc.mark_visible(n(sc.means3D), n(cam.viewmatrix)),
                          (n(sc.means3D) @ n(cam.viewmatrix)[:3, :3] + n(cam.viewmatrix)[3, :3])[:, 2] > 0.2)
    pts = np.random.RandomState(1).randn(300, 3).astype(np.float32)
    d2 = ((pts[:, None, :].astype(np.float64) - pts[None]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    ref = np.sort(d2, axis=1)[:, :3].mean(1)
    assert np.allclose(orc.knn_mean_dist2(pts), ref, rtol=1e-5)


def test_module_autograd_plumbing_with_a_stub_native_layer(monkeypatch):
    """The nn.Module / autograd.Function layer on CPU tensors with the native `_C` functions stubbed out: argument counts,
    the 7-tuple, which inputs receive gradients, and the multi-view gradient sink (sunk groups get None from autograd)."""
    import types
    import diff_LangSurf_rasterization as D
    P, H, W = 5, 4, 4
    seen = {}

    def fwd(*a):
        assert len(a) == 24
        z = lambda *s: torch.zeros(*s)
        return (3, z(3, H, W) + a[1].sum(), z(3, H, W) + a[3].sum(), z(3, H, W), torch.ones(P, dtype=torch.int32),
                torch.ones(P, dtype=torch.int32), z(5, H, W) + a[10].sum(), z(1, H, W) + a[5].sum(), z(1), z(1), z(1))

    def bwd(*a, grad_buffers=None, accumulate=False):
        assert len(a) == 31
        seen["sink"] = (grad_buffers, accumulate)
        o = lambda *s: torch.ones(*s)
        return (o(P, 3), o(P, 3), o(P, 3), o(P, 3), o(P, 3), o(P, 1), o(P, 3), o(P, 6), o(P, 16, 3), o(P, 3), o(P, 4), o(P, 5))

    monkeypatch.setattr(D, "_C", types.SimpleNamespace(rasterize_gaussians=fwd, rasterize_gaussians_backward=bwd))
    leaf = lambda *s: torch.randn(*s, requires_grad=True)
    s = D.GaussianRasterizationSettings(H, W, 1., 1., torch.zeros(3), 1., torch.eye(4), torch.eye(4), 3, torch.zeros(3),
                                        False, True, False, True)
    for sink in (None, {"means3D": torch.zeros(P, 3), "sh": torch.zeros(P, 48), "opacity": torch.zeros(P, 1)}):
        p = dict(m3=leaf(P, 3), sh=leaf(P, 16, 3), lf=leaf(P, 3), li=leaf(P, 3), op=leaf(P, 1), sc=leaf(P, 3), ro=leaf(P, 4),
                 am=leaf(P, 5))
        m2 = torch.zeros(P, 3, requires_grad=True)
        out = D.GaussianRasterizer(s, grad_buffers=sink, accumulate=True)(
            means3D=p["m3"], means2D=m2, means2D_abs=m2, opacities=p["op"], shs=p["sh"], language_feature_precomp=p["lf"],
            language_feature_instance_precomp=p["li"], scales=p["sc"], rotations=p["ro"], all_map=p["am"])
        assert len(out) == 7
        (out[0].sum() + out[1].sum() + out[5].sum() + out[6].sum()).backward()
        sunk = {"m3", "sh", "op"} if sink else set()
        for k, v in p.items():
            assert (v.grad is None) == (k in sunk), k
        assert m2.grad is not None                      # screen-space gradients always reach autograd (densification)
        assert (seen["sink"][0] is sink) and seen["sink"][1] == bool(sink)


def test_binning_capacity_is_recovered_from_the_buffer_size():
    """The backward pass and the debug helpers get only the binning buffer: its list capacity (a multiple of 64) must be
    recoverable from its size, for every capacity and image size (pure host arithmetic)."""
    from lsx_b200 import _lib
    lib = _lib.load()
    lay = _lib.ScratchLayout()
    for W, H in ((64, 64), (720, 480), (1920, 1080)):
        for cap in (0, 64, 128, 4096, 1_000_000 // 64 * 64, 26_500_032):
            assert lib.lsx_scratch_layout_query(10, W, H, cap, 27, ctypes.byref(lay)) == 0
            assert lib.lsx_binning_capacity(lay.binning_bytes, W, H) == cap, (W, H, cap)
            assert lib.lsx_binning_capacity(lay.binning_bytes + 8, W, H) == -1        # not a size any capacity produces


def test_reference_arm_imports_do_not_load_the_product_library():
    """`bench.py --impl reference` must not map liblsx_b200.so (the driver records which in-tree libraries a process loaded):
    everything that arm imports before it starts timing — bench, harness, the pure-torch arena / scene modules, the loop
    bench with its reference-style arm — is imported in a fresh interpreter and /proc/self/maps is inspected."""
    code = (
        "import sys\n"
        f"sys.path[:0] = [{os.path.join(hz.REPO, 'langscene-x_b200')!r}, {hz.REPO!r}, {os.path.join(hz.REPO, 'tests')!r}]\n"
        "import torch, harness, bench, bench_loop\n"
        "from lsx_b200.multiview import GradArena\n"
        "from lsx_b200.synthetic import CONFIGS\n"
        "from lsx_b200.field_loop import LoopConfig, View\n"
        "GradArena.allocate(10, 16, 3, 3, 'cpu')\n"
        "print('LOADED' if 'liblsx_b200' in open('/proc/self/maps').read() else 'CLEAN')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("CLEAN"), out.stdout
