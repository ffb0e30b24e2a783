#!/usr/bin/env python
"""Generates tests/golden/densify.npz by executing the REFERENCE's own densification code on CPU.

field_construction/scene/gaussian_model.py is imported from /root/reference (absent third-party modules stubbed, exactly as
oracle/make_golden_gaussian_head.py does) and these methods run UNMODIFIED on CPU tensors:
    add_densification_stats (:720-724), densify_and_prune (:700-718) -> densify_and_clone (:664-698), densify_and_split
    (:612-662), densification_postfix / cat_tensors_to_optimizer (:561-610), prune_points / _prune_optimizer (:520-559),
    reset_opacity (:443-446) / replace_tensor_to_optimizer (:506-518), with a real torch.optim.Adam holding non-zero moments.
Two things are substituted so that the code can run here and be compared value by value:
  * `device="cuda"` arguments of torch.zeros are redirected to the CPU (the reference hard-codes the device);
  * torch.normal(mean=0, std=stds) is replaced by `stds * z` with z ~ N(0, 1) drawn from a recorded generator — the same
    distribution; the unit noise is stored so that the restatement and the CUDA kernels consume identical samples.
The max_radii2D update of field_construction/gaussian_field.py:521-523 (two lines of the training loop) is re-typed inline.
"""
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

# ---- substitutions (see the docstring) ---------------------------------------------------------------------------
_zeros = torch.zeros


def _zeros_cpu(*a, **k):
    if str(k.get("device", "cpu")).startswith("cuda"):
        k["device"] = "cpu"
    return _zeros(*a, **k)


torch.zeros = _zeros_cpu
NOISE = {"gen": None, "log": []}


def _normal(mean=None, std=None, **k):
    z = torch.randn(std.shape, generator=NOISE["gen"])
    NOISE["log"].append(z)
    return mean + std * z


torch.normal = _normal

GROUPS = ("xyz", "knn_f", "f_dc", "f_rest", "opacity", "scaling", "rotation", "language_feature", "instance_feature")
ATTR = {"xyz": "_xyz", "knn_f": "_knn_f", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation", "language_feature": "_language_feature",
        "instance_feature": "_instance_feature"}

CASES = {
    # name: P, seed, extent, percent_dense, max_grad, abs_max_grad, min_opacity, max_screen_size, max_all_points,
    #       abs_split_radii2D_threshold, max_abs_split_points, F
    "plain": dict(P=400, seed=0, extent=5.0, pd=0.01, max_grad=0.0055, abs_grad=0.006, min_op=0.05, size=20, max_all=10**7,
                  abs_r=20, max_abs=0, F=3),
    "abs": dict(P=450, seed=1, extent=5.0, pd=0.01, max_grad=0.006, abs_grad=0.005, min_op=0.05, size=None, max_all=10**7,
                abs_r=20, max_abs=10**6, F=16),
    "abs_capped": dict(P=450, seed=2, extent=5.0, pd=0.01, max_grad=0.006, abs_grad=0.005, min_op=0.05, size=20, max_all=10**7,
                       abs_r=20, max_abs=12, F=3),
    "clone_capped": dict(P=400, seed=3, extent=5.0, pd=0.01, max_grad=0.004, abs_grad=0.006, min_op=0.02, size=20, max_all=430,
                         abs_r=20, max_abs=0, F=3),
    "split_capped": dict(P=400, seed=4, extent=5.0, pd=0.01, max_grad=0.004, abs_grad=0.006, min_op=0.02, size=20, max_all=505,
                         abs_r=20, max_abs=0, F=3),
    "nothing": dict(P=120, seed=5, extent=5.0, pd=0.01, max_grad=1.0, abs_grad=1.0, min_op=0.0, size=None, max_all=10**7,
                    abs_r=20, max_abs=0, F=3),
}


def make_model(c):
    P, F = c["P"], c["F"]
    g = torch.Generator().manual_seed(c["seed"])
    r = lambda *s: torch.randn(*s, generator=g)
    gm = object.__new__(GaussianModel)
    gm.setup_functions()
    gm._xyz = torch.nn.Parameter(r(P, 3) * 2)
    gm._knn_f = torch.nn.Parameter(r(P, 6))
    gm._features_dc = torch.nn.Parameter(r(P, 1, 3) * 0.5)
    gm._features_rest = torch.nn.Parameter(r(P, 15, 3) * 0.1)
    gm._opacity = torch.nn.Parameter(r(P, 1) * 2.0)
    sc = r(P, 3) * 0.7 + float(np.log(c["pd"] * c["extent"]))
    sc[::37] += 3.0                                   # a few world-size outliers (> 0.1 * extent)
    gm._scaling = torch.nn.Parameter(sc)
    gm._rotation = torch.nn.Parameter(r(P, 4) * 1.3)  # not normalised
    gm._language_feature = torch.nn.Parameter(r(P, F))
    gm._instance_feature = torch.nn.Parameter(r(P, 3))
    gm.percent_dense = c["pd"]
    gm.abs_split_radii2D_threshold = c["abs_r"]
    gm.max_abs_split_points = c["max_abs"]
    gm.max_all_points = c["max_all"]
    z = lambda *s: torch.zeros(*s)
    gm.xyz_gradient_accum, gm.xyz_gradient_accum_abs, gm.denom, gm.denom_abs = z(P, 1), z(P, 1), z(P, 1), z(P, 1)
    gm.max_radii2D, gm.max_weight = z(P), z(P)
    lrs = {"xyz": 1.6e-4, "knn_f": 0.01, "f_dc": 2.5e-3, "f_rest": 1.25e-4, "opacity": 0.05, "scaling": 5e-3, "rotation": 1e-3,
           "language_feature": 2.5e-3, "instance_feature": 2.5e-3}
    gm.optimizer = torch.optim.Adam([{"params": [getattr(gm, ATTR[n])], "lr": lrs[n], "name": n} for n in GROUPS],
                                    lr=0.0, eps=1e-15)
    for _ in range(2):                                # two real steps -> non-zero moments
        for n in GROUPS:
            p = getattr(gm, ATTR[n])
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        gm.optimizer.step()
    gm.optimizer.zero_grad(set_to_none=True)
    return gm, g


def snapshot(gm, prefix, out):
    for n in GROUPS:
        p = getattr(gm, ATTR[n])
        st = gm.optimizer.state.get(gm.optimizer.param_groups[GROUPS.index(n)]["params"][0], None)
        assert gm.optimizer.param_groups[GROUPS.index(n)]["params"][0] is p
        out[f"{prefix}_{n}"] = p.detach().reshape(p.shape[0], -1).numpy().copy()
        out[f"{prefix}_m_{n}"] = st["exp_avg"].reshape(p.shape[0], -1).numpy().copy()
        out[f"{prefix}_v_{n}"] = st["exp_avg_sq"].reshape(p.shape[0], -1).numpy().copy()
    out[f"{prefix}_grad_accum"] = gm.xyz_gradient_accum.reshape(-1).numpy().copy()
    out[f"{prefix}_grad_accum_abs"] = gm.xyz_gradient_accum_abs.reshape(-1).numpy().copy()
    out[f"{prefix}_denom"] = gm.denom.reshape(-1).numpy().copy()
    out[f"{prefix}_max_radii2D"] = gm.max_radii2D.reshape(-1).numpy().copy()


out = {}
for name, c in CASES.items():
    gm, g = make_model(c)
    P = c["P"]
    # ---- per-view statistics: three "views" through the reference's own add_densification_stats ----
    for v in range(3):
        radii = torch.randint(0, 45, (P,), generator=g, dtype=torch.int32)
        radii[torch.rand(P, generator=g) < 0.3] = 0
        observe = torch.randint(0, 3, (P,), generator=g, dtype=torch.int32)
        vs = types.SimpleNamespace(grad=torch.randn(P, 3, generator=g) * 0.004)
        vsa = types.SimpleNamespace(grad=torch.randn(P, 3, generator=g).abs() * 0.006)
        visibility_filter = radii > 0
        mask = (observe > 0) & visibility_filter                                             # gaussian_field.py:521-523
        gm.max_radii2D[mask] = torch.max(gm.max_radii2D[mask], radii[mask])
        gm.add_densification_stats(vs, vsa, visibility_filter)
        out.update({f"{name}_view{v}_radii": radii.numpy(), f"{name}_view{v}_observe": observe.numpy(),
                    f"{name}_view{v}_g2d": vs.grad.numpy(), f"{name}_view{v}_g2d_abs": vsa.grad.numpy()})
    snapshot(gm, f"{name}_in", out)
    NOISE["gen"], NOISE["log"] = torch.Generator().manual_seed(1000 + c["seed"]), []
    gm.densify_and_prune(c["max_grad"], c["abs_grad"], c["min_op"], c["extent"], c["size"])
    snapshot(gm, f"{name}_out", out)
    zs = [z.numpy() for z in NOISE["log"]]
    # densify_and_clone draws only when something was selected; densify_and_split always draws (possibly 0 rows)
    if len(zs) == 1:
        zs = [np.zeros((0, 3), np.float32)] + zs
    out[f"{name}_z_clone"], out[f"{name}_z_split"] = zs
    out[f"{name}_cfg"] = np.array([c["extent"], c["pd"], c["max_grad"], c["abs_grad"], c["min_op"],
                                   -1.0 if c["size"] is None else c["size"], c["max_all"], c["abs_r"], c["max_abs"]], np.float64)
    print(f"{name}: P {P} -> {gm._xyz.shape[0]}  (clone noise {zs[0].shape[0]}, split noise {zs[1].shape[0]})")
    # ---- reset_opacity on the densified model ----
    gm.reset_opacity()
    out[f"{name}_reset_opacity"] = gm._opacity.detach().numpy().copy()
dst = os.path.join(HERE, "..", "tests", "golden", "densify.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), len(out), "arrays", os.path.getsize(dst) // 1024, "KiB")
