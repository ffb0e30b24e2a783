"""Densification / pruning on flat arenas — host side of csrc/densify.cu.

Mirrors, for parameters held in one flat fp32 arena (+ the two Adam-moment arenas of `ArenaAdam`):
    GaussianModel.add_densification_stats    field_construction/scene/gaussian_model.py:720-724
                                             (+ the max_radii2D update, field_construction/gaussian_field.py:521-523)
    GaussianModel.densify_and_prune          gaussian_model.py:700-718  (clone :664-698, split :612-662, prune :520-559)
    GaussianModel.reset_opacity              gaussian_model.py:443-446
The reference concatenates and masks every parameter tensor and Adam moment group by group (re-creating nn.Parameters each
time); here one planning pass decides the fate of every row and ONE gather per arena builds the new set.  Row order, selection
rules (including the max_all_points / max_abs_split_points quantile branches) and the order in which normal samples are
consumed are the reference's, so with the same unit noise the result is the same set of Gaussians.  No CPU path.
"""
import ctypes
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from . import _lib
from .multiview import DensifyStats, GradArena


@dataclass
class DensifyConfig:
    """training_args fields read by the densification code (configs/field_construction.yaml values as defaults)."""
    percent_dense: float = 0.001
    max_all_points: int = 12_000_000
    abs_split_radii2D_threshold: float = 20
    max_abs_split_points: int = 0


GROUP_ALIGN = 64   # elements


class ParamArena(GradArena):
    """A flat fp32 arena of P rows with arbitrary named groups (the reference's optimizer groups: xyz, knn_f, f_dc, f_rest,
    opacity, scaling, rotation, language_feature, instance_feature — gaussian_model.py:313-323 — or lsx_b200.multiview's)."""

    @staticmethod
    def allocate(P: int, widths: Dict[str, int], device, zero: bool = True) -> "ParamArena":
        offs, total = {}, 0
        for name, w in widths.items():
            n = P * w
            offs[name] = (total, n)
            total += (n + GROUP_ALIGN - 1) // GROUP_ALIGN * GROUP_ALIGN          # 256-B aligned groups, like GradArena
        flat = (torch.zeros if zero else torch.empty)(max(total, 1), dtype=torch.float32, device=device)
        views = {name: flat[o:o + n].view(P, widths[name]) for name, (o, n) in offs.items()}
        a = ParamArena(flat, views, offs)
        a.widths, a.P = dict(widths), P
        return a

    @staticmethod
    def like(arena: GradArena, P: int) -> "ParamArena":
        return ParamArena.allocate(P, arena_widths(arena), arena.flat.device)


def arena_widths(arena: GradArena) -> Dict[str, int]:
    return {name: (v.shape[1] if v.dim() == 2 else 0) for name, v in arena.views.items()}


def arena_rows(arena: GradArena) -> int:
    return next(iter(arena.views.values())).shape[0]


# which group plays which part in the sampling of new rows, for the two naming schemes in use
_ROLE_NAMES = {"xyz": ("xyz", "means3D"), "scaling": ("scaling", "scales"), "rotation": ("rotation", "rotations"),
               "opacity": ("opacity",)}


def _role_groups(widths):
    out = {}
    for role, names in _ROLE_NAMES.items():
        hit = [n for n in names if n in widths]
        if len(hit) != 1:
            raise RuntimeError(f"arena needs exactly one of the groups {names} (has {list(widths)})")
        out[role] = hit[0]
    return out


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def add_densification_stats(stats: DensifyStats, means2D_grad, means2D_abs_grad, radii, out_observe=None):
    """One view: the fused form of `max_radii2D[mask] = max(...)` + add_densification_stats(...)."""
    if not stats.denom.is_cuda:
        raise RuntimeError("add_densification_stats needs CUDA tensors (this operator has no CPU path)")
    P = stats.denom.numel()
    g, ga = means2D_grad.contiguous(), means2D_abs_grad.contiguous()
    r = radii.contiguous().to(torch.int32)
    ob = None if out_observe is None else out_observe.contiguous().to(torch.int32)
    if g.shape != (P, 3) or ga.shape != (P, 3) or r.numel() != P or g.dtype != torch.float32:
        raise RuntimeError("add_densification_stats: expected (P,3) float32 screen gradients and P radii")
    dev = stats.denom.device
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsx_densify_stats_update(P, g.data_ptr(), ga.data_ptr(), r.data_ptr(),
                                                        None if ob is None else ob.data_ptr(), stats.grad_accum.data_ptr(),
                                                        stats.grad_accum_abs.data_ptr(), stats.denom.data_ptr(),
                                                        stats.max_radii2D.data_ptr(), _stream(dev)), "densify_stats_update")
    return stats


@dataclass
class DensifyResult:
    params: ParamArena
    exp_avg: Optional[ParamArena]
    exp_avg_sq: Optional[ParamArena]
    stats: DensifyStats
    n_clone: int
    n_split: int
    n_split_abs: int
    n_kept_original: int
    n_kept_clone: int
    n_kept_split: int
    capped: Dict[str, bool]


def densify_and_prune(params: GradArena, exp_avg: Optional[GradArena], exp_avg_sq: Optional[GradArena], stats: DensifyStats,
                      cfg: DensifyConfig, max_grad: float, abs_max_grad: float, min_opacity: float, extent: float,
                      max_screen_size, noise_fn: Callable[[str, int], torch.Tensor] = None, generator=None) -> DensifyResult:
    """GaussianModel.densify_and_prune on arenas.  `noise_fn(kind, n)` returns the (n, 3) unit-normal samples for
    kind = "clone" / "split" (default: torch.randn on the device, optionally from `generator` — identical on every rank of a
    view-sharded job when the generators are seeded alike).  Returns NEW arenas; the inputs are left untouched."""
    if params is None or not params.flat.is_cuda:
        raise RuntimeError("densify_and_prune needs CUDA arenas (this operator has no CPU path)")
    if (exp_avg is None) != (exp_avg_sq is None):
        raise RuntimeError("give both Adam moment arenas or neither")
    dev = params.flat.device
    widths = arena_widths(params)
    roles = _role_groups(widths)
    P = arena_rows(params)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws_bytes = int(lib.lsx_densify_workspace_bytes(P))
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a = _lib.DensifyPlanArgs()
        a.P = P
        a.grad_accum, a.grad_accum_abs = stats.grad_accum.data_ptr(), stats.grad_accum_abs.data_ptr()
        a.denom, a.max_radii2D = stats.denom.data_ptr(), stats.max_radii2D.data_ptr()
        a.scaling_raw = params.views[roles["scaling"]].data_ptr()
        a.opacity_raw = params.views[roles["opacity"]].data_ptr()
        a.max_grad, a.abs_max_grad, a.min_opacity, a.extent = max_grad, abs_max_grad, min_opacity, extent
        a.percent_dense, a.abs_split_radii2D_threshold = cfg.percent_dense, cfg.abs_split_radii2D_threshold
        a.max_all_points, a.max_abs_split_points = int(cfg.max_all_points), int(cfg.max_abs_split_points)
        a.prune_world_size = 1 if max_screen_size else 0
        a.workspace, a.workspace_bytes, a.stream = workspace.data_ptr(), ws_bytes, _stream(dev)
        res = _lib.DensifyPlanResult()
        _lib.check(lib.lsx_densify_plan(ctypes.byref(a), ctypes.byref(res)), "densify_plan")

        P_new = int(res.P_new)
        if noise_fn is None:
            noise_fn = lambda kind, n: torch.randn(n, 3, device=dev, dtype=torch.float32, generator=generator)
        z_clone = noise_fn("clone", int(res.n_clone)).contiguous()
        z_split = noise_fn("split", 2 * int(res.n_split)).contiguous()
        if z_clone.shape != (int(res.n_clone), 3) or z_split.shape != (2 * int(res.n_split), 3) or not z_clone.is_cuda:
            raise RuntimeError("noise_fn must return CUDA tensors of shape (n, 3)")
        # uninitialised: the gather writes every row and zero-fills the alignment padding (group_align)
        new_p = ParamArena.allocate(P_new, widths, dev, zero=P_new == 0)
        new_m = ParamArena.allocate(P_new, widths, dev, zero=False) if exp_avg is not None else None
        new_v = ParamArena.allocate(P_new, widths, dev, zero=False) if exp_avg is not None else None
        names = [n for n, w in widths.items() if w > 0]
        G = len(names)
        ap = _lib.DensifyApplyArgs()
        ap.P_new, ap.n_new_rows, ap.n_groups = P_new, P_new - int(res.n_kept_original), G
        ob = (ctypes.c_int64 * G)(*[params.offsets[n][0] for n in names])
        nb = (ctypes.c_int64 * G)(*[new_p.offsets[n][0] for n in names])
        wd = (ctypes.c_int32 * G)(*[widths[n] for n in names])
        role_of = {roles["xyz"]: "xyz", roles["scaling"]: "scaling", roles["rotation"]: "rotation"}
        rl = (ctypes.c_int32 * G)(*[_lib.DENSIFY_ROLE[role_of.get(n, "copy")] for n in names])
        ap.old_begin, ap.new_begin, ap.width, ap.role = ob, nb, wd, rl
        ap.row_map, ap.noise_index = res.row_map, res.noise_index
        ap.z_clone, ap.z_split = z_clone.data_ptr(), z_split.data_ptr()
        ap.old_params, ap.new_params = params.flat.data_ptr(), new_p.flat.data_ptr()
        if exp_avg is not None:
            if exp_avg.flat.numel() != params.flat.numel() or exp_avg_sq.flat.numel() != params.flat.numel():
                raise RuntimeError("moment arenas and parameter arena have different layouts")
            ap.old_exp_avg, ap.old_exp_avg_sq = exp_avg.flat.data_ptr(), exp_avg_sq.flat.data_ptr()
            ap.new_exp_avg, ap.new_exp_avg_sq = new_m.flat.data_ptr(), new_v.flat.data_ptr()
        ap.stream, ap.group_align = _stream(dev), GROUP_ALIGN
        _lib.check(lib.lsx_densify_apply(ctypes.byref(ap)), "densify_apply")
        workspace.record_stream(torch.cuda.current_stream(dev))
        new_stats = DensifyStats.allocate(P_new, dev)      # densification_postfix: all statistics restart at zero
    return DensifyResult(new_p, new_m, new_v, new_stats, int(res.n_clone), int(res.n_split), int(res.n_split_abs),
                         int(res.n_kept_original), int(res.n_kept_clone), int(res.n_kept_split),
                         {"clone": bool(res.clone_capped), "split": bool(res.split_capped), "abs": bool(res.abs_capped)})


def reset_opacity(params: GradArena, exp_avg: Optional[GradArena] = None, exp_avg_sq: Optional[GradArena] = None):
    """GaussianModel.reset_opacity in place: opacity <- inverse_sigmoid(min(sigmoid(opacity), 0.01)), its moments zeroed."""
    if not params.flat.is_cuda:
        raise RuntimeError("reset_opacity needs CUDA arenas (this operator has no CPU path)")
    op = params.views["opacity"]
    dev = params.flat.device
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsx_reset_opacity(op.shape[0], op.data_ptr(),
                                                 None if exp_avg is None else exp_avg.views["opacity"].data_ptr(),
                                                 None if exp_avg_sq is None else exp_avg_sq.views["opacity"].data_ptr(),
                                                 _stream(dev)), "reset_opacity")
    return params
