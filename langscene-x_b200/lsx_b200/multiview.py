"""View-sharded multi-GPU optimisation support — a NEW capability, the reference is single-GPU.

One process per GPU.  Every rank holds a full replica of the Gaussian parameters and renders the views
`{v : v mod N == rank}` of a batch; per-Gaussian parameter gradients of the local views are summed into ONE
flat fp32 arena (the layout the rasterizer's backward already produces), which is all-reduced with a single
NCCL call over NVLink/NVSwitch.  Densification statistics keep the reference's per-view semantics
(field_construction/scene/gaussian_model.py:720-724, field_construction/gaussian_field.py:523-524):
sums of per-view gradient NORMS and visibility counts, max of radii.

The host-side logic here is device-agnostic so it can be exercised with the gloo backend on CPU tensors.
"""
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

# per-Gaussian parameter-gradient groups, in arena order (name, trailing shape as a function of (M, F, Fi)).
# `all_map` is NOT a parameter: the render wrapper derives it per view from rotation / scale / position and the camera
# (field_construction/gaussian_renderer/__init__.py:188-196), so its gradient is consumed by the wrapper's backward on the
# rank that rendered the view and never crosses ranks (summing it over cameras would be meaningless).
PARAM_GROUPS = ("means3D", "sh", "opacity", "scales", "rotations", "language_feature", "instance_feature")


def shard_views(n_views: int, world_size: int, rank: int) -> List[int]:
    """Views rendered by `rank`: round-robin so that neighbouring (similar-cost) views spread over ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_views, world_size))


def group_widths(M: int, F: int, Fi: int) -> Dict[str, int]:
    return {"means3D": 3, "sh": 3 * M, "opacity": 1, "scales": 3, "rotations": 4, "language_feature": F,
            "instance_feature": Fi}


@dataclass
class GradArena:
    """Flat accumulation buffer for the parameter gradients of P Gaussians."""
    flat: torch.Tensor
    views: Dict[str, torch.Tensor]
    offsets: Dict[str, Tuple[int, int]] = None   # group name -> (first element, element count) inside `flat`

    @staticmethod
    def allocate(P: int, M: int, F: int, Fi: int, device, extra: Dict[str, Tuple[int, ...]] = None) -> "GradArena":
        """`extra`: further groups that are not per-Gaussian, name -> shape — e.g. {"pose": (n_views, 7)} for the camera-pose
        gradients of pose optimisation (GaussianModel.P, field_construction/scene/gaussian_model.py:243-262): every rank
        writes the rows of the views it rendered, the same all-reduce sums them."""
        widths = group_widths(M, F, Fi)
        offs, shapes, total = {}, {}, 0
        for name in PARAM_GROUPS:
            shapes[name] = (P, widths[name])
        for name, shp in (extra or {}).items():
            if name in shapes:
                raise ValueError(f"extra group {name!r} collides with a parameter group")
            shapes[name] = tuple(int(x) for x in shp)
        for name, shp in shapes.items():
            n = 1
            for x in shp:
                n *= x
            offs[name] = (total, n)
            total += (n + 63) // 64 * 64  # 256-B aligned groups (float4 stores, NCCL-friendly)
        flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        views = {name: flat[o:o + n].view(shapes[name]) for name, (o, n) in offs.items()}
        return GradArena(flat, views, offs)

    def zero_(self):
        self.flat.zero_()
        return self

    def accumulate(self, grads: Dict[str, torch.Tensor]):
        """Add one view's gradients (any tensors reshapeable to (P, width)); missing / empty groups are skipped."""
        for name, v in self.views.items():
            g = grads.get(name)
            if g is None or g.numel() != v.numel() or v.numel() == 0:
                continue
            v.add_(g.reshape(v.shape))
        return self

    def grad_buffers(self) -> Dict[str, torch.Tensor]:
        """The arena's groups keyed like `ops.rasterize_gaussians_backward(..., grad_buffers=, accumulate=True)` expects:
        the backward kernel then adds each view's parameter gradients straight into the arena."""
        return {name: v for name, v in self.views.items() if v.numel() > 0}

    def all_reduce(self, group=None, async_op=False):
        """Sum over ranks with ONE collective on the flat buffer."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    def span(self, names: Sequence[str]) -> torch.Tensor:
        """The contiguous slice of `flat` covering the given groups (they must be adjacent in arena order)."""
        order = list(self.offsets)
        idx = sorted(order.index(n) for n in names)
        if idx != list(range(idx[0], idx[0] + len(idx))):
            raise ValueError(f"groups {list(names)} are not adjacent in the arena")
        first = self.offsets[order[idx[0]]][0]
        last_o, last_n = self.offsets[order[idx[-1]]]
        end = (last_o + last_n + 63) // 64 * 64 if idx[-1] + 1 < len(order) else self.flat.numel()
        return self.flat[first:end]

    def all_reduce_spans(self, spans: Sequence[Sequence[str]], group=None) -> "PendingReduce":
        """Asynchronous all-reduce of the arena in pieces: one collective per span of adjacent groups, issued now in the
        given order on the process group's communication stream (NCCL: its own stream, which waits for the work already
        enqueued on the current stream), so that compute enqueued AFTER this call overlaps the transfers.  The caller
        issues a span as soon as the last kernel that writes it has been enqueued, and calls `.wait()` on the result
        before the first kernel that reads the reduced values (the optimiser step)."""
        works = []
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for names in spans:
                works.append(dist.all_reduce(self.span(names), op=dist.ReduceOp.SUM, group=group, async_op=True))
        return PendingReduce(works)


class PendingReduce:
    """Handles of collectives in flight; `wait()` makes the CURRENT stream wait for them (no host block with NCCL)."""

    def __init__(self, works=()):
        self.works = list(works)

    def extend(self, other: "PendingReduce"):
        self.works.extend(other.works)
        return self

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


@dataclass
class DensifyStats:
    """xyz_gradient_accum, xyz_gradient_accum_abs, denom (sums) and max_radii2D (max), per Gaussian."""
    grad_accum: torch.Tensor
    grad_accum_abs: torch.Tensor
    denom: torch.Tensor
    max_radii2D: torch.Tensor

    @staticmethod
    def allocate(P: int, device) -> "DensifyStats":
        z = lambda: torch.zeros(P, dtype=torch.float32, device=device)
        return DensifyStats(z(), z(), z(), z())

    def add_view(self, means2D_grad, means2D_abs_grad, radii, out_observe=None):
        """Per-view update, exactly the reference's: norms of the (x,y) screen gradients of visible Gaussians."""
        visible = radii > 0
        self.grad_accum += torch.where(visible, means2D_grad[:, :2].norm(dim=-1), torch.zeros_like(self.grad_accum))
        self.grad_accum_abs += torch.where(visible, means2D_abs_grad[:, :2].norm(dim=-1), torch.zeros_like(self.grad_accum))
        self.denom += visible.to(self.denom.dtype)
        mask = visible if out_observe is None else (visible & (out_observe > 0))
        self.max_radii2D = torch.where(mask, torch.maximum(self.max_radii2D, radii.to(self.max_radii2D.dtype)),
                                       self.max_radii2D)

    def all_reduce(self, group=None):
        """Sum / max over ranks, IN PLACE.  Only ever call this on a per-step DELTA (statistics of the views of one step,
        started from zero): the persistent statistics are cumulative over a whole densification interval
        (gaussian_model.py:720-724), so reducing them again every step would count earlier steps world_size times.
        `multiview_step` does the right thing: delta -> all_reduce -> merge_into(persistent)."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        packed = torch.stack([self.grad_accum, self.grad_accum_abs, self.denom])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        self.grad_accum.copy_(packed[0])
        self.grad_accum_abs.copy_(packed[1])
        self.denom.copy_(packed[2])
        dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=group)

    def zero_(self):
        for t in (self.grad_accum, self.grad_accum_abs, self.denom, self.max_radii2D):
            t.zero_()
        return self

    def merge_into(self, total: "DensifyStats"):
        """Fold this (all-reduced) per-step delta into the persistent statistics: sums add, radii take the maximum."""
        total.grad_accum += self.grad_accum
        total.grad_accum_abs += self.grad_accum_abs
        total.denom += self.denom
        torch.maximum(total.max_radii2D, self.max_radii2D, out=total.max_radii2D)
        return total


BWD_TO_GROUP = {"means3D": "means3D", "sh": "sh", "opacity": "opacity", "scales": "scales", "rotations": "rotations",
                "language_feature": "language_feature", "instance_feature": "instance_feature"}


def accumulate_views(render_view, view_ids: Sequence[int], arena: GradArena, stats: DensifyStats = None
                     ) -> Tuple[GradArena, int]:
    """Run `render_view(v) -> (fwd dict, bwd dict)` for the local views and sum their gradients into the arena.

    `bwd` uses the names of the native backward tuple (means2D, means2D_abs, colors, language_feature,
    instance_feature, opacity, means3D, cov3D, sh, scales, rotations, all_map)."""
    n = 0
    for v in view_ids:
        fwd, bwd = render_view(v)
        arena.accumulate({g: bwd[k] for k, g in BWD_TO_GROUP.items() if k in bwd})
        if stats is not None:
            stats.add_view(bwd["means2D"], bwd["means2D_abs"], fwd["radii"], fwd.get("out_observe"))
        n += 1
    return arena, n


def multiview_step(render_view, n_views: int, arena: GradArena, stats: DensifyStats = None, group=None,
                   _delta: DensifyStats = None):
    """One optimisation step's gradient: local views -> arena -> one all-reduce.  Returns #local views.

    `stats` are the PERSISTENT densification statistics (cumulative over the steps of a densification interval): the
    step's views go into a zeroed per-step delta, the delta is all-reduced, and only then folded into `stats` — so after
    every step each rank holds exactly what single-process accumulation over all views of all steps so far would hold."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    arena.zero_()
    delta = None
    if stats is not None:
        delta = _delta.zero_() if _delta is not None else DensifyStats.allocate(stats.denom.numel(), stats.denom.device)
    _, n_local = accumulate_views(render_view, shard_views(n_views, world, rank), arena, delta)
    arena.all_reduce(group)
    if stats is not None:
        delta.all_reduce(group)
        delta.merge_into(stats)
    return n_local
