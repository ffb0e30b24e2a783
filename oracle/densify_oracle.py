"""CPU restatement (torch) of the reference's densification step.  TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows field_construction/scene/gaussian_model.py:
    add_densification_stats :720-724 (+ the max_radii2D update of field_construction/gaussian_field.py:521-523),
    densify_and_prune :700-718, densify_and_clone :664-698, densify_and_split :612-662,
    densification_postfix / cat_tensors_to_optimizer :561-610, prune_points / _prune_optimizer :520-559,
    reset_opacity :443-446, and build_rotation of field_construction/utils/general_utils.py:80-101.
It is written in the "plan" form the CUDA kernels use — every decision is taken on the ORIGINAL P rows and the result is one
gather — rather than the reference's grow / grow / prune sequence, so that agreement with the recorded vectors
(tests/golden/densify.npz, made by oracle/make_golden_densify.py from the reference's own methods) checks that reformulation:

  * rows appended by the clone step carry a zero padded gradient and the parent's scale, so the split step can never select
    them (thresholds > 0); the split step's quantiles see them only as `n_clone` extra zeros in front of the sorted values;
  * densification_postfix zeroes max_radii2D before the final prune, so `big_points_vs` (:713) is always false and the
    screen-size argument only switches the world-size test (:714) on — reproduced, not fixed;
  * a clone / split child is pruned by the same opacity (and, for children, the 1/1.6-scaled size) test as any other row;
  * output order: surviving originals, surviving clones, surviving first children, surviving second children;
    noise rows are indexed by the PRE-prune rank of the parent (the reference samples before it prunes).
PINNING: pinned by the recorded vectors above.  torch.quantile's linear interpolation is restated in quantile_padded().
"""
import math

import torch


def stats_update(stats, g2d, g2d_abs, radii, observe=None):
    """stats = dict(grad_accum, grad_accum_abs, denom, max_radii2D), each (P,) float32; updated in place."""
    vis = radii > 0
    stats["grad_accum"] += torch.where(vis, g2d[:, :2].norm(dim=-1), torch.zeros(()))
    stats["grad_accum_abs"] += torch.where(vis, g2d_abs[:, :2].norm(dim=-1), torch.zeros(()))
    stats["denom"] += vis.float()
    m = vis if observe is None else (vis & (observe > 0))
    stats["max_radii2D"] = torch.where(m, torch.maximum(stats["max_radii2D"], radii.float()), stats["max_radii2D"])
    return stats


def quantile_padded(values, n_pad, q):
    """torch.quantile(cat(zeros(n_pad), values), q) for values >= 0 (ATen quantile_compute: float32 rank, lerp)."""
    n = values.numel() + n_pad
    s = torch.sort(values.float())[0]
    rank = torch.tensor(q, dtype=torch.float32) * (n - 1)
    lo, hi = int(torch.floor(rank)), int(torch.ceil(rank))
    w = float(rank - torch.floor(rank))
    at = lambda i: torch.zeros(()) if i < n_pad else s[i - n_pad]
    return torch.lerp(at(lo), at(hi), torch.tensor(w, dtype=torch.float32))


def build_rotation(r):
    q = r / r.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)


def select(stats, scaling_raw, cfg, max_grad, abs_max_grad, extent):
    """-> (clone mask, split mask) over the original rows; cfg = dict(percent_dense, max_all_points,
    abs_split_radii2D_threshold, max_abs_split_points)."""
    P = scaling_raw.shape[0]
    g = stats["grad_accum"] / stats["denom"]
    ga = stats["grad_accum_abs"] / stats["denom"]
    g[g.isnan()] = 0.0
    ga[ga.isnan()] = 0.0
    smax = torch.exp(scaling_raw).max(dim=1).values
    big = smax > cfg["percent_dense"] * extent
    clone = (g >= max_grad) & ~big
    if int(clone.sum()) + P > cfg["max_all_points"]:                                     # :671-677
        tmp = torch.where(clone, g, torch.zeros(()))
        ratio = min((cfg["max_all_points"] - P) / float(P), 1)
        clone = tmp > quantile_padded(tmp, 0, 1.0 - ratio)
    n_clone = int(clone.sum())
    n_init = P + n_clone
    split = (g >= max_grad) & big
    if int(split.sum()) + n_init > cfg["max_all_points"]:                                # :625-630
        tmp = torch.where(split, g, torch.zeros(()))
        ratio = (cfg["max_all_points"] - n_init) / float(n_init)
        split = tmp > quantile_padded(tmp, n_clone, 1.0 - ratio)
    else:                                                                                # :632-642
        elig = big & (stats["max_radii2D"] > cfg["abs_split_radii2D_threshold"]) & ~split
        tmp = torch.where(elig, ga, torch.zeros(()))
        sel_abs = tmp >= abs_max_grad
        limited = min(cfg["max_all_points"] - n_init - int(split.sum()), cfg["max_abs_split_points"])
        if int(sel_abs.sum()) > limited:
            ratio = limited / float(n_init)
            sel_abs = tmp > quantile_padded(tmp, n_clone, 1.0 - ratio)
        split = split | sel_abs
    return clone, split


def densify_and_prune(params, m, v, stats, cfg, max_grad, abs_max_grad, min_opacity, extent, max_screen_size, z_clone, z_split,
                      roles=None):
    """params / m / v: dict name -> (P, w) tensors; roles: dict(xyz=, scaling=, rotation=, opacity=) -> group names.
    z_clone (n_clone, 3), z_split (2 n_split, 3): unit normal noise.  Returns (params, m, v, stats) of the new set."""
    roles = roles or {"xyz": "xyz", "scaling": "scaling", "rotation": "rotation", "opacity": "opacity"}
    xyz, sraw, rot, op = (params[roles[k]] for k in ("xyz", "scaling", "rotation", "opacity"))
    P = xyz.shape[0]
    clone, split = select(stats, sraw, cfg, max_grad, abs_max_grad, extent)
    scale = torch.exp(sraw)
    R = build_rotation(rot)
    low_op = torch.sigmoid(op[:, 0]) < min_opacity
    ws = bool(max_screen_size)
    prune_self = low_op | (ws & (scale.max(dim=1).values > 0.1 * extent))
    child_sraw = torch.log(scale / (0.8 * 2))
    prune_child = low_op | (ws & (torch.exp(child_sraw).max(dim=1).values > 0.1 * extent))
    keep_o = ~split & ~prune_self
    keep_c = clone & ~prune_self
    keep_s = split & ~prune_child
    rank_c = torch.cumsum(clone.long(), 0) - 1
    rank_s = torch.cumsum(split.long(), 0) - 1
    n_split = int(split.sum())
    io, ic, is_ = (torch.nonzero(k)[:, 0] for k in (keep_o, keep_c, keep_s))
    src = torch.cat([io, ic, is_, is_])
    kind = torch.cat([torch.zeros_like(io), torch.ones_like(ic), 2 * torch.ones_like(is_), 3 * torch.ones_like(is_)])
    zc = z_clone[rank_c[ic]] if ic.numel() else torch.zeros(0, 3)
    zs = torch.cat([z_split[rank_s[is_]], z_split[rank_s[is_] + n_split]]) if is_.numel() else torch.zeros(0, 3)
    new_rows = torch.cat([ic, is_, is_])
    offs = torch.bmm(R[new_rows], (scale[new_rows] * torch.cat([zc, zs])).unsqueeze(-1)).squeeze(-1)
    out_p, out_m, out_v = {}, {}, {}
    is_new = (kind > 0)[:, None]
    for name, t in params.items():
        r = t[src].clone()
        if name == roles["xyz"]:
            r[io.numel():] = offs + t[new_rows]
        if name == roles["scaling"]:
            r[kind >= 2] = child_sraw[src[kind >= 2]]
        out_p[name] = r
        out_m[name] = torch.where(is_new, torch.zeros(()), m[name][src])
        out_v[name] = torch.where(is_new, torch.zeros(()), v[name][src])
    n = src.numel()
    new_stats = {k: torch.zeros(n) for k in ("grad_accum", "grad_accum_abs", "denom", "max_radii2D")}
    return out_p, out_m, out_v, new_stats


def reset_opacity(opacity_raw):
    x = torch.minimum(torch.sigmoid(opacity_raw), torch.ones_like(opacity_raw) * 0.01)
    return torch.log(x / (1 - x))
