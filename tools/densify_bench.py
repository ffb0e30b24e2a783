"""Densification on the flat arenas (plan + one gather) vs the reference's torch-op formulation (cat / boolean-mask of every
parameter and Adam-moment tensor, twice growing and twice pruning), on CUDA, same selection, same noise.  Also times the
per-view statistics kernel against its torch-op form."""
import json
import math
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.densify import DensifyConfig, ParamArena, add_densification_stats, densify_and_prune  # noqa: E402
from lsx_b200.multiview import DensifyStats  # noqa: E402

dev = torch.device("cuda:0")
WIDTHS = {"xyz": 3, "f_dc": 3, "f_rest": 45, "opacity": 1, "scaling": 3, "rotation": 4, "language_feature": 16, "instance_feature": 3}
EXTENT, PD = 5.0, 0.001


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def build_rotation(r):
    q = r / r.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z),
                        1 - 2 * (x * x + z * z), 2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x),
                        1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)


def torch_ops_densify(p, m, v, st, max_grad, min_opacity, zc, zs):
    """The reference's sequence (gaussian_model.py:612-718, uncapped branches, abs split off as in the shipped config) with
    torch ops on CUDA tensors: grow by cat (params + both moments), grow again, prune by mask, prune by mask."""
    g = st.grad_accum / st.denom
    g[g.isnan()] = 0.0
    scale = torch.exp(p["scaling"])
    big = scale.max(dim=1).values > PD * EXTENT
    clone = (g >= max_grad) & ~big
    new = {n: t[clone] for n, t in p.items()}
    new["xyz"] = torch.bmm(build_rotation(p["rotation"][clone]), (scale[clone] * zc).unsqueeze(-1)).squeeze(-1) + p["xyz"][clone]
    p = {n: torch.cat((t, new[n])) for n, t in p.items()}
    m = {n: torch.cat((t, torch.zeros_like(new[n]))) for n, t in m.items()}
    v = {n: torch.cat((t, torch.zeros_like(new[n]))) for n, t in v.items()}
    n_init = p["xyz"].shape[0]
    pg = torch.zeros(n_init, device=dev)
    pg[:g.shape[0]] = g
    scale = torch.exp(p["scaling"])
    split = (pg >= max_grad) & (scale.max(dim=1).values > PD * EXTENT)
    new = {n: t[split].repeat(2, 1) for n, t in p.items()}
    new["xyz"] = torch.bmm(build_rotation(p["rotation"][split]).repeat(2, 1, 1),
                           (scale[split].repeat(2, 1) * zs).unsqueeze(-1)).squeeze(-1) + p["xyz"][split].repeat(2, 1)
    new["scaling"] = torch.log(scale[split].repeat(2, 1) / 1.6)
    p = {n: torch.cat((t, new[n])) for n, t in p.items()}
    m = {n: torch.cat((t, torch.zeros_like(new[n]))) for n, t in m.items()}
    v = {n: torch.cat((t, torch.zeros_like(new[n]))) for n, t in v.items()}
    keep = ~torch.cat((split, torch.zeros(2 * int(split.sum()), device=dev, dtype=torch.bool)))
    p, m, v = ({n: t[keep] for n, t in d.items()} for d in (p, m, v))
    prune = (torch.sigmoid(p["opacity"][:, 0]) < min_opacity) | (torch.exp(p["scaling"]).max(dim=1).values > 0.1 * EXTENT)
    keep = ~prune
    p, m, v = ({n: t[keep] for n, t in d.items()} for d in (p, m, v))
    return p, m, v


for P in (500_000, 1_000_000, 5_000_000):
    g = torch.Generator(device=dev).manual_seed(P)
    arenas = [ParamArena.allocate(P, WIDTHS, dev) for _ in range(3)]
    for a, s in zip(arenas, (1.0, 1e-3, 1e-6)):
        a.flat.normal_(generator=g).mul_(s)
    arenas[2].flat.abs_()
    arenas[0].views["scaling"].mul_(0.8).add_(math.log(PD * EXTENT))
    arenas[0].views["opacity"].mul_(2.0)
    denom = torch.randint(0, 4, (P,), generator=g, device=dev).float()
    st = DensifyStats(denom * torch.rand(P, generator=g, device=dev) * 0.008, denom * torch.rand(P, generator=g, device=dev) * 0.03,
                      denom, torch.randint(0, 45, (P,), generator=g, device=dev).float())
    zc, zs = torch.randn(P, 3, generator=g, device=dev), torch.randn(2 * P, 3, generator=g, device=dev)
    cfg = DensifyConfig(percent_dense=PD)
    noise = lambda kind, n: (zc if kind == "clone" else zs)[:n]
    res = densify_and_prune(arenas[0], arenas[1], arenas[2], st, cfg, 0.004, 0.016, 0.05, EXTENT, 20, noise_fn=noise)
    t_new = timeit(lambda: densify_and_prune(arenas[0], arenas[1], arenas[2], st, cfg, 0.004, 0.016, 0.05, EXTENT, 20, noise_fn=noise))
    dicts = [{n: a.views[n].clone() for n in WIDTHS} for a in arenas]
    rp, rm, rv = torch_ops_densify(*dicts, st, 0.004, 0.05, zc[:res.n_clone], zs[:2 * res.n_split])
    same_rows = rp["xyz"].shape[0] == res.params.P
    same = same_rows and all(torch.equal(rp[n], res.params.views[n]) for n in WIDTHS if n not in ("xyz", "scaling")) and \
        all(torch.equal(rm[n], res.exp_avg.views[n]) for n in WIDTHS)
    t_ref = timeit(lambda: torch_ops_densify(*dicts, st, 0.004, 0.05, zc[:res.n_clone], zs[:2 * res.n_split]), n=5, warm=2)
    floats = sum(WIDTHS.values())
    alg_bytes = P * 32 + 24 * floats * res.params.P        # classify inputs + read/write of parameter and both moments
    # per-view statistics
    g2, g2a = torch.randn(P, 3, device=dev), torch.randn(P, 3, device=dev).abs()
    radii, obs = torch.randint(0, 40, (P,), device=dev, dtype=torch.int32), torch.randint(0, 3, (P,), device=dev, dtype=torch.int32)
    st2 = DensifyStats.allocate(P, dev)
    t_stats = timeit(lambda: add_densification_stats(st2, g2, g2a, radii, obs), n=50)
    st3 = DensifyStats.allocate(P, dev)
    t_stats_ref = timeit(lambda: st3.add_view(g2, g2a, radii, obs), n=50)
    print(json.dumps({"P": P, "P_new": res.params.P, "n_clone": res.n_clone, "n_split": res.n_split,
                      "arena_densify_ms": round(t_new, 3), "torch_ops_densify_ms": round(t_ref, 3), "speedup": round(t_ref / t_new, 2),
                      "same_rows_and_bits_as_torch_ops": bool(same), "algorithmic_GBps": round(alg_bytes / t_new / 1e6, 1),
                      "stats_update_ms": round(t_stats, 4), "stats_update_torch_ops_ms": round(t_stats_ref, 4),
                      "stats_GBps": round(P * 56 / t_stats / 1e6, 1)}), flush=True)
