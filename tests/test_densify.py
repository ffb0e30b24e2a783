"""Next row: densification / pruning on the flat arenas.
CPU tier: the torch restatement (oracle/densify_oracle.py, "plan" formulation) against vectors recorded from the REFERENCE's own
GaussianModel.add_densification_stats / densify_and_prune / reset_opacity with a real torch.optim.Adam
(oracle/make_golden_densify.py).  GPU tier: the CUDA kernels (C ABI, lsx_b200.densify) against the same vectors — row sets,
row order, copied parameters and Adam moments BIT-EXACT; sampled positions, child scales and the reset opacities to 1e-6
(expf / logf vs torch's CPU exp / log) — and, at 1 M Gaussians, against the restatement plus conservation properties."""
import os

import numpy as np
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)
from oracle import densify_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify.npz")
CASES = ("plain", "abs", "abs_capped", "clone_capped", "split_capped", "nothing")
GROUPS = ("xyz", "knn_f", "f_dc", "f_rest", "opacity", "scaling", "rotation", "language_feature", "instance_feature")
STATS = ("grad_accum", "grad_accum_abs", "denom", "max_radii2D")


def _case(name):
    z = np.load(GOLD)
    c = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "_")}
    e, pd, mg, ag, mo, size, max_all, abs_r, max_abs = [float(x) for x in c["cfg"]]
    c["args"] = dict(cfg=dict(percent_dense=pd, max_all_points=int(max_all), abs_split_radii2D_threshold=abs_r,
                              max_abs_split_points=int(max_abs)),
                     max_grad=mg, abs_max_grad=ag, min_opacity=mo, extent=e, max_screen_size=None if size < 0 else size)
    return c


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30)) if b.numel() else 0.0


@pytest.mark.parametrize("name", CASES)
def test_restatement_stats_match_reference_vectors(name):
    c = _case(name)
    P = c["in_xyz"].shape[0]
    st = {k: torch.zeros(P) for k in STATS}
    for v in range(3):
        orc.stats_update(st, c[f"view{v}_g2d"], c[f"view{v}_g2d_abs"], c[f"view{v}_radii"], c[f"view{v}_observe"])
    for k in STATS:
        assert torch.equal(st[k], c["in_" + k]), k


@pytest.mark.parametrize("name", CASES)
def test_restatement_densify_matches_reference_vectors(name):
    c = _case(name)
    params = {n: c["in_" + n] for n in GROUPS}
    m = {n: c["in_m_" + n] for n in GROUPS}
    v = {n: c["in_v_" + n] for n in GROUPS}
    st = {k: c["in_" + k].clone() for k in STATS}
    op, om, ov, ost = orc.densify_and_prune(params, m, v, st, z_clone=c["z_clone"], z_split=c["z_split"], **c["args"])
    for n in GROUPS:
        assert op[n].shape == c["out_" + n].shape, (n, op[n].shape, c["out_" + n].shape)
        if n in ("xyz", "scaling"):
            assert _rel(op[n], c["out_" + n]) < 1e-6, n
        else:
            assert torch.equal(op[n], c["out_" + n]), n
        assert torch.equal(om[n], c["out_m_" + n]) and torch.equal(ov[n], c["out_v_" + n]), n
    for k in STATS:
        assert torch.equal(ost[k], c["out_" + k]), k
    assert _rel(orc.reset_opacity(op["opacity"]), c["reset_opacity"]) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------------
# GPU tier
# ---------------------------------------------------------------------------------------------------------------------------
def _arena_from(c, prefix, dev):
    """Flat arenas (parameters, exp_avg, exp_avg_sq) with the reference's nine groups."""
    from lsx_b200.densify import ParamArena
    widths = {n: c[f"{prefix}_{n}"].shape[1] for n in GROUPS}
    P = c[f"{prefix}_xyz"].shape[0]
    arenas = [ParamArena.allocate(P, widths, dev) for _ in range(3)]
    for a, pre in zip(arenas, (f"{prefix}_", f"{prefix}_m_", f"{prefix}_v_")):
        for n in GROUPS:
            a.views[n].copy_(c[pre + n])
    return arenas


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_stats_match_reference_vectors(name):
    from lsx_b200.densify import add_densification_stats
    from lsx_b200.multiview import DensifyStats
    c = _case(name)
    dev = torch.device("cuda:0")
    st = DensifyStats.allocate(c["in_xyz"].shape[0], dev)
    for v in range(3):
        add_densification_stats(st, c[f"view{v}_g2d"].to(dev), c[f"view{v}_g2d_abs"].to(dev), c[f"view{v}_radii"].to(dev),
                                c[f"view{v}_observe"].to(dev))
    for k in STATS:
        # sqrt(x*x + y*y): correctly rounded sqrtf of an fma-free sum on both sides -> 1 ulp at most
        assert _rel(getattr(st, k).cpu(), c["in_" + k]) < 2e-7, k
    assert torch.equal(st.denom.cpu(), c["in_denom"]) and torch.equal(st.max_radii2D.cpu(), c["in_max_radii2D"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_densify_matches_reference_vectors(name):
    from lsx_b200.densify import DensifyConfig, densify_and_prune, reset_opacity
    from lsx_b200.multiview import DensifyStats
    c = _case(name)
    dev = torch.device("cuda:0")
    params, m, v = _arena_from(c, "in", dev)
    st = DensifyStats(*(c["in_" + k].to(dev).clone() for k in STATS))
    a = c["args"]
    noise = {"clone": c["z_clone"].to(dev), "split": c["z_split"].to(dev)}
    res = densify_and_prune(params, m, v, st, DensifyConfig(**a["cfg"]), a["max_grad"], a["abs_max_grad"], a["min_opacity"],
                            a["extent"], a["max_screen_size"], noise_fn=lambda kind, n: noise[kind][:n])
    assert res.params.P == c["out_xyz"].shape[0]
    assert res.n_clone == c["z_clone"].shape[0] and 2 * res.n_split == c["z_split"].shape[0]
    for n in GROUPS:
        got, want = res.params.views[n].cpu(), c["out_" + n]
        if n in ("xyz", "scaling"):
            assert _rel(got, want) < 1e-6, n
        else:
            assert torch.equal(got, want), n
        assert torch.equal(res.exp_avg.views[n].cpu(), c["out_m_" + n]), n
        assert torch.equal(res.exp_avg_sq.views[n].cpu(), c["out_v_" + n]), n
    for k in STATS:
        assert torch.equal(getattr(res.stats, k).cpu(), c["out_" + k]), k
    for arena in (res.params, res.exp_avg, res.exp_avg_sq):      # alignment padding between groups is zero-filled
        used = torch.zeros_like(arena.flat, dtype=torch.bool)
        for o, n in arena.offsets.values():
            used[o:o + n] = True
        assert float(arena.flat[~used].abs().sum()) == 0.0
    reset_opacity(res.params, res.exp_avg, res.exp_avg_sq)
    assert _rel(res.params.views["opacity"].cpu(), c["reset_opacity"]) < 1e-6
    assert float(res.exp_avg.views["opacity"].abs().max()) == 0.0 and float(res.exp_avg_sq.views["opacity"].abs().max()) == 0.0
    assert torch.equal(res.exp_avg.views["xyz"].cpu(), c["out_m_xyz"])     # other groups untouched


@pytest.mark.gpu
def test_cuda_densify_1M_properties_and_restatement():
    """1 M Gaussians with the reference's own configuration values: agreement with the restatement (same noise) and
    size-independent properties: survivors keep their order and bits, new rows have zero moments, counts add up."""
    from lsx_b200.densify import DensifyConfig, ParamArena, densify_and_prune
    from lsx_b200.multiview import DensifyStats
    dev = torch.device("cuda:0")
    P, F = 1_000_000, 16
    g = torch.Generator().manual_seed(11)
    widths = {"xyz": 3, "f_dc": 3, "f_rest": 45, "opacity": 1, "scaling": 3, "rotation": 4, "language_feature": F,
              "instance_feature": 3}
    extent, pd = 5.0, 0.001
    host = {n: torch.randn(P, w, generator=g) for n, w in widths.items()}
    host["scaling"] = host["scaling"] * 0.8 + float(np.log(pd * extent))
    host["opacity"] *= 2.0
    hm = {n: torch.randn(P, w, generator=g) * 1e-3 for n, w in widths.items()}
    hv = {n: torch.rand(P, w, generator=g) * 1e-6 for n, w in widths.items()}
    arenas = [ParamArena.allocate(P, widths, dev) for _ in range(3)]
    for a, src in zip(arenas, (host, hm, hv)):
        for n in widths:
            a.views[n].copy_(src[n])
    denom = torch.randint(0, 4, (P,), generator=g).float()
    hs = {"grad_accum": denom * torch.rand(P, generator=g) * 0.008, "grad_accum_abs": denom * torch.rand(P, generator=g) * 0.03,
          "denom": denom, "max_radii2D": torch.randint(0, 45, (P,), generator=g).float()}
    st = DensifyStats(*(hs[k].to(dev) for k in STATS))
    cfg = dict(percent_dense=pd, max_all_points=12_000_000, abs_split_radii2D_threshold=20, max_abs_split_points=50_000)
    zc, zs = torch.randn(P, 3, generator=g), torch.randn(2 * P, 3, generator=g)
    calls = {}

    def noise_fn(kind, n):
        calls[kind] = n
        return (zc if kind == "clone" else zs)[:n].to(dev)

    res = densify_and_prune(arenas[0], arenas[1], arenas[2], st, DensifyConfig(**cfg), 0.004, 0.016, 0.05, extent, 20,
                            noise_fn=noise_fn)
    op, om, ov, _ = orc.densify_and_prune(host, hm, hv, {k: x.clone() for k, x in hs.items()}, cfg, 0.004, 0.016, 0.05, extent, 20,
                                          zc[:calls["clone"]], zs[:calls["split"]])
    assert res.params.P == op["xyz"].shape[0] and res.n_clone > 1000 and res.n_split > 1000
    for n in widths:
        got = res.params.views[n].cpu()
        if n in ("xyz", "scaling"):
            assert _rel(got, op[n]) < 1e-6, n
        else:
            assert torch.equal(got, op[n]), n
        assert torch.equal(res.exp_avg.views[n].cpu(), om[n]) and torch.equal(res.exp_avg_sq.views[n].cpu(), ov[n]), n
    n_new = res.n_kept_clone + 2 * res.n_kept_split
    assert res.params.P == res.n_kept_original + n_new
    assert float(res.exp_avg.views["f_rest"][res.n_kept_original:].abs().max()) == 0.0
    assert float(res.stats.denom.abs().max()) == 0.0 and res.stats.denom.numel() == res.params.P
    with pytest.raises(RuntimeError):
        densify_and_prune(ParamArena.allocate(4, widths, "cpu"), None, None, None, DensifyConfig(**cfg), 0.004, 0.016, 0.05,
                          extent, 20)   # CPU arenas: no CPU path
