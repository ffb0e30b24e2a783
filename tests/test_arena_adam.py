"""Arena Adam against torch.optim.Adam — the optimiser the reference constructs unmodified
(gaussian_model.py:313-328: per-group lr, eps=1e-15) — on the same parameters and gradients, several steps."""
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)
from lsx_b200.multiview import PARAM_GROUPS, GradArena

LRS = {"means3D": 1.6e-4, "sh": 2.5e-3, "opacity": 5e-2, "scales": 5e-3, "rotations": 1e-3, "language_feature": 2.5e-3,
       "instance_feature": 2.5e-3, "all_map": 0.0}


@pytest.mark.gpu
def test_matches_torch_adam_over_steps():
    from lsx_b200.optim import ArenaAdam
    P, M, F, Fi = 1003, 16, 16, 3
    dev = "cuda:0"
    g = torch.Generator().manual_seed(3)
    params = GradArena.allocate(P, M, F, Fi, dev)
    params.flat.copy_(torch.randn(params.flat.numel(), generator=g))
    # the reference formulation: one nn.Parameter per group, torch.optim.Adam with per-group lr
    ref_p = {n: torch.nn.Parameter(params.views[n].detach().clone()) for n in PARAM_GROUPS if params.views[n].numel()}
    ref = torch.optim.Adam([{"params": [ref_p[n]], "lr": LRS[n], "name": n} for n in ref_p], lr=0.0, eps=1e-15)
    opt = ArenaAdam(params, LRS)
    grads = GradArena.allocate(P, M, F, Fi, dev)
    for step in range(5):
        grads.flat.copy_(torch.randn(grads.flat.numel(), generator=g) * (10.0 ** (step - 3)))
        for n in ref_p:
            ref_p[n].grad = grads.views[n].detach().clone()
        ref.step()
        opt.step(grads)
        for n in ref_p:
            a, b = params.views[n], ref_p[n].detach()
            assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max())), (step, n)
    # padding between groups is never written with NaNs and all_map (lr 0) did not move
    assert bool(torch.isfinite(params.flat).all())


@pytest.mark.gpu
def test_rejects_cpu_and_mismatched_arenas():
    from lsx_b200.optim import ArenaAdam
    with pytest.raises(RuntimeError):
        ArenaAdam(GradArena.allocate(10, 16, 3, 3, "cpu"), LRS)
    opt = ArenaAdam(GradArena.allocate(64, 16, 3, 3, "cuda:0"), LRS)
    with pytest.raises(RuntimeError):
        opt.step(GradArena.allocate(65, 16, 3, 3, "cuda:0"))
