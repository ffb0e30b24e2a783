"""Adam step over the 1M-Gaussian parameter set: ArenaAdam (one kernel) vs torch.optim.Adam as the reference builds it
(default implementation) and with fused=True."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.multiview import PARAM_GROUPS, GradArena  # noqa: E402
from lsx_b200.optim import ArenaAdam  # noqa: E402

LRS = {"means3D": 1.6e-4, "sh": 2.5e-3, "opacity": 5e-2, "scales": 5e-3, "rotations": 1e-3, "language_feature": 2.5e-3,
       "instance_feature": 2.5e-3, "all_map": 0.0}


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for P in (1_000_000, 500_000):
    F = 16 if P == 1_000_000 else 3
    params = GradArena.allocate(P, 16, F, 3, "cuda:0")
    params.flat.normal_()
    grads = GradArena.allocate(P, 16, F, 3, "cuda:0")
    grads.flat.normal_()
    opt = ArenaAdam(params, LRS)
    row = {"P": P, "elements": params.flat.numel(), "arena_adam_ms": timeit(lambda: opt.step(grads))}
    row["arena_GBps"] = params.flat.numel() * 28 / row["arena_adam_ms"] / 1e6
    for fused in (False, True):
        ps = {n: torch.nn.Parameter(params.views[n].detach().clone()) for n in PARAM_GROUPS if params.views[n].numel()}
        for n in ps:
            ps[n].grad = grads.views[n].detach().clone()
        o = torch.optim.Adam([{"params": [ps[n]], "lr": LRS[n]} for n in ps], lr=0.0, eps=1e-15, fused=fused)
        row["torch_fused_ms" if fused else "torch_default_ms"] = timeit(o.step)
    print(json.dumps(row))
