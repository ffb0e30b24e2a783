"""Experiment: does running view i+1's forward on a second stream while view i's backward runs shorten a multi-view step?"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402
from lsx_b200.multiview import GradArena  # noqa: E402

dev = torch.device("cuda:0")
V = 8
CFG = sys.argv[1] if len(sys.argv) > 1 else "C3"
c, scene, grads, bg, views = bench.build_views(CFG, dev, 0, 1, V)
arena = GradArena.allocate(c["P"], 16, c["F"], 3, dev)
seq = bench.batch_stepper(ops, views, grads, arena, 1)


def make_pipelined(prio):
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    streams = [torch.cuda.Stream(device=dev, priority=(-1 if prio else 0)) for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(V)]

    def fwd_on(i):
        with torch.cuda.stream(streams[i & 1]):
            return dict(zip(hz.FWD_NAMES, ops.rasterize_gaussians(*views[i]["fargs"])))

    def step():
        main = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(main)
        f = fwd_on(0)
        for i in range(V):
            s = streams[i & 1]
            with torch.cuda.stream(s):
                if i > 0:
                    s.wait_event(done[i - 1])
                bargs = hz.native_backward_args(views[i]["fargs"], f, grads)
                ops.rasterize_gaussians_backward(*bargs, grad_buffers=arena.grad_buffers(), accumulate=i > 0)
                done[i].record(s)
            if i + 1 < V:
                f = fwd_on(i + 1)
        main.wait_event(done[V - 1])
    return step


def timeit(fn, n=6):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / V


print("sequential  ms/view", timeit(seq))
ref = arena.flat.clone()
print("pipelined   ms/view", timeit(make_pipelined(False)))
print("max rel diff of the arena vs sequential:", float((arena.flat - ref).abs().max() / ref.abs().max()))
print("pipelined (high-priority streams) ms/view", timeit(make_pipelined(True)))
