"""Diagnostic: per-step device times and allocator activity of bench.py's batch step, to locate sporadic long steps."""
import gc
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from lsx_b200 import ops  # noqa: E402
from lsx_b200.multiview import GradArena  # noqa: E402

dev = torch.device("cuda:0")
V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c, scene, grads, bg, views = bench.build_views("C3", dev, 0, 1, V)
arena = GradArena.allocate(c["P"], 16, c["F"], 3, dev)
step = bench.batch_stepper(ops, views, grads, arena, 1)
gc.collect()
gc.disable()
for rep in range(2):
    n = 16
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    allocs, reserved = [], []
    ev[0].record()
    for i in range(n):
        step()
        ev[i + 1].record()
        st = torch.cuda.memory_stats()
        allocs.append(st.get("num_device_alloc", 0))
        reserved.append(st.get("reserved_bytes.all.current", 0) / 1e9)
    torch.cuda.synchronize()
    print("rep", rep, "ms/view:", " ".join(f"{ev[i].elapsed_time(ev[i + 1]) / V:.2f}" for i in range(n)))
    print("   cudaMalloc calls (cumulative):", allocs)
    print("   reserved GB:", " ".join(f"{r:.1f}" for r in reserved))
