"""Diagnostic: per-step wall/device times of the raw native step loop, to locate sporadic long steps."""
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
import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
c, scene, cam, grads, bg, am, fargs = bench.build_case("C3", dev)
step = bench.native_stepper(ops, fargs, grads)
sampler = None
for rep in range(6):
    if rep == 3:
        sampler = bench.ClockSampler(0)
        print('sampler started')
    gc.collect()
    if rep == 2:
        gc.disable()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    wall = []
    ev[0].record()
    for i in range(40):
        t0 = time.perf_counter()
        step()
        wall.append((time.perf_counter() - t0) * 1e3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    devms = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
    print("rep", rep, "device ms:", " ".join(f"{x:.1f}" for x in devms))
    print("rep", rep, "wall   ms:", " ".join(f"{x:.1f}" for x in wall))
    print("mem reserved GB", torch.cuda.memory_reserved() / 1e9, "num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"),
          "segments", torch.cuda.memory_stats().get("segment.all.allocated"))

if sampler: print(sampler.stop())
