"""Warm per-kernel device times of FieldLoop steps (torch.profiler / CUPTI), top kernels and the GPU-busy fraction:
    python tools/loop_profile.py C4 [views] [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench_loop as bl  # noqa: E402
from lsx_b200.field_loop import FieldLoop, LoopConfig  # noqa: E402
from lsx_b200.synthetic import CONFIGS, make_scene  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
c = CONFIGS[name]
dev = torch.device("cuda:0")
scene = make_scene(c["P"], c["W"], c["H"], F=c["F"], seed=0, s_med=c["s_med"]).to(dev)
cfg = LoopConfig(cls3d=(name == "C4"))
loop = FieldLoop(bl.make_raw(scene), bl.LRS, torch.zeros(3, device=dev), cfg, n_views=c["views"], poses=bl.make_poses(c["views"], dev))
views = [bl.make_view(v, c["views"], c["W"], c["H"], c["F"], dev) for v in range(V)]
si = [bl.sample_indices(v, 0, c["P"], 800, dev) for v in range(V)] if cfg.cls3d else None
for _ in range(5):
    loop.step(views, si)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loop.step(views, si)
e1.record()
torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / (steps * V)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        loop.step(views, si)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3 / (steps * V), e.count / (steps * V)) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print(f"{name}: {wall:.4f} ms per view (events, no profiler); sum of kernel times {total:.4f} ms per view -> GPU busy {100 * total / wall:.1f} %")
for k, ms, n in rows[:28]:
    print(f"  {ms:8.4f} ms  x{n:5.2f}  {k[:110]}")
