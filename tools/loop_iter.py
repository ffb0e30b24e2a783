"""A few FieldLoop steps at a BASELINE loop config for a launch list:  python tools/loop_iter.py C4 [views] [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench_loop as bl  # noqa: E402
from lsx_b200.field_loop import FieldLoop, LoopConfig  # noqa: E402
from lsx_b200.synthetic import CONFIGS, make_scene  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
c = CONFIGS[name]
dev = torch.device("cuda:0")
scene = make_scene(c["P"], c["W"], c["H"], F=c["F"], seed=0, s_med=c["s_med"]).to(dev)
cfg = LoopConfig(cls3d=(name == "C4"), cls3d_tree=("tree" in sys.argv[4:]), fused_wrapper=("unfused" not in sys.argv[4:]))
loop = FieldLoop(bl.make_raw(scene), bl.LRS, torch.zeros(3, device=dev), cfg, n_views=c["views"], poses=bl.make_poses(c["views"], dev))
views = [bl.make_view(v, c["views"], c["W"], c["H"], c["F"], dev) for v in range(V)]
si = [bl.sample_indices(v, 0, c["P"], 800, dev) for v in range(V)] if cfg.cls3d else None
for _ in range(steps):
    loop.step(views, si)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loop.step(views, si)
e1.record()
torch.cuda.synchronize()
print("loop_iter done", name, V, steps, "ms_per_view", round(e0.elapsed_time(e1) / (steps * V), 4), "lib", os.environ.get("LSX_B200_LIB", "default"))
