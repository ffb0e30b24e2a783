"""Diagnostic: which allocation pattern makes torch's caching allocator grow without bound?"""
import gc
import os
import sys

mode = sys.argv[1] if len(sys.argv) > 1 else "base"
if mode == "expandable":
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = "expandable_segments:True"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
c, scene, grads, bg, views = bench.build_views("C3", dev, 0, 1, 1)
fargs = views[0]["fargs"]
mod = hz.ref_rast_for(16) if mode == "ref" else ops
gc.collect()
gc.disable()
n = 40
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
rows = []
ev[0].record()
fwd = bwd = None
for i in range(n):
    if mode == "del":
        fwd = bwd = None
    fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*fargs)))
    bwd = dict(zip(hz.BWD_NAMES, mod.rasterize_gaussians_backward(*hz.native_backward_args(fargs, fwd, grads))))
    ev[i + 1].record()
    st = torch.cuda.memory_stats()
    rows.append((st.get("num_device_alloc", 0), st.get("reserved_bytes.all.current", 0) / 1e9, st.get("allocated_bytes.all.current", 0) / 1e9))
torch.cuda.synchronize()
print(mode, "ms:", " ".join(f"{ev[i].elapsed_time(ev[i + 1]):.1f}" for i in range(n)))
print(mode, "cudaMalloc:", [r[0] for r in rows][::4], "reserved GB:", [round(r[1], 1) for r in rows][::4], "allocated GB:", [round(r[2], 2) for r in rows][::8])
