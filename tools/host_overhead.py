"""Host-side cost of one forward + backward call (tiny scene: GPU time negligible), new binding vs reference pybind."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402
from lsx_b200.synthetic import make_all_map, make_camera, make_scene, make_upstream_grads  # noqa: E402

dev = torch.device("cuda:0")
P, W, H, F = 2000, 64, 64, 3
scene = make_scene(P, W, H, F=F, seed=0).to(dev)
cam = make_camera(W, H).to(dev)
grads = make_upstream_grads(W, H, F, device=dev)
fargs = hz.native_forward_args(scene, cam, torch.zeros(3, device=dev), F, all_map=make_all_map(scene, cam))
for name, mod in (("new (ctypes -> C ABI)", ops), ("reference (pybind)", hz.ref_rast_for(F))):
    if mod is None:
        continue
    for _ in range(20):
        hz.run_native(mod, fargs, grads)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*fargs)))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(n):
        mod.rasterize_gaussians_backward(*hz.native_backward_args(fargs, fwd, grads))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: forward {1e6 * (t1 - t0) / n:.0f} us/call, backward {1e6 * (t2 - t1) / n:.0f} us/call (wall, tiny scene)")
