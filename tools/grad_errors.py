"""Per-tensor error report of the new operator against the recompiled reference (oracle/_ref) on the edge cases
and one mid-size scene.  Diagnostic for the gradient tolerance; run on a GPU box."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402
from lsx_b200.synthetic import make_camera, make_scene, make_upstream_grads  # noqa: E402

dev = torch.device("cuda:0")


def report(name, scene, cam, grads, F, bg):
    ref = hz.ref_rast_for(F)
    fargs = hz.native_forward_args(scene, cam, bg, F)
    rf, rb = hz.run_native(ref, fargs, grads)
    rb2 = dict(zip(hz.BWD_NAMES, ref.rasterize_gaussians_backward(*hz.native_backward_args(fargs, rf, grads))))
    nf, nb = hz.run_native(ops, fargs, grads)
    torch.cuda.synchronize()
    print(f"== {name}: R={rf['num_rendered']}")
    for k in ["color", "language_feature", "instance_feature", "all_map", "plane_depth"]:
        if rf[k].numel() > 1 and float(rf[k].abs().max()) > 0:
            print(f"   fwd {k:18s} {hz.rel_err(nf[k], rf[k]):.3e}")
    for k in hz.BWD_NAMES:
        if rb[k].numel() > 1 and float(rb[k].abs().max()) > 0:
            print(f"   bwd {k:18s} err {hz.rel_err(nb[k], rb[k]):.3e}  ref self-spread {hz.rel_err(rb2[k], rb[k]):.3e}  max|ref| {float(rb[k].abs().max()):.3e}")


def mk(P, W, H, F, seed, s_med=None):
    scene = make_scene(P, W, H, F=F, seed=seed, s_med=s_med).to(dev)
    cam = make_camera(W, H).to(dev)
    grads = make_upstream_grads(W, H, F, seed=seed + 1, device=dev)
    return scene, cam, grads


W, H, F = 64, 48, 3
bg = torch.tensor([0.5, 0.25, 0.0], device=dev)
scene, cam, grads = mk(2, W, H, F, 4, 0.05)
scene.means3D[0] = torch.tensor([0.0, 0.0, 1.0], device=dev)
scene.scales[0] = torch.tensor([2.0, 2.0, 2.0], device=dev)
scene.means3D[1] = torch.tensor([50.0, 0.0, 3.0], device=dev)
report("huge splat", scene, cam, grads, F, bg)
scene, cam, grads = mk(64, W, H, F, 5, 0.05)
scene.means3D[:] = scene.means3D[0]
report("64 identical", scene, cam, grads, F, bg)
scene, cam, grads = mk(30_000, 320, 240, 16, 2)
report("30k F16", scene, cam, grads, 16, torch.tensor([0.1, 0.3, 0.7], device=dev))
scene, cam, grads = mk(100_000, 800, 800, 3, 3)
report("C2", scene, cam, grads, 3, torch.tensor([0.1, 0.3, 0.7], device=dev))
