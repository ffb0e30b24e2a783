"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): the rasterizer forward + backward at C1's shape and on
ragged / empty / degenerate inputs, distCUDA2, and one FieldLoop step.
    compute-sanitizer --tool memcheck  --log-file gpurun_out/memcheck.log  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck.log python tools/sanitize_case.py render"""
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
only_render = len(sys.argv) > 1 and sys.argv[1] == "render"
cases = [(10_000, 256, 256, 3), (3_000, 113, 67, 16), (1, 64, 64, 3), (500, 16, 16, 29)]   # C1, ragged tiles, one splat, widest record
if only_render:
    cases = cases[1:2]
for P, W, H, F in cases:
    scene = make_scene(P, W, H, F=F, seed=P).to(dev)
    cam = make_camera(W, H, yaw_deg=2.0).to(dev)
    grads = make_upstream_grads(W, H, F, device=dev)
    fargs = hz.native_forward_args(scene, cam, torch.zeros(3, device=dev), F)
    fwd, bwd = hz.run_native(ops, fargs, grads)
    st = ops.render_stats(fwd["num_rendered"], fwd["geom"], fwd["binning"], fwd["img"], P, H, W, 3 + F + 3 + 5)
    torch.cuda.synchronize()
    print("case", P, W, H, F, "R =", fwd["num_rendered"], "B =", st["B"], "grad checksum", float(bwd["means3D"].abs().sum()), flush=True)
if not only_render:
    # everything culled (behind the camera) and the empty input
    scene = make_scene(2_000, 96, 64, F=3, seed=1).to(dev)
    scene.means3D[:, 2] = -5.0
    cam = make_camera(96, 64).to(dev)
    fwd, bwd = hz.run_native(ops, hz.native_forward_args(scene, cam, torch.zeros(3, device=dev), 3), make_upstream_grads(96, 64, 3, device=dev))
    print("all culled: R =", fwd["num_rendered"], flush=True)
    print("knn", float(ops.distCUDA2(make_scene(2_000, 64, 64, seed=2).means3D.to(dev)).sum()), flush=True)
    import bench_loop as bl
    from lsx_b200.field_loop import FieldLoop, LoopConfig
    P, W, H, F = 4_000, 96, 64, 3
    scene = make_scene(P, W, H, F=F, seed=9).to(dev)
    views = [bl.make_view(v, 2, W, H, F, dev) for v in range(2)]
    loop = FieldLoop(bl.make_raw(scene), bl.LRS, torch.zeros(3, device=dev), LoopConfig(), n_views=2, poses=bl.make_poses(2, dev))
    out = loop.step(views, [bl.sample_indices(v, 0, P, 800, dev) for v in range(2)])
    torch.cuda.synchronize()
    print("field loop step", {k: round(float(v), 5) for k, v in out.items()}, flush=True)
print("sanitize_case done", flush=True)
