"""Minimal driver for profiler captures: N forward+backward steps of one configuration, nothing else.
Usage (GPU box):  python tools/one_step.py [--config C3] [--iters 2] [--knn 0]
Under ncu:        ncu --set full -k regex:render_ -s 2 -c 2 ... python tools/one_step.py --iters 2
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402
from lsx_b200.synthetic import CONFIGS, make_all_map, make_camera, make_scene, make_upstream_grads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--knn", type=int, default=0, help="also run distCUDA2 on this many points")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    c = CONFIGS[a.config]
    scene = make_scene(c["P"], c["W"], c["H"], F=c["F"], seed=0, s_med=c["s_med"]).to(dev)
    cam = make_camera(c["W"], c["H"]).to(dev)
    grads = make_upstream_grads(c["W"], c["H"], c["F"], device=dev)
    fargs = hz.native_forward_args(scene, cam, torch.zeros(3, device=dev), c["F"], all_map=make_all_map(scene, cam))
    for _ in range(a.iters):
        fwd, bwd = hz.run_native(ops, fargs, grads)
    torch.cuda.synchronize()
    print("R", fwd["num_rendered"], "checksum", float(fwd["color"].sum()), float(bwd["means3D"].abs().sum()))
    if a.knn:
        pts = make_scene(a.knn, 1920, 1080, seed=3).means3D.to(dev)
        d = ops.distCUDA2(pts)
        torch.cuda.synchronize()
        print("knn", float(d.sum()))


if __name__ == "__main__":
    main()
