"""A few single-view fwd+bwd iterations of one BASELINE config for ncu (launch list / --set full):  python tools/profile_iter.py C3 [iters]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from lsx_b200 import ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c, scene, cam, grads, bg, am, fargs = bench.build_case(name, torch.device("cuda:0"))
step = bench.native_stepper(ops, fargs, grads)
for _ in range(iters):
    step()
torch.cuda.synchronize()
print("profile_iter done", name, iters)
