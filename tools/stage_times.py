"""Per-stage device times (CUDA events recorded by the library) of single-view fwd+bwd on the BASELINE configs."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from lsx_b200 import _lib, ops  # noqa: E402

for name in (sys.argv[1:] or ["C3", "C4"]):
    dev = torch.device("cuda:0")
    c, scene, cam, grads, bg, am, fargs = bench.build_case(name, dev)
    step = bench.native_stepper(ops, fargs, grads)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    _lib.profile_read()
    n = 20
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    st = {k: round(v / n, 4) for k, v in _lib.profile_read().items() if v > 0}
    _lib.profile_enable(False)
    print(json.dumps({"config": name, "sum_ms": round(sum(st.values()), 4), "stages_ms": st}), flush=True)
    del scene, cam, grads, am, fargs
    torch.cuda.empty_cache()
