"""distCUDA2: new library vs the recompiled reference (oracle/_ref/ref_knn.so) — bit-exactness and time.
Usage (GPU box): python tools/knn_bench.py [P ...]"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402
from lsx_b200.synthetic import make_scene  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(n):
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100_000, 1_000_000, 5_000_000]
    ref = hz.load_ref("ref_knn")
    for P in sizes:
        pts = make_scene(P, 1920, 1080, seed=3).means3D.cuda().contiguous()
        t_new, d_new = timeit(lambda: ops.distCUDA2(pts))
        row = {"P": P, "new_ms": t_new}
        if ref is not None:
            t_ref, d_ref = timeit(lambda: ref.distCUDA2(pts), n=2 if P > 2_000_000 else 3)
            row.update(ref_ms=t_ref, speedup=t_ref / t_new,
                       bit_mismatches=int((d_new.view(torch.int32) != d_ref.view(torch.int32)).sum()))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
