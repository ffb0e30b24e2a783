"""fwd+bwd ms/iter of the new operator and the recompiled reference on every BASELINE config (single view each)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import harness as hz  # noqa: E402
from lsx_b200 import ops  # noqa: E402


def timeit(step, n):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name in (sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]):
    dev = torch.device("cuda:0")
    c, scene, cam, grads, bg, am, fargs = bench.build_case(name, dev)
    ref = hz.ref_rast_for(c["F"])
    fwd, _ = bench.native_stepper(ops, fargs, grads)()
    row = {"config": name, "P": c["P"], "W": c["W"], "H": c["H"], "F": c["F"], "R": int(fwd["num_rendered"]),
           "P_vis": int((fwd["radii"] > 0).sum())}
    st = ops.render_stats(fwd["num_rendered"], fwd["geom"], fwd["binning"], fwd["img"], c["P"], c["H"], c["W"], 3 + c["F"] + 3 + 5)
    row.update({k: st[k] for k in ("S", "B", "V", "Vb", "L", "Hmax", "Hsum")})
    row["lane_utilisation_in_blending_visits"] = round(st["B"] / max(1, 32 * st["Vb"]), 3)
    del fwd
    n = 20 if c["P"] <= 1_000_000 else 5
    row["new_ms"] = timeit(bench.native_stepper(ops, fargs, grads), n)
    if ref is not None:
        row["ref_ms"] = timeit(bench.native_stepper(ref, fargs, grads), max(3, n // 4))
        row["speedup"] = row["ref_ms"] / row["new_ms"]
    row["new_MPix_s"] = c["W"] * c["H"] / row["new_ms"] / 1e3
    print(json.dumps(row), flush=True)
    del scene, cam, grads, am, fargs
    torch.cuda.empty_cache()
