"""First-light GPU check: new operator vs the recompiled reference on identical inputs + rough timings.
Usage (on a GPU box): python tools/first_light.py [C1 C2 C3 ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as hz  # noqa: E402
from lsx_b200 import ops as new_ops, _lib  # noqa: E402
from lsx_b200.synthetic import CONFIGS, make_camera, make_scene, make_upstream_grads  # noqa: E402


def compare(name, cfg, timing_iters=0):
    dev = torch.device("cuda:0")
    P, W, H, F = cfg["P"], cfg["W"], cfg["H"], cfg["F"]
    scene = make_scene(P, W, H, F=F, seed=0, s_med=cfg["s_med"]).to(dev)
    cam = make_camera(W, H).to(dev)
    bg = torch.zeros(3, device=dev)
    grads = make_upstream_grads(W, H, F, device=dev)
    fargs = hz.native_forward_args(scene, cam, bg, F)
    ref = hz.ref_rast_for(F)
    res = {"config": name, "P": P, "W": W, "H": H, "F": F}
    nf, nb = hz.run_native(new_ops, fargs, grads)
    torch.cuda.synchronize()
    res["R_new"] = nf["num_rendered"]
    if ref is not None:
        rf, rb = hz.run_native(ref, fargs, grads)
        torch.cuda.synchronize()
        R = rf["num_rendered"]
        res["R_ref"] = R
        nch = 3 + F + 3 + 5
        rbuf = hz.parse_ref_buffers(rf["geom"], rf["binning"], rf["img"], P, R, W, H)
        nbuf = hz.parse_new_buffers(nf["geom"], nf["binning"], nf["img"], P, nf["num_rendered"], W, H, nch)
        vis = rf["radii"] > 0
        res["P_vis"] = int(vis.sum())
        res["radii_mismatch"] = int((rf["radii"] != nf["radii"]).sum())
        res["tiles_mismatch"] = int((rbuf["tiles_touched"] != nbuf["tiles_touched"]).sum())
        for k in ["depths", "means2D", "conic_opacity", "rgb", "cov3D"]:
            a, b = rbuf[k][vis], nbuf[k][vis]
            res[k + "_bit_mismatch"] = int((a.view(torch.int32) != b.view(torch.int32)).sum())
        res["clamped_mismatch"] = int((rbuf["clamped"][vis] != nbuf["clamped"][vis]).sum())
        if R == nf["num_rendered"] and R > 0:
            res["keys_mismatch"] = int((rbuf["keys"] != nbuf["keys"]).sum())
            res["point_list_mismatch"] = int((rbuf["point_list"] != nbuf["point_list"]).sum())
        res["ranges_mismatch"] = int((rbuf["ranges"] != nbuf["ranges"]).sum())
        res["n_contrib_mismatch"] = int((rbuf["n_contrib"] != nbuf["n_contrib"]).sum())
        res["final_T_bit_mismatch"] = int((rbuf["final_T"].view(torch.int32) != nbuf["final_T"].view(torch.int32)).sum())
        res["out_observe_mismatch"] = int((rf["out_observe"] != nf["out_observe"]).sum())
        res["S"] = int(rbuf["n_contrib"].long().sum())
        for k in ["color", "language_feature", "instance_feature", "all_map", "plane_depth"]:
            res["fwd_rel_" + k] = hz.rel_err(nf[k], rf[k])
        for k in hz.BWD_NAMES:
            if rb[k].numel() > 1:
                res["bwd_rel_" + k] = hz.rel_err(nb[k], rb[k])
        # reference's own run-to-run spread of the backward
        _, rb2 = hz.run_native(ref, fargs, grads)
        res["ref_self_spread_means2D"] = hz.rel_err(rb2["means2D"], rb["means2D"])
        res["ref_self_spread_lang"] = hz.rel_err(rb2["language_feature"], rb["language_feature"])
    if timing_iters:
        for label, mod in (("new", new_ops), ("ref", ref)):
            if mod is None:
                continue
            for _ in range(3):
                hz.run_native(mod, fargs, grads)
            torch.cuda.synchronize()
            tf, tb = [], []
            for _ in range(timing_iters):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*fargs)))
                e1.record()
                mod.rasterize_gaussians_backward(*hz.native_backward_args(fargs, fwd, grads))
                e2.record()
                torch.cuda.synchronize()
                tf.append(e0.elapsed_time(e1))
                tb.append(e1.elapsed_time(e2))
            tf.sort(); tb.sort()
            res[f"ms_fwd_{label}"] = tf[len(tf) // 2]
            res[f"ms_bwd_{label}"] = tb[len(tb) // 2]
    return res


if __name__ == "__main__":
    names = sys.argv[1:] or ["C1", "C2", "C3"]
    os.makedirs("gpurun_out", exist_ok=True)
    out = []
    for n in names:
        t0 = time.time()
        r = compare(n, CONFIGS[n], timing_iters=10)
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)
        out.append(r)
        with open("gpurun_out/first_light.json", "w") as f:
            json.dump(out, f, indent=1)
    print("launches", _lib.kernel_launch_count())
