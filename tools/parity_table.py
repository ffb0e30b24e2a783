"""Per-tensor parity table of the rasterizer against the UNMODIFIED reference CUDA rasterizer (oracle/_ref) on the BASELINE
configs C1..C5 (single view) and on a near-transparent scene:  python tools/parity_table.py [C1 C2 ...] > profiles/parity_table.md

For every output / gradient tensor three error statistics of (new - reference), each next to the reference's OWN
run-to-run spread in the same statistic (the reference accumulates gradients with fp32 atomics in arbitrary order):
    scale   max|a-b| / max|b|                      tensor-scale: what BASELINE.json's 1e-5 / 1e-4 are quoted in
    mixed   max |a-b| / (|b| + rms(b))             element-wise, the tensor's rms as the floor for small elements
    rms     rms(a-b) / rms(b)
Integer / index outputs are compared bit for bit."""
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

DEV = torch.device("cuda:0")


def stats(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    if b.numel() == 0 or float(b.abs().max()) == 0.0:
        return (float((a - b).abs().max()) if b.numel() else 0.0,) * 3
    rms = float(b.pow(2).mean().sqrt())
    d = (a - b).abs()
    return float(d.max() / b.abs().max()), float((d / (b.abs() + rms)).max()), float(d.pow(2).mean().sqrt() / rms)


def fmt(x):
    return "0" if x == 0 else f"{x:.1e}"


def run_case(title, scene, cam, grads, F, W, H, P):
    ref = hz.ref_rast_for(F)
    bg = torch.zeros(3, device=DEV)
    fargs = hz.native_forward_args(scene, cam, bg, F, all_map=make_all_map(scene, cam))
    rf, rb = hz.run_native(ref, fargs, grads)
    rf2, rb2 = hz.run_native(ref, fargs, grads)
    nf, nb = hz.run_native(ops, fargs, grads)
    torch.cuda.synchronize()
    R, Ct = rf["num_rendered"], 3 + F + 3 + 5
    rbuf = hz.parse_ref_buffers(rf["geom"], rf["binning"], rf["img"], P, R, W, H)
    nbuf = hz.parse_new_buffers(nf["geom"], nf["binning"], nf["img"], P, R, W, H, Ct)
    vis = rf["radii"] > 0
    beq = lambda a, b: bool((a.contiguous().view(torch.int32) == b.contiguous().view(torch.int32)).all())
    exact = {"num_rendered": nf["num_rendered"] == R, "radii": torch.equal(rf["radii"], nf["radii"]),
             "tiles_touched": torch.equal(rbuf["tiles_touched"], nbuf["tiles_touched"]),
             "depths": beq(rbuf["depths"][vis], nbuf["depths"][vis]), "means2D": beq(rbuf["means2D"][vis], nbuf["means2D"][vis]),
             "conic_opacity": beq(rbuf["conic_opacity"][vis], nbuf["conic_opacity"][vis]),
             "sorted 64-bit keys": torch.equal(rbuf["keys"], nbuf["keys"]), "point_list": torch.equal(rbuf["point_list"], nbuf["point_list"]),
             "tile ranges": torch.equal(rbuf["ranges"], nbuf["ranges"]), "n_contrib": torch.equal(rbuf["n_contrib"], nbuf["n_contrib"]),
             "final_T": beq(rbuf["final_T"], nbuf["final_T"]), "out_observe": torch.equal(rf["out_observe"], nf["out_observe"])}
    print(f"\n## {title}\n\nP = {P}, {W}x{H}, F = {F}, R = {R}, visible = {int(vis.sum())}\n")
    print("bit-exact: " + ", ".join(f"{k} {'✓' if v else '✗'}" for k, v in exact.items()) + "\n")
    print("| tensor | scale (new vs ref) | scale (ref vs ref) | mixed | mixed (ref vs ref) | rms | rms (ref vs ref) |")
    print("|---|---|---|---|---|---|---|")
    for k in ("color", "language_feature", "instance_feature", "all_map", "plane_depth"):
        s, s0 = stats(nf[k], rf[k]), stats(rf2[k], rf[k])
        print(f"| out {k} | {fmt(s[0])} | {fmt(s0[0])} | {fmt(s[1])} | {fmt(s0[1])} | {fmt(s[2])} | {fmt(s0[2])} |")
    for k in hz.BWD_NAMES:
        s, s0 = stats(nb[k], rb[k]), stats(rb2[k], rb[k])
        print(f"| grad {k} | {fmt(s[0])} | {fmt(s0[0])} | {fmt(s[1])} | {fmt(s0[1])} | {fmt(s[2])} | {fmt(s0[2])} |")
    sys.stdout.flush()


def main():
    names = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]
    print("# Parity of the rasterizer against the reference CUDA rasterizer, tensor by tensor\n")
    print(__doc__.split("\n\n", 1)[1])
    for name in names:
        c = CONFIGS[name]
        P, W, H, F = c["P"], c["W"], c["H"], c["F"]
        scene = make_scene(P, W, H, F=F, seed=0, s_med=c["s_med"]).to(DEV)
        cam = make_camera(W, H, yaw_deg=0.0).to(DEV)
        grads = make_upstream_grads(W, H, F, device=DEV)
        run_case(f"{name}", scene, cam, grads, F, W, H, P)
        del scene, cam, grads
        torch.cuda.empty_cache()
    # near-transparent scene (the footprint-culling stress case of tests/test_parity_gpu.py): gradients WITH the plane-depth
    # upstream gradient zeroed — through 1/<n, ray> ~ 1e3 it turns both implementations' gradients into amplified noise
    P, W, H, F = 20_000, 256, 192, 3
    scene = make_scene(P, W, H, F=F, seed=42, s_med=0.003).to(DEV)
    scene.opacities = torch.sigmoid(torch.logit(scene.opacities.clamp(1e-4, 1 - 1e-4)) - 3.5).contiguous()
    grads = make_upstream_grads(W, H, F, seed=43, device=DEV)
    grads["plane_depth"] = torch.zeros_like(grads["plane_depth"])
    run_case("near-transparent scene (opacity logits - 3.5), dL/dplane_depth = 0", scene, make_camera(W, H).to(DEV), grads, F, W, H, P)


if __name__ == "__main__":
    main()
