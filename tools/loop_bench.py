"""Skeleton of LangScene-X's optimisation iteration (render wrapper -> rasterizer -> depth normal + losses -> backward -> Adam,
field_construction/gaussian_field.py:184-543 reduced to its compute) on a BASELINE-config-4-shaped synthetic scene:
    reference-style : reference CUDA rasterizer (oracle/_ref) + the torch-op wrapper / losses the reference uses + torch.optim.Adam
    new             : this repository's rasterizer + fused head / depth-normal / image-loss kernels + ArenaAdam
Same parameters, same camera, same targets.  A measurement harness (tool), not a trainer."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import harness as hz  # noqa: E402
from lsx_b200 import loss as fused_loss, render_utils  # noqa: E402
from lsx_b200.multiview import PARAM_GROUPS, GradArena  # noqa: E402
from lsx_b200.optim import ArenaAdam  # noqa: E402
from lsx_b200.synthetic import CONFIGS, make_camera, make_scene  # noqa: E402
from oracle import depth_normal_oracle, gaussian_head_oracle, image_loss_oracle  # noqa: E402  (tool: reference-style torch ops)

LRS = {"means3D": 1.6e-4, "sh": 2.5e-3, "opacity": 5e-2, "scales": 5e-3, "rotations": 1e-3, "language_feature": 2.5e-3,
       "instance_feature": 2.5e-3, "all_map": 0.0}


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def make_rasterizer(style, F):
    if style == "new":
        from diff_LangSurf_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        return GaussianRasterizationSettings, GaussianRasterizer
    # the reference's Python operator layer restated around its own pybind module (its package cannot be imported as a
    # package here: both feature widths are built as single .so files)
    mod = hz.ref_rast_for(F)
    from diff_LangSurf_rasterization import GaussianRasterizationSettings

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, sh, lang, inst, opac, scales, rots, all_map, s):
            e = torch.Tensor([])
            args = (s.bg, means3D, e, lang, inst, opac, scales, rots, s.scale_modifier, e, all_map, s.viewmatrix, s.projmatrix,
                    s.tanfovx, s.tanfovy, s.image_height, s.image_width, sh, s.sh_degree, s.campos, False, True, False, True)
            (R, color, lf, li, radii, obs, amap, depth, geom, binning, img) = mod.rasterize_gaussians(*args)
            ctx.s, ctx.R = s, R
            ctx.save_for_backward(amap, lang, inst, all_map, means3D, scales, rots, radii, sh, geom, binning, img)
            return color, lf, li, amap, depth

        @staticmethod
        def backward(ctx, gc, glf, gli, gam, gd):
            s = ctx.s
            amap, lang, inst, all_map, means3D, scales, rots, radii, sh, geom, binning, img = ctx.saved_tensors
            e = torch.Tensor([])
            out = mod.rasterize_gaussians_backward(s.bg, amap, means3D, radii, e, lang, inst, all_map, scales, rots, s.scale_modifier,
                                                   e, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, gc, glf, gli, gam, gd, sh,
                                                   s.sh_degree, s.campos, geom, ctx.R, binning, img, True, False, True)
            (g2d, g2da, gcol, glang, ginst, gop, gm3, gcov, gsh, gsc, grot, gall) = out
            return gm3, gsh, glang, ginst, gop, gsc, grot, gall, None

    class _Rast:
        def __init__(self, s):
            self.s = s

        def __call__(self, means3D, means2D, means2D_abs, opacities, shs, language_feature_precomp,
                     language_feature_instance_precomp, scales, rotations, all_map):
            c, lf, li, am, d = _Fn.apply(means3D, shs, language_feature_precomp, language_feature_instance_precomp, opacities,
                                         scales, rotations, all_map, self.s)
            return c, lf, li, None, None, am, d
    return GaussianRasterizationSettings, _Rast


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    c = CONFIGS[a.config]
    P, W, H, F = c["P"], c["W"], c["H"], c["F"]
    scene = make_scene(P, W, H, F=F, seed=0, s_med=c["s_med"]).to(dev)
    cam = make_camera(W, H, yaw_deg=3.0).to(dev)
    fx, fy = W / (2 * cam.tanfovx), H / (2 * cam.tanfovy)
    g = torch.Generator().manual_seed(1)
    gt_image = torch.rand(3, H, W, generator=g).to(dev)
    gt_lang = torch.rand(F, H, W, generator=g).to(dev)
    bg = torch.zeros(3, device=dev)
    raw0 = {"means3D": scene.means3D, "sh": scene.shs.reshape(P, -1), "opacity": torch.logit(scene.opacities.clamp(1e-4, 1 - 1e-4)),
            "scales": torch.log(scene.scales), "rotations": scene.rotations, "language_feature": scene.language_feature,
            "instance_feature": scene.instance_feature}
    results = {}
    for style in ("reference-style", "new"):
        if style == "reference-style" and hz.ref_rast_for(F) is None:
            continue
        Settings, Rast = make_rasterizer("new" if style == "new" else "ref", F)
        head = render_utils.gaussian_head if style == "new" else gaussian_head_oracle.gaussian_head
        if style == "new":
            params = GradArena.allocate(P, 16, F, 3, dev)
            grads = GradArena.allocate(P, 16, F, 3, dev)
            for n, v in raw0.items():
                params.views[n].copy_(v.reshape(params.views[n].shape))
            leaf = {n: params.views[n].detach().requires_grad_(True) for n in raw0}      # parameters ARE arena views
            opt = ArenaAdam(params, LRS)
        else:
            leaf = {n: v.detach().clone().requires_grad_(True) for n, v in raw0.items()}
            opt = torch.optim.Adam([{"params": [leaf[n]], "lr": LRS[n]} for n in leaf], lr=0.0, eps=1e-15)

        def iteration():
            if style == "new":
                grads.flat.zero_()
                for n in leaf:
                    leaf[n].grad = grads.views[n].view(leaf[n].shape)                      # autograd accumulates in place
            else:
                opt.zero_grad(set_to_none=True)
            scales, rots, opac, all_map = head(leaf["means3D"], leaf["scales"], leaf["rotations"], leaf["opacity"], cam.viewmatrix,
                                               cam.campos)
            s = Settings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix, cam.projmatrix, 3, cam.campos, False, True, False,
                         True)
            m2 = torch.zeros_like(leaf["means3D"], requires_grad=True)
            out = Rast(s)(means3D=leaf["means3D"], means2D=m2, means2D_abs=m2, opacities=opac, shs=leaf["sh"].view(P, 16, 3),
                          language_feature_precomp=leaf["language_feature"],
                          language_feature_instance_precomp=leaf["instance_feature"], scales=scales, rotations=rots, all_map=all_map)
            color, lf, li, _, _, amap, depth = out
            if style == "new":
                dn = render_utils.depth_to_normal(depth[0], fx, fy, W / 2, H / 2, alpha=amap[3])
                img_loss = fused_loss.image_loss(color, gt_image, 0.2)[0]
            else:
                dn = depth_normal_oracle.depth_to_normal(depth[0], fx, fy, W / 2, H / 2, amap[3])
                img_loss = image_loss_oracle.image_loss(color, gt_image, 0.2)
            loss = img_loss + 0.015 * (amap[:3] - dn).abs().sum(0).mean() + torch.abs(lf - gt_lang).mean()
            loss.backward()
            if style == "new":
                opt.step(grads)
            else:
                opt.step()
            return loss

        ms = timeit(iteration, a.iters)
        results[style] = {"ms_per_iter": ms, "loss": float(iteration().detach())}
    if len(results) == 2:
        results["speedup"] = results["reference-style"]["ms_per_iter"] / results["new"]["ms_per_iter"]
    print(json.dumps({"config": a.config, "P": P, "W": W, "H": H, "F": F, **results}))


if __name__ == "__main__":
    main()
