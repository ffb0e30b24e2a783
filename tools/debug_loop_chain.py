"""Where do FieldLoop's hand-chained gradients differ from the autograd composition of the same operators?  One view:
image-space gradients, rasterizer backward outputs, final arena groups — each compared (rms / tensor-scale)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench_loop as bl  # noqa: E402
from diff_LangSurf_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from lsx_b200 import loss as L, render_utils as RU  # noqa: E402
from lsx_b200.field_loop import FieldLoop, LoopConfig  # noqa: E402
from lsx_b200.synthetic import make_scene  # noqa: E402

DEV = "cuda:0"
P, W, H, F = 20_000, 208, 144, 3
scene = make_scene(P, W, H, F=F, seed=5).to(DEV)
raw = bl.make_raw(scene)
vw = bl.make_view(1, 3, W, H, F, DEV)
cfg = LoopConfig(optimise_pose=False, cls3d=False)
bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)


def err(a, b):
    a, b = a.double(), b.double()
    d = (a - b)
    return f"rms {float(d.pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-300)):.2e} scale {float(d.abs().max() / (b.abs().max() + 1e-300)):.2e}"


def manual():
    loop = FieldLoop(raw, bl.LRS, bg, cfg)
    loop.debug_tap = []
    loop.gradient([vw], None)
    return loop.debug_tap[0], {k: v.clone() for k, v in loop.grads.views.items()}


def auto():
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}
    scales, rots, opac, all_map = RU.gaussian_head(leaf["means3D"], leaf["scales"], leaf["rotations"], leaf["opacity"], vw.viewmatrix, vw.campos)
    inter = dict(scales=scales, rots=rots, opac=opac, all_map=all_map)
    for t in inter.values():
        t.retain_grad()
    s = GaussianRasterizationSettings(H, W, vw.tanfovx, vw.tanfovy, bg, 1.0, vw.viewmatrix, vw.projmatrix, 3, vw.campos, False, True, False, True)
    m2 = torch.zeros(P, 3, device=DEV, requires_grad=True)
    xyz_in = leaf["means3D"] * 1.0
    xyz_in.retain_grad()
    color, lf, li, radii, obs, amap, depth = GaussianRasterizer(s)(
        means3D=xyz_in, means2D=m2, means2D_abs=m2, opacities=opac, shs=leaf["sh"].view(P, -1, 3), language_feature_precomp=leaf["language_feature"],
        language_feature_instance_precomp=leaf["instance_feature"], scales=scales, rotations=rots, all_map=all_map)
    for t in (color, lf, amap, depth):
        t.retain_grad()
    loss = L.image_loss(color, vw.gt_image, cfg.lambda_dssim)[0]
    dn = RU.depth_to_normal(depth[0], vw.fx, vw.fy, W * 0.5, H * 0.5, alpha=amap[3])
    loss = loss + cfg.normal_weight * (vw.image_weight * (dn - amap[:3]).abs().sum(0)).mean()
    loss = loss + L.masked_l1_loss(lf, vw.gt_language, vw.language_mask)
    loss.backward()
    tap = dict(color=color, lang=lf, amap=amap, depth=depth, g_color=color.grad, g_lang=lf.grad, g_amap=amap.grad, g_depth=depth.grad,
               g_m3d=xyz_in.grad, g_scales=scales.grad, g_rot=rots.grad, g_opac=opac.grad, g_allmap=all_map.grad)
    return tap, {k: v.grad for k, v in leaf.items()}


m_tap, m_g = manual()
m_tap2, m_g2 = manual()
a_tap, a_g = auto()
a_tap2, a_g2 = auto()
print("== intermediates: manual vs autograd | manual vs manual | autograd vs autograd")
for k in a_tap:
    print(f"{k:10s} {err(m_tap[k], a_tap[k]):34s} | {err(m_tap2[k], m_tap[k]):34s} | {err(a_tap2[k], a_tap[k])}")
print("== parameter gradients")
for k in a_g:
    if a_g[k] is None:
        continue
    v = m_g[k].reshape(a_g[k].shape)
    print(f"{k:18s} {err(v, a_g[k]):34s} | {err(m_g2[k].reshape(a_g[k].shape), v):34s} | {err(a_g2[k], a_g[k])}")
