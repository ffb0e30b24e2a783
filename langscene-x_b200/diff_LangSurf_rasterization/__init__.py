"""Drop-in replacement for the reference package `diff_LangSurf_rasterization`
(diff-langsurf-rasterizer/diff_LangSurf_rasterization/__init__.py).

Public surface kept identical:
  GaussianRasterizationSettings   14-field NamedTuple, same order            (reference :189-203)
  GaussianRasterizer(settings)    .forward(...) keyword names, 7-tuple result, .markVisible(positions)  (:205-264)
  rasterize_gaussians(...)        functional form                            (:21-50)
Gradients are produced for means3D, means2D, means2D_abs, sh, colors_precomp, language_feature_precomp,
language_feature_instance_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map (:171-185).

Differences (all behind the same outputs):
  * the language-feature width is taken from the tensor at run time (the reference bakes 3 into config.h);
  * `debug=True` works (the reference's debug branch unpacks the wrong tuple arity, :103): every stage is
    synchronised and checked, and on failure the arguments are dumped to snapshot_fw.dump / snapshot_bw.dump;
  * work is enqueued on torch's current stream (the reference uses the legacy default stream);
  * multi-view extension (keyword-only, absent in the reference): `GaussianRasterizer(settings, grad_buffers=, accumulate=)`
    lets the backward kernel write / add the per-Gaussian parameter gradients straight into caller-owned buffers (e.g. the
    views of lsx_b200.multiview.GradArena); autograd then receives None for those inputs.  Use it when the rasterizer's
    inputs ARE the leaves (no wrapper in between), as in a view-sharded gradient accumulation.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    render_geo: bool
    debug: bool
    include_feature: bool


def _snapshot(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def _call_native(fn, args, debug, dump_name, what):
    if not debug:
        return fn(*args)
    saved = _snapshot(args)  # taken before the call so a fault cannot corrupt it
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_name)
        print(f"\nAn error occured in {what}. Arguments written to {dump_name} for debugging.")
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, means2D_abs, sh, colors_precomp, language_feature_precomp,
                language_feature_instance_precomp, opacities, scales, rotations, cov3Ds_precomp, all_maps,
                raster_settings, grad_sink=None):
        s = raster_settings
        ctx.grad_sink = grad_sink
        native_args = (
            s.bg, means3D, colors_precomp, language_feature_precomp, language_feature_instance_precomp, opacities,
            scales, rotations, s.scale_modifier, cov3Ds_precomp, all_maps, s.viewmatrix, s.projmatrix,
            s.tanfovx, s.tanfovy, s.image_height, s.image_width, sh, s.sh_degree, s.campos,
            s.prefiltered, s.render_geo, s.debug, s.include_feature,
        )
        (num_rendered, color, language_feature, language_feature_instance, radii, out_observe, out_all_map,
         out_plane_depth, geomBuffer, binningBuffer, imgBuffer) = _call_native(
            _C.rasterize_gaussians, native_args, s.debug, "snapshot_fw.dump", "forward")

        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(out_all_map, colors_precomp, language_feature_precomp,
                              language_feature_instance_precomp, all_maps, means3D, scales, rotations,
                              cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii, out_observe)
        return color, language_feature, language_feature_instance, radii, out_observe, out_all_map, out_plane_depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_language_feature, grad_out_language_feature_instance, grad_radii,
                 grad_out_observe, grad_out_all_map, grad_out_plane_depth):
        s = ctx.raster_settings
        (all_map_pixels, colors_precomp, language_feature_precomp, language_feature_instance_precomp, all_maps,
         means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        native_args = (
            s.bg, all_map_pixels, means3D, radii, colors_precomp, language_feature_precomp,
            language_feature_instance_precomp, all_maps, scales, rotations, s.scale_modifier, cov3Ds_precomp,
            s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, grad_out_color, grad_out_language_feature,
            grad_out_language_feature_instance, grad_out_all_map, grad_out_plane_depth, sh, s.sh_degree, s.campos,
            geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, s.render_geo, s.debug, s.include_feature,
        )
        (grad_means2D, grad_means2D_abs, grad_colors_precomp, grad_language_feature_precomp,
         grad_language_feature_instance_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
         grad_scales, grad_rotations, grad_all_map) = _call_native(
            _sunk_backward(ctx.grad_sink), native_args, s.debug, "snapshot_bw.dump", "backward")
        sunk = set(ctx.grad_sink[0]) if ctx.grad_sink else ()

        def _for(inp, g, name=None):
            # placeholders (empty CPU tensors standing for "absent") never require grad; give autograd None.  Gradients the
            # kernel already wrote into the caller's buffers are not handed to autograd a second time.
            if name in sunk:
                return None
            return g if (inp is not None and inp.numel() != 0) else None

        return (
            _for(means3D, grad_means3D, "means3D"), grad_means2D, grad_means2D_abs,
            _for(sh, grad_sh, "sh"), _for(colors_precomp, grad_colors_precomp, "colors"),
            _for(language_feature_precomp, grad_language_feature_precomp, "language_feature") if s.include_feature else None,
            _for(language_feature_instance_precomp, grad_language_feature_instance_precomp, "instance_feature")
            if s.include_feature else None,
            _for(means3D, grad_opacities, "opacity"), _for(scales, grad_scales, "scales"),   # opacities: always present
            _for(rotations, grad_rotations, "rotations"),
            _for(cov3Ds_precomp, grad_cov3Ds_precomp, "cov3D"),
            _for(all_maps, grad_all_map, "all_map") if s.render_geo else None,
            None, None,
        )


def _sunk_backward(grad_sink):
    if not grad_sink:
        return _C.rasterize_gaussians_backward
    buffers, accumulate = grad_sink

    def call(*args):
        return _C.rasterize_gaussians_backward(*args, grad_buffers=buffers, accumulate=accumulate)
    return call


def rasterize_gaussians(means3D, means2D, means2D_abs, sh, colors_precomp, language_feature_precomp,
                        language_feature_instance_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map,
                        raster_settings, grad_sink=None):
    return _RasterizeGaussians.apply(means3D, means2D, means2D_abs, sh, colors_precomp, language_feature_precomp,
                                     language_feature_instance_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     all_map, raster_settings, grad_sink)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, *, grad_buffers=None, accumulate=False):
        super().__init__()
        self.raster_settings = raster_settings
        # multi-view extension: see the module docstring and lsx_b200.ops.rasterize_gaussians_backward
        self.grad_buffers = grad_buffers
        self.accumulate = accumulate

    def markVisible(self, positions):
        """Boolean mask of the points in front of the near plane (view-space z > 0.2)."""
        with torch.no_grad():
            s = self.raster_settings
            return _C.mark_visible(positions, s.viewmatrix, s.projmatrix)

    def forward(self, means3D, means2D, means2D_abs, opacities, shs=None, colors_precomp=None,
                language_feature_precomp=None, language_feature_instance_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, all_map=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (has_sr and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        def _absent(t):
            return torch.Tensor([]) if t is None else t  # empty CPU tensor == "not provided", as in the reference

        return rasterize_gaussians(
            means3D, means2D, means2D_abs, _absent(shs), _absent(colors_precomp),
            _absent(language_feature_precomp), _absent(language_feature_instance_precomp), opacities,
            _absent(scales), _absent(rotations), _absent(cov3D_precomp), _absent(all_map), self.raster_settings,
            (self.grad_buffers, bool(self.accumulate)) if self.grad_buffers else None)
