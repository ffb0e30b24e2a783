"""Tensor-level entry points with the reference's `_C` signatures, implemented on the C ABI.

Mirrors (argument order, return tuples, error behaviour)
  diff_LangSurf_rasterization._C.rasterize_gaussians            rasterize_points.cu:35-143   (24 args -> 11-tuple)
  diff_LangSurf_rasterization._C.rasterize_gaussians_backward   rasterize_points.cu:145-259  (31 args -> 12-tuple)
  diff_LangSurf_rasterization._C.mark_visible                   rasterize_points.cu:261-280
  simple_knn._C.distCUDA2                                       simple-knn/spatial.cu:15-25

What this layer does: tensor allocation, `.contiguous()`, current-stream lookup, raw pointers into the
ctypes argument blocks, status -> RuntimeError.  Nothing is computed here.
"""
import ctypes

import torch

from . import _lib

_FLOAT = torch.float32


def _ptr(t):
    """Device pointer of a tensor, or None ("absent") for the reference's empty placeholders."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t, name, device):
    """contiguous fp32 CUDA tensor on `device`, or the empty placeholder untouched."""
    if t is None or t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this operator has no CPU path)")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, expected {device}")
    if t.dtype != _FLOAT:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _aligned16(t):
    """The kernels read `rotations` as one 16-B word per Gaussian (and write its gradient the same way): a contiguous tensor
    that starts at an odd float of its storage is copied to a fresh allocation (the C ABI rejects it, include/lsx_rasterizer.h)."""
    if t is not None and t.numel() and t.data_ptr() % 16:
        return t.clone()
    return t


class _ScratchAlloc:
    """The resizeFunctional analogue (rasterize_points.cu:27-33): torch owns the scratch bytes.

    The ctypes callback must not capture `self`: self -> fn -> closure -> self would be a reference cycle, and the
    scratch tensor (hundreds of MB) would then survive until CPython's cyclic collector runs instead of being
    released by reference counting when the caller drops it."""

    def __init__(self, device):
        holder, failure = [], []

        def _cb(_user, nbytes):
            # ctypes swallows exceptions raised inside a callback (the C side would just see NULL): keep the exception —
            # typically torch.cuda.OutOfMemoryError, which training loops catch to empty the cache and retry — and let
            # `reraise()` surface it with its type once the library has returned its "allocation failed" status.
            try:
                t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            except Exception as e:  # noqa: BLE001
                failure[:] = [e]
                return 0
            holder[:] = [t]
            return t.data_ptr()

        self._holder = holder
        self._failure = failure
        self._device = device
        self.fn = _lib.ALLOC_FN(_cb)

    @property
    def tensor(self):
        return self._holder[0] if self._holder else torch.empty(0, dtype=torch.uint8, device=self._device)

    def reraise(self):
        if self._failure:
            raise self._failure[0]


def _check(status, what, *allocs):
    """status -> exception; an allocation failure inside a scratch callback is re-raised with its original type."""
    if status != 0:
        for a in allocs:
            a.reraise()
    _lib.check(status, what)


# ---- speculative binning capacity (lsx_forward_args.binning_capacity_hint), OFF by default -------------------------------
# The list length of a forward call is close to that of the previous calls on the same scene (same Gaussian count, same image
# size): with LSX_SPECULATIVE_BINNING=1, 1.25 x the largest of the last few is handed to the library as the capacity, so that
# it need not drain the stream in the middle of the forward pass to learn the exact length.  A wrong guess costs one repeated
# binning pass inside the library, never a wrong result (tests/test_parity_gpu.py).  Measured (profiles/r6f_sweep_*.jsonl):
# no gain in a training-shaped loop — the host is blocked in that read while the GPU is still busy with the previous
# backward pass, so the bubble is one kernel-launch latency, and the 25 % of padded list slots cost as much in the tile sort:
# C3 3.761 vs 3.758 ms, C4 0.954 vs 0.960, C1 0.360 vs 0.321.  Kept for callers whose stream is otherwise empty at that point.
import os as _os

_SPECULATE = _os.environ.get("LSX_SPECULATIVE_BINNING", "0") == "1"
_recent_rendered = {}


def _capacity_hint(key):
    hist = _recent_rendered.get(key)
    if not _SPECULATE or not hist:
        return 0
    return min(int(1.25 * max(hist)) + 4096, 0x7fffff00)


def _remember_rendered(key, rendered):
    hist = _recent_rendered.setdefault(key, [])
    hist.append(int(rendered))
    del hist[:-4]


def _stream_handle(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def rasterize_gaussians(
    background, means3D, colors, language_feature, language_feature_instance, opacity, scales, rotations,
    scale_modifier, cov3D_precomp, all_map, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
    sh, degree, campos, prefiltered, render_geo, debug, include_feature,
    *, raw_params=False, pose=None,
):
    """Positional signature of the reference's `_C.rasterize_gaussians`.  Keyword-only extension (fused render wrapper,
    lsx_forward_args.raw_params): with `raw_params=True`, `means3D / opacity / scales / rotations` are the reference's RAW
    parameters (positions, opacity logits, log-scales, un-normalised quaternions), `all_map` and `cov3D_precomp` must be empty,
    `pose` is an optional (7,) camera pose [quaternion | translation]; the per-Gaussian kernel applies pose transform,
    activations, plane normal and all_map itself."""
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (this operator has no CPU path)")
    device = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)

    means3D = _prep(means3D, "means3D", device)
    background = _prep(background, "bg", device)
    colors = _prep(colors, "colors_precomp", device)
    language_feature = _prep(language_feature, "language_feature_precomp", device)
    language_feature_instance = _prep(language_feature_instance, "language_feature_instance_precomp", device)
    opacity = _prep(opacity, "opacities", device)
    scales = _prep(scales, "scales", device)
    rotations = _aligned16(_prep(rotations, "rotations", device))
    cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
    all_map = _prep(all_map, "all_map", device)
    viewmatrix = _prep(viewmatrix, "viewmatrix", device)
    projmatrix = _prep(projmatrix, "projmatrix", device)
    sh = _prep(sh, "sh", device)
    campos = _prep(campos, "campos", device)

    M = int(sh.size(1)) if (sh is not None and sh.dim() >= 2 and sh.size(0) != 0) else 0
    F = Fi = 0
    if include_feature:
        if language_feature is None or language_feature.dim() != 2 or language_feature.size(0) != P:
            raise RuntimeError("language_feature_precomp must have dimensions (num_points, F) when include_feature is set")
        if language_feature_instance is None or language_feature_instance.dim() != 2 or language_feature_instance.size(0) != P:
            raise RuntimeError("language_feature_instance_precomp must have dimensions (num_points, Fi) when include_feature is set")
        F, Fi = int(language_feature.size(1)), int(language_feature_instance.size(1))
    n_blend = 3 + F + Fi + (5 if render_geo else 0)
    if n_blend > _lib.MAX_BLEND_CHANNELS:
        raise RuntimeError(f"{n_blend} blended channels exceed the supported maximum of {_lib.MAX_BLEND_CHANNELS}")

    with torch.cuda.device(device):
        fopts = dict(dtype=_FLOAT, device=device)
        iopts = dict(dtype=torch.int32, device=device)
        out_color = torch.empty((3, H, W), **fopts)
        if include_feature:
            out_lang = torch.empty((F, H, W), **fopts)
            out_inst = torch.empty((Fi, H, W), **fopts)
        else:
            out_lang = torch.zeros((1,), **fopts)
            out_inst = torch.zeros((1,), **fopts)
        radii = torch.empty((P,), **iopts)
        out_observe = torch.empty((P,), **iopts)
        out_all_map = torch.empty((5, H, W), **fopts)
        out_plane_depth = torch.empty((1, H, W), **fopts)

        geom, binning, image = _ScratchAlloc(device), _ScratchAlloc(device), _ScratchAlloc(device)
        a = _lib.ForwardArgs()
        a.P, a.D, a.M, a.W, a.H, a.F, a.Fi = P, int(degree), M, W, H, F, Fi
        a.tanfovx, a.tanfovy, a.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
        a.prefiltered, a.render_geo, a.debug, a.include_feature = (
            int(bool(prefiltered)), int(bool(render_geo)), int(bool(debug)), int(bool(include_feature)))
        a.background, a.means3D, a.shs, a.colors_precomp = _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors)
        a.language_feature, a.language_feature_instance = _ptr(language_feature), _ptr(language_feature_instance)
        a.opacities, a.scales, a.rotations = _ptr(opacity), _ptr(scales), _ptr(rotations)
        a.cov3D_precomp, a.all_map = _ptr(cov3D_precomp), _ptr(all_map)
        a.viewmatrix, a.projmatrix, a.campos = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
        a.out_color = out_color.data_ptr()
        a.out_language_feature = out_lang.data_ptr() if include_feature else None
        a.out_language_feature_instance = out_inst.data_ptr() if include_feature else None
        a.radii, a.out_observe = _ptr(radii), _ptr(out_observe)
        a.out_all_map, a.out_plane_depth = out_all_map.data_ptr(), out_plane_depth.data_ptr()
        a.geom_alloc, a.binning_alloc, a.image_alloc = geom.fn, binning.fn, image.fn
        a.stream = _stream_handle(device)
        if raw_params:
            if (all_map is not None and all_map.numel()) or (cov3D_precomp is not None and cov3D_precomp.numel()):
                raise RuntimeError("raw_params=True takes neither all_map nor cov3D_precomp")
            a.raw_params = 1
            pose = _prep(pose, "pose", device)
            if pose is not None and pose.numel() not in (0, 7):
                raise RuntimeError("pose must have 7 elements [quaternion | translation]")
            a.pose = _ptr(pose)
        hint_key = (device.index, P, H, W, n_blend)
        a.binning_capacity_hint = 0 if debug else _capacity_hint(hint_key)
        rendered = ctypes.c_int32(0)
        _check(lib.lsx_rasterize_forward(ctypes.byref(a), ctypes.byref(rendered)), "rasterize_gaussians", geom, binning, image)

        _remember_rendered(hint_key, rendered.value)

    return (int(rendered.value), out_color, out_lang, out_inst, radii, out_observe, out_all_map, out_plane_depth,
            geom.tensor, binning.tensor, image.tensor)


def _grad_arena(device, shapes):
    """One allocation for all gradient tensors (views, 256-B aligned) — also the layout the multi-view
    all-reduce wants."""
    offs, total = [], 0
    for shp in shapes:
        n = 1
        for s in shp:
            n *= int(s)
        offs.append((total, n))
        total += (n + 63) // 64 * 64
    arena = torch.empty((max(total, 1),), dtype=_FLOAT, device=device)
    return arena, [arena[o:o + n].view(shp) for (o, n), shp in zip(offs, shapes)]


def rasterize_gaussians_backward(
    background, all_map_pixels, means3D, radii, colors, language_feature, language_feature_instance, all_maps,
    scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
    dL_dout_color, dL_dout_language_feature, dL_dout_language_feature_instance, dL_dout_all_map, dL_dout_plane_depth,
    sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, render_geo, debug, include_feature,
    *, grad_buffers=None, accumulate=False, raw_params=False, pose=None, pose_grad=None, accumulate_pose=False,
):
    """Positional signature of the reference's `_C.rasterize_gaussians_backward`.  Keyword-only extensions for
    multi-view optimisation: `grad_buffers` maps any of {"means3D", "sh", "opacity", "scales", "rotations", "colors",
    "language_feature", "instance_feature", "all_map", "cov3D"} to a caller-owned contiguous fp32 tensor (e.g. a view
    into lsx_b200.multiview.GradArena) that receives that gradient instead of a fresh tensor; with `accumulate=True`
    the gradients of exactly those groups are ADDED to the buffers (the kernel does the read-modify-write, no extra pass);
    all other outputs are fresh tensors holding this view's gradient.
    Fused render wrapper (`raw_params=True`, same `pose` as the forward call): `means3D / scales / rotations` are the RAW
    parameters and the returned (or sunk) means3D / scales / rotations / opacity gradients are those of the raw parameters;
    the cov3D and all_map entries of the tuple are not written (empty); `pose_grad` (7,) receives the pose gradient (added to
    when `accumulate_pose`)."""
    lib = _lib.load()
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (this operator has no CPU path)")
    device = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    sh = _prep(sh, "sh", device)
    M = int(sh.size(1)) if (sh is not None and sh.dim() >= 2 and sh.size(0) != 0) else 0
    language_feature = _prep(language_feature, "language_feature_precomp", device)
    language_feature_instance = _prep(language_feature_instance, "language_feature_instance_precomp", device)
    F = Fi = 0
    if include_feature:
        F, Fi = int(language_feature.size(1)), int(language_feature_instance.size(1))

    means3D = _prep(means3D, "means3D", device)
    background = _prep(background, "bg", device)
    all_map_pixels = _prep(all_map_pixels, "all_map_pixels", device)
    colors = _prep(colors, "colors_precomp", device)
    all_maps = _prep(all_maps, "all_map", device)
    scales = _prep(scales, "scales", device)
    rotations = _aligned16(_prep(rotations, "rotations", device))
    cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
    viewmatrix = _prep(viewmatrix, "viewmatrix", device)
    projmatrix = _prep(projmatrix, "projmatrix", device)
    campos = _prep(campos, "campos", device)
    dL_dout_color = _prep(dL_dout_color, "dL_dout_color", device)
    dL_dout_language_feature = _prep(dL_dout_language_feature, "dL_dout_language_feature", device)
    dL_dout_language_feature_instance = _prep(dL_dout_language_feature_instance, "dL_dout_language_feature_instance", device)
    dL_dout_all_map = _prep(dL_dout_all_map, "dL_dout_all_map", device)
    dL_dout_plane_depth = _prep(dL_dout_plane_depth, "dL_dout_plane_depth", device)
    radii = radii.contiguous()

    with torch.cuda.device(device):
        shapes = [
            (P, 3), (P, 3), (P, 3),                      # means3D, means2D, means2D_abs
            (P, 3),                                      # colors
            (P, F) if include_feature else (1,),         # language feature
            (P, Fi) if include_feature else (1,),        # instance feature
            (P, 5), (P, 2, 2), (P, 1), (P, 6), (P, M, 3), (P, 3), (P, 4),
        ]
        _, (g_means3D, g_means2D, g_means2D_abs, g_colors, g_lang, g_inst, g_all_map, g_conic, g_opacity, g_cov3D,
            g_sh, g_scales, g_rot) = _grad_arena(device, shapes)
        if not include_feature:
            g_lang.zero_()
            g_inst.zero_()
        if grad_buffers:
            own = {"means3D": g_means3D, "sh": g_sh, "opacity": g_opacity, "scales": g_scales, "rotations": g_rot,
                   "colors": g_colors, "language_feature": g_lang, "instance_feature": g_inst, "all_map": g_all_map,
                   "cov3D": g_cov3D}
            for name, buf in grad_buffers.items():
                ref = own[name]
                if buf.numel() != ref.numel() or buf.dtype != _FLOAT or not buf.is_contiguous() or buf.device != device:
                    raise RuntimeError(f"grad_buffers[{name!r}] must be a contiguous float32 tensor of {ref.numel()} elements on {device}")
                own[name] = buf.view(ref.shape)
            (g_means3D, g_sh, g_opacity, g_scales, g_rot, g_colors, g_lang, g_inst, g_all_map, g_cov3D) = (
                own[k] for k in ("means3D", "sh", "opacity", "scales", "rotations", "colors", "language_feature",
                                 "instance_feature", "all_map", "cov3D"))
        acc_mask = 0
        if accumulate:
            if not grad_buffers:
                raise RuntimeError("accumulate=True needs caller-owned grad_buffers to accumulate into")
            for name in grad_buffers:       # only the caller's buffers are added to; library-owned outputs are overwritten
                acc_mask |= _lib.ACC_BITS[name]

        if P != 0:
            a = _lib.BackwardArgs()
            a.P, a.D, a.M, a.W, a.H, a.F, a.Fi, a.R = P, int(degree), M, W, H, F, Fi, int(R)
            a.tanfovx, a.tanfovy, a.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
            a.render_geo, a.debug, a.include_feature = int(bool(render_geo)), int(bool(debug)), int(bool(include_feature))
            a.background, a.means3D, a.shs, a.colors_precomp = _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors)
            a.language_feature, a.language_feature_instance = _ptr(language_feature), _ptr(language_feature_instance)
            a.all_map, a.scales, a.rotations, a.cov3D_precomp = _ptr(all_maps), _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp)
            a.viewmatrix, a.projmatrix, a.campos, a.radii = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), _ptr(radii)
            a.out_all_map = _ptr(all_map_pixels)
            a.geom_buffer, a.binning_buffer, a.image_buffer = _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer)
            a.dL_dout_color = _ptr(dL_dout_color)
            a.dL_dout_language_feature = _ptr(dL_dout_language_feature) if include_feature else None
            a.dL_dout_language_feature_instance = _ptr(dL_dout_language_feature_instance) if include_feature else None
            a.dL_dout_all_map, a.dL_dout_plane_depth = _ptr(dL_dout_all_map), _ptr(dL_dout_plane_depth)
            a.dL_dmeans2D, a.dL_dmeans2D_abs, a.dL_dconic = g_means2D.data_ptr(), g_means2D_abs.data_ptr(), g_conic.data_ptr()
            a.dL_dopacity, a.dL_dcolors = g_opacity.data_ptr(), g_colors.data_ptr()
            a.dL_dlanguage_feature = g_lang.data_ptr() if include_feature else None
            a.dL_dlanguage_feature_instance = g_inst.data_ptr() if include_feature else None
            a.dL_dmeans3D, a.dL_dcov3D = g_means3D.data_ptr(), g_cov3D.data_ptr()
            a.dL_dsh = g_sh.data_ptr() if M > 0 else None
            a.dL_dscales, a.dL_drotations, a.dL_dall_map = g_scales.data_ptr(), g_rot.data_ptr(), g_all_map.data_ptr()
            a.stream = _stream_handle(device)
            a.accumulate_param_grads = acc_mask
            a.binning_bytes = int(binningBuffer.numel()) if binningBuffer is not None else 0
            if raw_params:
                a.raw_params = 1
                pose = _prep(pose, "pose", device)
                a.pose = _ptr(pose)
                a.dL_dcov3D, a.dL_dall_map = None, None
                if pose_grad is not None:
                    if pose_grad.numel() != 7 or pose_grad.dtype != _FLOAT or not pose_grad.is_contiguous():
                        raise RuntimeError("pose_grad must be a contiguous float32 tensor of 7 elements")
                    a.dL_dpose = pose_grad.data_ptr()
                    a.accumulate_pose = int(bool(accumulate_pose))
            _lib.check(lib.lsx_rasterize_backward(ctypes.byref(a)), "rasterize_gaussians_backward")

    return (g_means2D, g_means2D_abs, g_colors, g_lang, g_inst, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot,
            g_all_map)


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = _lib.load()
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (this operator has no CPU path)")
    device = means3D.device
    P = int(means3D.size(0))
    means3D = _prep(means3D, "means3D", device)
    viewmatrix = _prep(viewmatrix, "viewmatrix", device)
    projmatrix = _prep(projmatrix, "projmatrix", device)
    with torch.cuda.device(device):
        present = torch.empty((P,), dtype=torch.bool, device=device)
        if P != 0:
            _lib.check(lib.lsx_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr(),
                                            _stream_handle(device)), "mark_visible")
    return present


def distCUDA2(points):
    lib = _lib.load()
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor (this operator has no CPU path)")
    device = points.device
    P = int(points.size(0))
    points = _prep(points, "points", device)
    with torch.cuda.device(device):
        means = torch.empty((P,), dtype=_FLOAT, device=device)
        if P != 0:
            scratch = _ScratchAlloc(device)
            _check(lib.lsx_knn_mean_dist2(P, points.data_ptr(), means.data_ptr(), scratch.fn, None,
                                          _stream_handle(device)), "distCUDA2", scratch)
            # `scratch.tensor` may be released now: the caching allocator keeps the block alive for work
            # already enqueued on this stream.
    return means


def render_stats(num_rendered, geomBuffer, binningBuffer, imageBuffer, P, image_height, image_width, n_blend_channels):
    """Workload counters of the view a forward call just rendered, counted on the device (one 64-byte read back):
    dict(S=(pixel, entry) tests per pass of the reference's render loops = sum of n_contrib, B=(pixel, entry) blends,
    V=(8x4 block, entry) visits of this library's backward, Vb=visits in which some pixel blends, L=total length of the
    per-block compacted lists, R=num_rendered).  SURVEY.md 8d: S, B and R accompany every reported number."""
    lib = _lib.load()
    device = geomBuffer.device
    with torch.cuda.device(device):
        out = torch.empty(8, dtype=torch.int64, device=device)
        _lib.check(lib.lsx_render_stats(int(P), int(image_width), int(image_height), int(num_rendered), int(n_blend_channels),
                                        _ptr(geomBuffer), _ptr(binningBuffer), int(binningBuffer.numel()), _ptr(imageBuffer),
                                        out.data_ptr(), _stream_handle(device)), "render_stats")
        v = out.cpu().tolist()
    return {"S": v[0], "B": v[1], "V": v[2], "Vb": v[3], "L": v[4], "R": int(num_rendered), "Hmax": v[5], "Hsum": v[6]}
