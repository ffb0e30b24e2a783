"""View-sharded optimisation step of LangScene-X's field construction on the fused kernels (SURVEY.md 8e + 8f; BASELINE
configs 4 and 5) — the caller side of the rasterizer hot path.

One step restates, for a BATCH of views sharded over the ranks of one node, what one iteration of the reference's loop
(field_construction/gaussian_field.py:184-543) does for one view:

    pose_transform (optim_pose, gaussian_renderer/__init__.py:79-87)          } inside lsx_rasterize_forward's per-Gaussian kernel
    activations + plane normal + all_map (gaussian_model.py:193-236,          } (raw_params; LoopConfig.fused_wrapper=False: the
        gaussian_renderer/__init__.py:188-196)                                   stand-alone lsx_pose_transform / lsx_gaussian_head)
    rasterizer forward                                                         lsx_rasterize_forward
    depth -> normal * alpha.detach() (gaussian_renderer/__init__.py:233-235)   lsx_depth_normal_forward
    (1 - l) L1 + l (1 - SSIM)                 (gaussian_field.py:238-246)      lsx_image_loss_forward
    w * mean(iw * sum_c |depth_normal - normal|)  (gaussian_field.py:262-285)  lsx_masked_l1_forward  (3 * mean over c, h, w;
                                                                               iw = the per-view image-gradient weight map)
    l1(language * mask, gt * mask)            (gaussian_field.py:447-451)      lsx_masked_l1_forward
    loss_cls_3d(xyz, language_feature)        (gaussian_field.py:461-465)      lsx_cls3d_forward
    backward of all of the above, rasterizer backward, head / pose backward    — the per-Gaussian parameter gradients are
        written or ADDED straight into ONE flat gradient arena by the kernels themselves (no autograd graph over the
        P-sized tensors, no accumulation passes)
    add_densification_stats + max_radii2D     (gaussian_field.py:519-524)      lsx_densify_stats_update (per-step delta)
then, once per step
    all-reduce of the arena over the ranks (NCCL), issued group by group as soon as the last local view has written a group
        so that the transfers overlap the rest of that view's backward, + the (V, 7) pose-gradient rows + the statistics delta
    Adam on the flat parameter arena          (gaussian_field.py:537-543)      lsx_arena_adam_step

Every rank holds a replica of the parameters; equal seeds give equal replicas and the all-reduced gradients are bitwise
identical on every rank, so the replicas stay identical without any parameter broadcast.  No CPU path.
"""
import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, ops
from .densify import add_densification_stats
from .loss import _Cls3d, _ImageLoss, _MaskedL1, knn_tree
from .multiview import DensifyStats, GradArena, PendingReduce
from .optim import ArenaAdam
from .render_utils import _DepthToNormal, _PoseTransform


class _Tape:
    """Stand-in for an autograd context: the fused operators' forward / backward staticmethods are called directly, so the
    image-space chain costs no autograd graph (at LangScene-X's own 720x480 the graph bookkeeping is a visible fraction
    of an iteration)."""
    needs_input_grad = (True,) * 16

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors

    def mark_non_differentiable(self, *tensors):
        pass


@dataclass
class LoopConfig:
    sh_degree: int = 3
    lambda_dssim: float = 0.2            # configs/field_construction.yaml:84
    normal_weight: float = 0.10          # single_view_weight (:95)
    language_weight: float = 1.0
    cls3d: bool = True                   # opt.loss_obj_3d
    reg3d_k: int = 5                     # :119
    reg3d_lambda: float = 4.0            # :120
    reg3d_samples: int = 800
    # neighbour search of loss_cls_3d through one Morton / box hierarchy per step (lsx_knn_tree_build, 0.14 ms at 500 k) and one
    # CTA per query (knn.cu) instead of a scan of all points per view: the same neighbours bit for bit (tests/test_cls3d.py);
    # 0.28 vs 0.50 ms per loss call at 500 k, 0.45 vs 1.64 ms at 2 M; a C4-loop view 1.61 vs 1.78 ms (profiles/r6o_*)
    cls3d_tree: bool = True
    optimise_pose: bool = True           # optim_pose (:67)
    densify_stats: bool = True
    # per-group asynchronous all-reduces issued while the last view's backward is still running, instead of ONE blocking
    # collective after it.  Measured at C4 on 2 GPUs (profiles/r6e_c4loop_n2*.json): 11.47 vs 11.08 ms per step — five small
    # collectives and NCCL's CTAs competing with the wrapper's backward cost more than the 0.28 ms they hide; off by default.
    overlap_allreduce: bool = False
    # the render wrapper's per-Gaussian work (pose transform, activations, plane normal, all_map) and its backward inside the
    # rasterizer's per-Gaussian kernels (lsx_forward_args.raw_params) instead of in separate passes over P
    fused_wrapper: bool = True


@dataclass
class View:
    """One training view, device-resident.  `index` = row of the pose table / global view id."""
    index: int
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor      # (4,4) world_view_transform in the reference's memory order
    projmatrix: torch.Tensor      # (4,4) full_proj_transform
    campos: torch.Tensor          # (3,)
    gt_image: torch.Tensor        # (3,H,W)
    gt_language: Optional[torch.Tensor] = None   # (F,H,W)
    language_mask: Optional[torch.Tensor] = None  # (H,W) float/bool or None
    image_weight: Optional[torch.Tensor] = None   # (H,W) >= 0: (1 - get_img_grad_weight(gt)).clamp(0,1)^2 of the single-view
                                                  # normal loss (gaussian_field.py:262-263); None = the wo_image_weight form

    @property
    def fx(self):
        return self.W / (2.0 * self.tanfovx)

    @property
    def fy(self):
        return self.H / (2.0 * self.tanfovy)


# arena spans in the order the last local view finalises them (each span = adjacent groups, one collective):
# the rasterizer's backward completes sh and the two feature groups (86 % of the bytes at F = 3, M = 16), the wrapper's
# backward the rest.  "pose" rows exist only under pose optimisation.
_SPANS_AFTER_RASTER = (("sh",), ("language_feature", "instance_feature"))
_SPANS_AFTER_HEAD = (("means3D",), ("opacity", "scales", "rotations"))


class FieldLoop:
    """Parameters, gradient arena, Adam state and densification statistics of one replica + the sharded step."""

    def __init__(self, raw: Dict[str, torch.Tensor], lrs: Dict[str, float], background: torch.Tensor, cfg: LoopConfig = None,
                 n_views: int = 0, poses: Optional[torch.Tensor] = None, group=None):
        """raw: means3D (P,3), sh (P,M,3) or (P,3M), opacity (P,1) logits, scales (P,3) log-scales, rotations (P,4)
        un-normalised quaternions, language_feature (P,F), instance_feature (P,Fi) — the reference's nn.Parameters.
        poses: (n_views, 7) [quaternion | translation] rows of GaussianModel.P (identity if None)."""
        self.cfg = cfg or LoopConfig()
        dev = raw["means3D"].device
        if dev.type != "cuda":
            raise RuntimeError("FieldLoop needs CUDA tensors (this path has no CPU fallback)")
        self.device, self.group = dev, group
        P = int(raw["means3D"].shape[0])
        sh = raw["sh"].reshape(P, -1)
        self.P, self.M = P, sh.shape[1] // 3
        self.F, self.Fi = int(raw["language_feature"].shape[1]), int(raw["instance_feature"].shape[1])
        extra = {"pose": (int(n_views), 7)} if self.cfg.optimise_pose else None
        if self.cfg.optimise_pose and n_views <= 0:
            raise ValueError("pose optimisation needs the number of views (rows of the pose table)")
        self.params = GradArena.allocate(P, self.M, self.F, self.Fi, dev, extra=extra)
        self.grads = GradArena.allocate(P, self.M, self.F, self.Fi, dev, extra=extra)
        for name in ("means3D", "opacity", "scales", "rotations", "language_feature", "instance_feature"):
            self.params.views[name].copy_(raw[name].reshape(self.params.views[name].shape))
        self.params.views["sh"].copy_(sh)
        if self.cfg.optimise_pose:
            if poses is None:
                poses = torch.zeros(n_views, 7, device=dev)
                poses[:, 0] = 1.0
            self.params.views["pose"].copy_(poses)
        self.opt = ArenaAdam(self.params, lrs)
        self.stats = DensifyStats.allocate(P, dev)
        self._delta = DensifyStats.allocate(P, dev)
        self.bg = background.to(dev, torch.float32).contiguous()
        lam = self.cfg.lambda_dssim
        self._g_ssim = torch.tensor([-lam], device=dev)
        self._g_l1 = torch.tensor([1.0 - lam], device=dev)
        self._g_normal = torch.tensor([3.0 * self.cfg.normal_weight], device=dev)   # sum over 3 channels = 3 * mean
        self._g_lang = torch.tensor([self.cfg.language_weight], device=dev)
        self._g_one = torch.tensor([1.0], device=dev)
        self._empty = torch.Tensor([])
        self._zero_img: Dict[tuple, torch.Tensor] = {}
        self.debug_tap = None
        self._tree = None
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    # ---- helpers ----------------------------------------------------------------------------------------------------
    def _zeros(self, *shape):
        t = self._zero_img.get(shape)
        if t is None:
            t = self._zero_img[shape] = torch.zeros(shape, device=self.device)
        return t

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _head_forward(self, xyz, rot_in, view: View):
        pv = self.params.views
        P, dev = self.P, self.device
        opts = dict(dtype=torch.float32, device=dev)
        scales, rotations = torch.empty((P, 3), **opts), torch.empty((P, 4), **opts)
        opacity, all_map = torch.empty((P, 1), **opts), torch.empty((P, 5), **opts)
        _lib.check(_lib.load().lsx_gaussian_head_forward(P, view.viewmatrix.data_ptr(), view.campos.data_ptr(), xyz.data_ptr(),
                                                         pv["scales"].data_ptr(), rot_in.data_ptr(), pv["opacity"].data_ptr(),
                                                         scales.data_ptr(), rotations.data_ptr(), opacity.data_ptr(),
                                                         all_map.data_ptr(), self._stream()), "gaussian_head")
        return scales, rotations, opacity, all_map

    # ---- one view: forward, losses, backward into the arena ----------------------------------------------------------
    def _view(self, view: View, first: bool, last: bool, sample_idx: Optional[torch.Tensor], pending: PendingReduce,
              stats: Optional[DensifyStats]):
        cfg, pv, gv, dev = self.cfg, self.params.views, self.grads.views, self.device
        P, H, W = self.P, view.H, view.W
        acc = not first
        losses = {}

        # -- forward --
        e = self._empty
        pose_row = pv["pose"][view.index] if cfg.optimise_pose else None
        xyz, rot_in, pose_tape = pv["means3D"], pv["rotations"], None
        if cfg.fused_wrapper:
            fwd = ops.rasterize_gaussians(self.bg, pv["means3D"], e, pv["language_feature"], pv["instance_feature"], pv["opacity"],
                                          pv["scales"], pv["rotations"], 1.0, e, e, view.viewmatrix, view.projmatrix, view.tanfovx,
                                          view.tanfovy, H, W, pv["sh"].view(P, self.M, 3), cfg.sh_degree, view.campos, False,
                                          True, False, True, raw_params=True, pose=pose_row)
        else:
            if cfg.optimise_pose:
                pose_tape = _Tape()
                xyz, rot_in = _PoseTransform.forward(pose_tape, pose_row, pv["means3D"], pv["rotations"])
            scales, rotations, opacity, all_map = self._head_forward(xyz, rot_in, view)
            fwd = ops.rasterize_gaussians(self.bg, xyz, e, pv["language_feature"], pv["instance_feature"], opacity, scales,
                                          rotations, 1.0, e, all_map, view.viewmatrix, view.projmatrix, view.tanfovx,
                                          view.tanfovy, H, W, pv["sh"].view(P, self.M, 3), cfg.sh_degree, view.campos, False,
                                          True, False, True)
        (R, color, lang, inst, radii, observe, amap, depth, geom, binning, img) = fwd

        # -- image-space losses: forward + their backward, called directly (no autograd graph) --
        t_img = _Tape()
        ssim_mean, l1_mean = _ImageLoss.forward(t_img, color, view.gt_image)
        (g_color, _) = _ImageLoss.backward(t_img, self._g_ssim, self._g_l1)
        losses["l1"], losses["ssim"] = l1_mean, ssim_mean

        g_amap = self._zeros(5, H, W)
        g_depth = self._zeros(1, H, W)
        if cfg.normal_weight > 0.0:
            t_dn = _Tape()
            dn = _DepthToNormal.forward(t_dn, depth[0], amap[3], view.fx, view.fy, W * 0.5, H * 0.5)
            t_nl = _Tape()
            # mean over (3, H, W) of w |dn - normal|  (w >= 0, so |w a - w b| = w |a - b|)
            normal_l1 = _MaskedL1.forward(t_nl, dn, amap[:3], view.image_weight)
            (g_dn, _, _) = _MaskedL1.backward(t_nl, self._g_normal)
            g_amap = torch.empty((5, H, W), device=dev)
            torch.neg(g_dn, out=g_amap[:3])                                 # d|a - b|/db = -d|a - b|/da
            g_amap[3:].zero_()
            g_depth = _DepthToNormal.backward(t_dn, g_dn)[0].view(1, H, W)
            losses["normal"] = normal_l1

        if view.gt_language is not None and cfg.language_weight > 0.0:
            t_ll = _Tape()
            lang_l1 = _MaskedL1.forward(t_ll, lang, view.gt_language, view.language_mask)
            (g_lang, _, _) = _MaskedL1.backward(t_ll, self._g_lang)
            losses["language"] = lang_l1
        else:
            g_lang = self._zeros(self.F, H, W)
        g_inst = self._zeros(self.Fi, H, W)

        if cfg.fused_wrapper:
            # -- rasterizer + wrapper backward in one: EVERY parameter gradient (and the pose row) goes straight into the arena --
            sink = {k: gv[k] for k in ("means3D", "sh", "opacity", "scales", "rotations", "language_feature", "instance_feature")}
            bwd = ops.rasterize_gaussians_backward(self.bg, amap, pv["means3D"], radii, e, pv["language_feature"],
                                                   pv["instance_feature"], e, pv["scales"], pv["rotations"], 1.0, e,
                                                   view.viewmatrix, view.projmatrix, view.tanfovx, view.tanfovy, g_color, g_lang,
                                                   g_inst, g_amap, g_depth, pv["sh"].view(P, self.M, 3), cfg.sh_degree, view.campos,
                                                   geom, R, binning, img, True, False, True, grad_buffers=sink, accumulate=acc,
                                                   raw_params=True, pose=pose_row,
                                                   pose_grad=gv["pose"][view.index] if cfg.optimise_pose else None,
                                                   accumulate_pose=True)
            g_m2d, g_m2d_abs = bwd[0], bwd[1]
            if self.debug_tap is not None:
                self.debug_tap.append({k: v.detach().clone() for k, v in dict(
                    color=color, lang=lang, amap=amap, depth=depth, g_color=g_color, g_lang=g_lang, g_amap=g_amap, g_depth=g_depth,
                    g_m2d=g_m2d).items()})
            if cfg.cls3d and sample_idx is not None:
                t_c = _Tape()
                cls_loss, _nbr = _Cls3d.forward(t_c, pv["means3D"], pv["language_feature"], sample_idx, cfg.reg3d_k, cfg.reg3d_lambda,
                                                self._tree)
                gv["language_feature"].add_(_Cls3d.backward(t_c, self._g_one, None)[1])
                losses["cls3d"] = cls_loss
            if last and self.world > 1 and cfg.overlap_allreduce:
                spans = list(_SPANS_AFTER_RASTER) + list(_SPANS_AFTER_HEAD) + ([("pose",)] if cfg.optimise_pose else [])
                pending.extend(self.grads.all_reduce_spans(spans, self.group))
            if stats is not None:
                add_densification_stats(stats, g_m2d, g_m2d_abs, radii, observe)
            return losses

        # -- rasterizer backward: sh / language / instance gradients go straight into the arena --
        sink = {"sh": gv["sh"], "language_feature": gv["language_feature"], "instance_feature": gv["instance_feature"]}
        bwd = ops.rasterize_gaussians_backward(self.bg, amap, xyz, radii, e, pv["language_feature"], pv["instance_feature"],
                                               all_map, scales, rotations, 1.0, e, view.viewmatrix, view.projmatrix,
                                               view.tanfovx, view.tanfovy, g_color, g_lang, g_inst, g_amap, g_depth,
                                               pv["sh"].view(P, self.M, 3), cfg.sh_degree, view.campos, geom, R, binning, img,
                                               True, False, True, grad_buffers=sink, accumulate=acc)
        (g_m2d, g_m2d_abs, _gc, _gl, _gi, g_opac, g_m3d, _gcov, _gsh, g_scales, g_rot, g_allmap) = bwd
        if self.debug_tap is not None:      # tests / tools: the intermediates of this view, cloned
            self.debug_tap.append({k: v.detach().clone() for k, v in dict(
                color=color, lang=lang, amap=amap, depth=depth, g_color=g_color, g_lang=g_lang, g_amap=g_amap, g_depth=g_depth,
                g_m3d=g_m3d, g_scales=g_scales, g_rot=g_rot, g_opac=g_opac, g_allmap=g_allmap, g_m2d=g_m2d).items()})

        # -- 3-D neighbourhood regulariser on the language feature parameter (adds into the arena) --
        if cfg.cls3d and sample_idx is not None:
            t_c = _Tape()
            cls_loss, _nbr = _Cls3d.forward(t_c, pv["means3D"], pv["language_feature"], sample_idx, cfg.reg3d_k, cfg.reg3d_lambda,
                                                self._tree)
            g_feat = _Cls3d.backward(t_c, self._g_one, None)[1]
            gv["language_feature"].add_(g_feat)
            losses["cls3d"] = cls_loss
        if last and self.world > 1 and cfg.overlap_allreduce:
            pending.extend(self.grads.all_reduce_spans(_SPANS_AFTER_RASTER, self.group))

        # -- wrapper backward: raw scaling / rotation / opacity / position gradients into the arena --
        lib = _lib.load()
        if not cfg.optimise_pose:
            _lib.check(lib.lsx_gaussian_head_backward_acc(
                P, view.viewmatrix.data_ptr(), view.campos.data_ptr(), xyz.data_ptr(), pv["scales"].data_ptr(),
                rot_in.data_ptr(), pv["opacity"].data_ptr(), g_scales.data_ptr(), g_rot.data_ptr(), g_opac.data_ptr(),
                g_allmap.data_ptr(), g_m3d.data_ptr(), gv["means3D"].data_ptr(), gv["scales"].data_ptr(),
                gv["rotations"].data_ptr(), gv["opacity"].data_ptr(), 0xF if acc else 0, self._stream()), "gaussian_head backward")
        else:
            d_xyz_w, d_rot_w = torch.empty_like(xyz), torch.empty_like(rot_in)
            _lib.check(lib.lsx_gaussian_head_backward_acc(
                P, view.viewmatrix.data_ptr(), view.campos.data_ptr(), xyz.data_ptr(), pv["scales"].data_ptr(),
                rot_in.data_ptr(), pv["opacity"].data_ptr(), g_scales.data_ptr(), g_rot.data_ptr(), g_opac.data_ptr(),
                g_allmap.data_ptr(), g_m3d.data_ptr(), d_xyz_w.data_ptr(), gv["scales"].data_ptr(), d_rot_w.data_ptr(),
                gv["opacity"].data_ptr(), 0xA if acc else 0, self._stream()), "gaussian_head backward")
            d_pose, d_xyz, d_rot = _PoseTransform.backward(pose_tape, d_xyz_w, d_rot_w)
            if acc:
                gv["means3D"].add_(d_xyz)
                gv["rotations"].add_(d_rot)
            else:
                gv["means3D"].copy_(d_xyz)
                gv["rotations"].copy_(d_rot)
            gv["pose"][view.index].add_(d_pose.reshape(7))                 # rows are zeroed at the start of the step
        if last and self.world > 1 and cfg.overlap_allreduce:
            spans = list(_SPANS_AFTER_HEAD) + ([("pose",)] if cfg.optimise_pose else [])
            pending.extend(self.grads.all_reduce_spans(spans, self.group))

        if stats is not None:
            add_densification_stats(stats, g_m2d, g_m2d_abs, radii, observe)
        return losses

    # ---- one optimisation step over this rank's views -----------------------------------------------------------------
    def gradient(self, views: Sequence[View], sample_idx: Optional[Sequence[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Forward + backward of the local views, all-reduce; leaves the summed gradient in self.grads (identical on every
        rank) and returns the per-loss sums over the LOCAL views (device scalars)."""
        if not views:
            raise ValueError("every rank needs at least one view per step")
        with torch.cuda.device(self.device):
            pending = PendingReduce()
            delta = self._delta.zero_() if self.cfg.densify_stats else None
            if self.cfg.optimise_pose:
                self.grads.views["pose"].zero_()
            totals: Dict[str, List[torch.Tensor]] = {}
            # loss_cls_3d searches neighbours among the positions, which do not change inside a step: one search structure
            # for all of the step's views (a full scan of the points per view was the largest kernel of a C4-sized iteration)
            self._tree = knn_tree(self.params.views["means3D"]) if (self.cfg.cls3d and self.cfg.cls3d_tree and sample_idx is not None) else None
            for i, vw in enumerate(views):
                si = sample_idx[i] if sample_idx is not None else None
                for k, v in self._view(vw, i == 0, i == len(views) - 1, si, pending, delta).items():
                    totals.setdefault(k, []).append(v)
            if self.world > 1:
                if not self.cfg.overlap_allreduce:
                    self.grads.all_reduce(self.group)
                if delta is not None:
                    delta.all_reduce(self.group)
            if delta is not None:
                delta.merge_into(self.stats)
            pending.wait()
            return {k: torch.stack(v).sum() for k, v in totals.items()}

    # ---- densification on the replica's arenas (gaussian_field.py:526-535) ------------------------------------------------
    def _gaussian_groups(self, flat: torch.Tensor) -> GradArena:
        """An arena object over the per-Gaussian groups of `flat` (same memory; the pose rows at the tail are left out)."""
        offs = {n: oc for n, oc in self.params.offsets.items() if n != "pose"}
        views = {n: flat[o:o + c].view(self.params.views[n].shape) for n, (o, c) in offs.items()}
        return GradArena(flat, views, offs)

    def densify_and_prune(self, max_grad: float, abs_max_grad: float, min_opacity: float, extent: float, max_screen_size=None,
                          dcfg=None, generator=None):
        """GaussianModel.densify_and_prune (gaussian_model.py:700-718) on this replica: plan from the PERSISTENT statistics
        (identical on every rank: only all-reduced deltas were ever merged into them), one gather per arena for parameters and
        both Adam moments, statistics restarted.  With equally seeded `generator`s every rank samples the same new Gaussians,
        so the replicas stay bit-identical without any broadcast.  Returns lsx_b200.densify.DensifyResult."""
        from .densify import DensifyConfig, densify_and_prune
        with torch.cuda.device(self.device):
            res = densify_and_prune(self._gaussian_groups(self.params.flat), self._gaussian_groups(self.opt.exp_avg),
                                    self._gaussian_groups(self.opt.exp_avg_sq), self.stats, dcfg or DensifyConfig(), max_grad,
                                    abs_max_grad, min_opacity, extent, max_screen_size, generator=generator)
            P_new = next(iter(res.params.views.values())).shape[0]
            extra = {"pose": tuple(self.params.views["pose"].shape)} if self.cfg.optimise_pose else None
            new = [GradArena.allocate(P_new, self.M, self.F, self.Fi, self.device, extra=extra) for _ in range(3)]
            for dst, src_new, src_old in zip(new, (res.params, res.exp_avg, res.exp_avg_sq),
                                             (self.params.flat, self.opt.exp_avg, self.opt.exp_avg_sq)):
                g = src_new.flat.numel() if P_new > 0 else 0
                dst.flat[:g].copy_(src_new.flat[:g])                       # identical group order and 256-B alignment
                if self.cfg.optimise_pose:
                    o_new, c = dst.offsets["pose"]
                    o_old, _ = self.params.offsets["pose"]
                    dst.flat[o_new:o_new + c].copy_(src_old[o_old:o_old + c])
            self.params = new[0]
            self.grads = GradArena.allocate(P_new, self.M, self.F, self.Fi, self.device, extra=extra)
            self.opt.rebind(self.params, new[1].flat, new[2].flat)
            self.stats, self._delta, self.P = res.stats, DensifyStats.allocate(P_new, self.device), P_new
        return res

    def reset_opacity(self):
        """GaussianModel.reset_opacity (gaussian_model.py:443-446) in place, Adam moments of the opacity group zeroed."""
        from .densify import reset_opacity
        reset_opacity(self.params, self._gaussian_groups(self.opt.exp_avg), self._gaussian_groups(self.opt.exp_avg_sq))

    def step(self, views: Sequence[View], sample_idx: Optional[Sequence[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        losses = self.gradient(views, sample_idx)
        with torch.cuda.device(self.device):
            self.opt.step(self.grads)
        return losses
