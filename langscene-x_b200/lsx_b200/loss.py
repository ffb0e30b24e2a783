"""Fused image losses of LangScene-X's training loop on the C ABI.

  ssim(img1, img2)                  drop-in for field_construction/utils/loss_utils.py:37-46 (11x11 window, size_average)
  image_loss(image, gt, lambda_)    (1 - lambda_) * l1_loss(image, gt) + lambda_ * (1 - ssim(image, gt)), the combination at
                                    field_construction/gaussian_field.py:238-246, one kernel forward and one backward;
                                    returns (loss, l1, ssim_value) so that the loop can keep logging both terms.
  masked_l1_loss(out, gt, mask)     the language-feature supervision (gaussian_field.py:450-451)
  loss_cls_3d(xyz, feature, ...)    the 3-D neighbourhood regulariser (loss_utils.py:158-186): tiled exact k-NN of the sampled
                                    rows + KL-style term, no (samples x N) distance matrix
Gradients flow to the FIRST image only (the ground-truth image never requires grad in the reference).  No CPU path.
"""
import ctypes

import torch

from . import _lib


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_pair(img1, img2):
    if not (img1.is_cuda and img2.is_cuda):
        raise RuntimeError("images must be CUDA tensors (this operator has no CPU path)")
    if img1.shape != img2.shape or img1.dim() != 3:
        raise RuntimeError("images must both have shape (C, H, W)")
    if img1.dtype != torch.float32 or img2.dtype != torch.float32:
        raise RuntimeError("images must be float32")


class _ImageLoss(torch.autograd.Function):
    """returns (mean ssim, mean |x - y|); backward takes the two upstream scalars."""

    @staticmethod
    def forward(ctx, img1, img2):
        _check_pair(img1, img2)
        if img2.requires_grad:
            raise RuntimeError("the second image is treated as ground truth: it must not require grad")
        img1c, img2c = img1.contiguous(), img2.contiguous()
        C, H, W = img1c.shape
        lib = _lib.load()
        nblk = int(lib.lsx_image_loss_num_blocks(C, H, W))
        dmaps = torch.empty((3, C, H, W), dtype=torch.float32, device=img1.device)
        partial = torch.empty((2, nblk), dtype=torch.float32, device=img1.device)
        with torch.cuda.device(img1.device):
            _lib.check(lib.lsx_image_loss_forward(C, H, W, img1c.data_ptr(), img2c.data_ptr(), dmaps.data_ptr(),
                                                  partial.data_ptr(), _stream(img1.device)), "image_loss")
        sums = partial.sum(dim=1) / float(C * H * W)
        ctx.save_for_backward(img1c, img2c, dmaps)
        return sums[0], sums[1]

    @staticmethod
    def backward(ctx, g_ssim, g_l1):
        img1, img2, dmaps = ctx.saved_tensors
        C, H, W = img1.shape
        dev = img1.device
        scal = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        gs, gl = scal(g_ssim), scal(g_l1)                          # device scalars: no host read in backward
        g = torch.empty_like(img1)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().lsx_image_loss_backward(C, H, W, img1.data_ptr(), img2.data_ptr(), dmaps.data_ptr(),
                                                           None if gs is None else gs.data_ptr(),
                                                           None if gl is None else gl.data_ptr(), g.data_ptr(), _stream(dev)),
                       "image_loss backward")
        return g, None


def ssim(img1, img2, window_size=11, size_average=True):
    if window_size != 11 or not size_average:
        raise NotImplementedError("the fused kernel implements the 11x11, size_average=True case used by the training loop")
    return _ImageLoss.apply(img1, img2)[0]


def image_loss(image, gt, lambda_dssim):
    s, l1 = _ImageLoss.apply(image, gt)
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - s), l1, s


class _MaskedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, mask):
        _check_pair(a, b)
        if b.requires_grad:
            raise RuntimeError("the second tensor is treated as ground truth: it must not require grad")
        ac, bc = a.contiguous(), b.contiguous()
        C, H, W = ac.shape
        mc = 1
        if mask is not None:
            mask = mask.detach().to(device=a.device, dtype=torch.float32)
            if mask.numel() == H * W:
                mask = mask.reshape(H, W).contiguous()
            elif tuple(mask.shape) == (C, H, W):
                mask, mc = mask.contiguous(), C
            else:
                raise RuntimeError("mask must have H*W elements (broadcast over channels) or shape (C, H, W)")
        lib = _lib.load()
        nblk = int(lib.lsx_masked_l1_num_blocks(C * H * W))
        partial = torch.empty(nblk, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            _lib.check(lib.lsx_masked_l1_forward(C, H, W, mc, ac.data_ptr(), bc.data_ptr(),
                                                 None if mask is None else mask.data_ptr(), partial.data_ptr(),
                                                 _stream(a.device)), "masked_l1")
        ctx.save_for_backward(ac, bc, mask if mask is not None else torch.empty(0, device=a.device))
        ctx.mc = mc
        return partial.sum() / float(C * H * W)

    @staticmethod
    def backward(ctx, g):
        a, b, mask = ctx.saved_tensors
        C, H, W = a.shape
        dev = a.device
        up = g.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()     # device scalar: no host read
        out = torch.empty_like(a)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().lsx_masked_l1_backward(C, H, W, ctx.mc, a.data_ptr(), b.data_ptr(),
                                                          mask.data_ptr() if mask.numel() else None, up.data_ptr(),
                                                          out.data_ptr(), _stream(dev)), "masked_l1 backward")
        return out, None, None


def masked_l1_loss(network_output, gt, mask=None):
    """l1_loss(network_output * mask, gt * mask) of the language-feature supervision (field_construction/gaussian_field.py:
    450-451; l1_loss = mean |a - b|, field_construction/utils/loss_utils.py:20-21) in one kernel each way.  `mask`: (H, W) /
    (1, H, W) bool or float (as Camera.get_language_feature returns it, field_construction/scene/cameras.py:137-151), or
    (C, H, W); None = plain l1_loss.  Gradient flows to `network_output` only."""
    return _MaskedL1.apply(network_output, gt, mask)


class _Cls3d(torch.autograd.Function):
    """lambda * mean |q[s] (log(q[s] + eps) - log(q[nbr] + eps))| for given sample rows; gradient to `predictions` only."""

    @staticmethod
    def forward(ctx, features, predictions, sample_idx, k, lambda_val, tree=None):
        N, C = predictions.shape
        S = int(sample_idx.numel())
        dev = predictions.device
        pts, pred = features.detach().contiguous(), predictions.contiguous()
        lib = _lib.load()
        nbytes = int(lib.lsx_cls3d_scratch_bytes(N, C, S, k))
        if nbytes <= 0:
            raise RuntimeError(f"loss_cls_3d: need 1 <= k <= min(N, 8) and non-empty inputs (N={N}, C={C}, samples={S}, k={k})")
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        nbr = torch.empty((S, k), dtype=torch.int32, device=dev)
        minmax = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if tree is not None:
                _lib.check(lib.lsx_cls3d_forward_tree(N, C, S, k, float(lambda_val), pts.data_ptr(), pred.data_ptr(),
                                                      sample_idx.data_ptr(), loss.data_ptr(), nbr.data_ptr(), minmax.data_ptr(),
                                                      scratch.data_ptr(), tree.data_ptr(), _stream(dev)), "loss_cls_3d")
            else:
                _lib.check(lib.lsx_cls3d_forward(N, C, S, k, float(lambda_val), pts.data_ptr(), pred.data_ptr(),
                                                 sample_idx.data_ptr(), loss.data_ptr(), nbr.data_ptr(), minmax.data_ptr(),
                                                 scratch.data_ptr(), _stream(dev)), "loss_cls_3d")
        ctx.save_for_backward(pred, sample_idx, nbr, minmax)
        ctx.k, ctx.lambda_val = k, float(lambda_val)
        ctx.mark_non_differentiable(nbr)
        return loss, nbr

    @staticmethod
    def backward(ctx, g, _g_nbr):
        pred, sample_idx, nbr, minmax = ctx.saved_tensors
        N, C = pred.shape
        S, dev = int(sample_idx.numel()), pred.device
        lib = _lib.load()
        up = g.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()          # device scalar: no host read
        scratch = torch.empty(int(lib.lsx_cls3d_scratch_bytes(N, C, S, ctx.k)), dtype=torch.uint8, device=dev)
        out = torch.empty_like(pred)
        with torch.cuda.device(dev):
            _lib.check(lib.lsx_cls3d_backward(N, C, S, ctx.k, ctx.lambda_val, pred.data_ptr(), sample_idx.data_ptr(),
                                              nbr.data_ptr(), minmax.data_ptr(), up.data_ptr(), out.data_ptr(),
                                              scratch.data_ptr(), _stream(dev)), "loss_cls_3d backward")
        return None, out, None, None, None, None


class KnnTree:
    """Opaque search structure of a point set (see `knn_tree`): the device bytes plus the point count they were built for."""

    def __init__(self, data: torch.Tensor, n_points: int):
        self.data, self.n_points = data, int(n_points)

    def data_ptr(self):
        return self.data.data_ptr()


def knn_tree(points) -> KnnTree:
    """The search structure (Morton order + box hierarchy) of an (N, 3) float32 CUDA point set for `loss_cls_3d(..., tree=)`.
    Valid while `points` keeps its values: build it once per optimisation step and share it between the step's views."""
    if not points.is_cuda or points.dim() != 2 or points.shape[1] != 3 or points.dtype != torch.float32:
        raise RuntimeError("knn_tree: points must be a float32 (N, 3) CUDA tensor (this operator has no CPU path)")
    pts = points.detach().contiguous()
    lib = _lib.load()
    N = int(pts.shape[0])
    tree = torch.empty(int(lib.lsx_knn_tree_bytes(N)), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(lib.lsx_knn_tree_build(N, pts.data_ptr(), tree.data_ptr(), _stream(pts.device)), "knn_tree_build")
    return KnnTree(tree, N)


def loss_cls_3d(features, predictions, k=5, lambda_val=2.0, max_points=200000, sample_size=800, *, sample_indices=None,
                return_neighbors=False, tree=None):
    """Drop-in for loss_cls_3d of field_construction/utils/loss_utils.py:158-186 (the 3-D neighbourhood regulariser of the
    language / instance features; called with features = xyz.detach(), predictions = the (N, C) feature parameter at
    field_construction/gaussian_field.py:461-465,482-485): same positional signature, same random draws — the optional
    down-sampling to `max_points` rows and the `sample_size` query rows both come from `torch.randperm(n)` on the CPU
    generator, in the reference's order, so equal seeds select equal rows — but no (samples x N) distance matrix, no topk over
    it and no host read of min / max: the k nearest neighbours are found by a tiled exact scan and the loss, its gradient
    (including the part through the min / max normalisation) and the neighbour indices stay on the device.
    `sample_indices` (1-D integer tensor) overrides the second draw; 1 <= k <= 8.  Gradient flows to `predictions` only.
    `tree` (from `knn_tree(features)`) replaces the scan of all points by a search through a prebuilt hierarchy: same
    neighbours, same order; reuse it for every call made while the positions are unchanged."""
    if not (features.is_cuda and predictions.is_cuda):
        raise RuntimeError("loss_cls_3d: tensors must be CUDA tensors (this operator has no CPU path)")
    if features.dim() != 2 or features.shape[1] != 3 or features.dtype != torch.float32:
        raise RuntimeError("loss_cls_3d: features must be a float32 (N, 3) tensor of positions")
    if predictions.dim() == 1:
        predictions = predictions.unsqueeze(1)
    if predictions.dim() != 2 or predictions.shape[0] != features.shape[0] or predictions.dtype != torch.float32:
        raise RuntimeError("loss_cls_3d: predictions must be a float32 (N, C) tensor with one row per position")
    if features.size(0) > max_points:                                   # loss_utils.py:160-163
        indices = torch.randperm(features.size(0))[:max_points].to(features.device)
        features = features[indices]
        predictions = predictions[indices]
    n = features.size(0)
    if sample_indices is None:                                          # loss_utils.py:172
        sample_indices = torch.randperm(n)[:sample_size]
    idx = sample_indices.to(device=features.device, dtype=torch.int32).contiguous()
    if tree is not None and int(features.size(0)) != tree.n_points:
        raise RuntimeError("loss_cls_3d: `tree` was built for another point count (it cannot be combined with down-sampling)")
    loss, nbr = _Cls3d.apply(features, predictions, idx, int(k), float(lambda_val), tree)
    return (loss, nbr) if return_neighbors else loss
