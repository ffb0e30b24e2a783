"""Fused image losses of LangScene-X's training loop on the C ABI.

  ssim(img1, img2)                  drop-in for field_construction/utils/loss_utils.py:37-46 (11x11 window, size_average)
  image_loss(image, gt, lambda_)    (1 - lambda_) * l1_loss(image, gt) + lambda_ * (1 - ssim(image, gt)), the combination at
                                    field_construction/gaussian_field.py:238-246, one kernel forward and one backward;
                                    returns (loss, l1, ssim_value) so that the loop can keep logging both terms.
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
