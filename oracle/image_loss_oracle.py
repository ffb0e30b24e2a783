"""CPU restatement (torch) of the reference's image losses.  TEST INFRASTRUCTURE ONLY (imported by tests/ and tools/).

Follows field_construction/utils/loss_utils.py: l1_loss :20-21; gaussian :32-34; create_window :37-41 (2-D window = outer
product of the normalised 1-D float32 Gaussian); ssim / _ssim :44-75 (five depthwise conv2d with zero padding 5).
PINNING: tests/golden/image_loss.npz holds outputs and autograd gradients of the reference's OWN functions, recorded on
CPU by oracle/make_golden_image_loss.py."""
from math import exp

import torch
import torch.nn.functional as F


def window(channel, dtype=torch.float32, device="cpu"):
    g = torch.tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, 11, 11).contiguous().to(dtype=dtype, device=device)


def ssim(img1, img2):
    C = img1.size(-3)
    w = window(C, img1.dtype, img1.device)
    conv = lambda t: F.conv2d(t, w, padding=5, groups=C)
    mu1, mu2 = conv(img1), conv(img2)
    s11 = conv(img1 * img1) - mu1.pow(2)
    s22 = conv(img2 * img2) - mu2.pow(2)
    s12 = conv(img1 * img2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1.pow(2) + mu2.pow(2) + C1) * (s11 + s22 + C2))
    return m.mean()


def l1_loss(a, b):
    return torch.abs(a - b).mean()


def image_loss(image, gt, lam):
    return (1.0 - lam) * l1_loss(image, gt) + lam * (1.0 - ssim(image, gt))
