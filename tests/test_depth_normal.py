"""Next row: plane depth -> depth normal.
CPU tier: the torch restatement (oracle/depth_normal_oracle.py) against vectors recorded from the REFERENCE's own
graphics_utils.normal_from_depth_image + autograd (oracle/make_golden_depth_normal.py).
GPU tier: the fused CUDA kernels (through the C ABI) against the same vectors and, at the headline image size, against
the restatement.  Tolerances (tensor-scale relative): forward 1e-5, gradient 1e-4."""
import os

import numpy as np
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)
from oracle import depth_normal_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_normal.npz")


def _cases():
    z = np.load(GOLD)
    for name in ("a", "b"):
        yield name, {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "_")}


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


@pytest.mark.parametrize("name,c", list(_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_restatement_matches_reference_vectors(name, c):
    fx, fy, cx, cy = (float(v) for v in c["intr"])
    n = orc.depth_to_normal(c["depth"], fx, fy, cx, cy)
    assert _rel(n, c["normal"]) < 1e-5   # the reference inverts K numerically (torch.inverse); the restatement divides
    assert float(n[:, 0, :].abs().max()) == 0.0 and float(n[:, :, -1].abs().max()) == 0.0   # zero border
    assert _rel(orc.depth_to_normal(c["depth"], fx, fy, cx, cy, c["alpha"]), c["normal_alpha"]) < 1e-5
    g = orc.depth_to_normal_backward(c["depth"], fx, fy, cx, cy, c["upstream"], c["alpha"])
    assert _rel(g, c["g_depth"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name,c", list(_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_cuda_matches_reference_vectors(name, c):
    from lsx_b200.render_utils import depth_to_normal
    fx, fy, cx, cy = (float(v) for v in c["intr"])
    depth = c["depth"].cuda().requires_grad_(True)
    n = depth_to_normal(depth, fx, fy, cx, cy)
    assert _rel(n.detach().cpu(), c["normal"]) < 1e-5
    na = depth_to_normal(depth, fx, fy, cx, cy, alpha=c["alpha"].cuda())
    assert _rel(na.detach().cpu(), c["normal_alpha"]) < 1e-5
    na.backward(c["upstream"].cuda())
    assert _rel(depth.grad.cpu(), c["g_depth"]) < 1e-4


@pytest.mark.gpu
def test_cuda_matches_restatement_at_1080p_and_api():
    from lsx_b200.render_utils import depth_to_normal, render_normal
    H, W = 1080, 1920
    g = torch.Generator().manual_seed(5)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 4.0 + 0.001 * xs + 0.002 * ys + 0.05 * torch.rand(H, W, generator=g)
    alpha = torch.rand(H, W, generator=g)
    up = torch.randn(3, H, W, generator=g)
    fx, fy, cx, cy = 1662.8, 1662.8, 960.0, 540.0
    want = orc.depth_to_normal(depth.double(), fx, fy, cx, cy, alpha.double())
    want_g = orc.depth_to_normal_backward(depth.double(), fx, fy, cx, cy, up.double(), alpha.double())
    d = depth.cuda().requires_grad_(True)

    class Cam:
        Fx, Fy, Cx, Cy = fx, fy, cx, cy
    got = render_normal(Cam(), d, alpha=alpha.cuda())
    got.backward(up.cuda())
    assert got.shape == (3, H, W)
    # `want` is evaluated in float64.  At this size the edge vectors are differences of nearly equal back-projected
    # points (depth 4 +- 0.05, rays up to 0.58), so ANY fp32 evaluation — the reference's included — carries ~1e-7 * 4 / 0.05
    # relative cancellation error; the 1e-5 / 1e-4 tolerances are enforced on the reference's own vectors above.
    ef, eg = _rel(got.detach().cpu(), want), _rel(d.grad.cpu(), want_g)
    ef32 = _rel(orc.depth_to_normal(depth, fx, fy, cx, cy, alpha), want)
    assert ef < max(1e-4, 4 * ef32), (ef, ef32)
    assert eg < 1e-3, eg
    with pytest.raises(NotImplementedError):
        render_normal(Cam(), d, scale=2)
    with pytest.raises(RuntimeError):
        depth_to_normal(depth, fx, fy, cx, cy)          # CPU tensor: no fallback
