"""Next rows (SURVEY.md 8f.1 pose half, 8f.2 language loss): camera-pose transform of the Gaussians and the masked L1.
CPU tier: the torch restatement (oracle/pose_oracle.py) against vectors recorded from the REFERENCE's own pose_utils /
loss_utils functions + autograd (oracle/make_golden_pose.py).  GPU tier: the CUDA kernels (C ABI) against the same vectors and,
at 1 M Gaussians / 1080p, against the restatement in float64.  Tolerances: forward 1e-5, gradients 1e-4 (tensor-scale relative)."""
import os

import numpy as np
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)
from oracle import pose_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_l1.npz")


def _case(name):
    z = np.load(GOLD)
    return {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "_")}


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


def _run_pose(fn, c, device, dtype=torch.float32):
    leaf = lambda t: t.to(device=device, dtype=dtype).clone().requires_grad_(True)
    pose, xyz, rot = leaf(c["pose"]), leaf(c["xyz"]), leaf(c["rot"])
    m, r = fn(pose, xyz, rot)
    ((m * c["up_means3D"].to(device=device, dtype=dtype)).sum() + (r * c["up_rotations"].to(device=device, dtype=dtype)).sum()).backward()
    return (m.detach(), r.detach()), (pose.grad, xyz.grad, rot.grad)


@pytest.mark.parametrize("name", ["a", "b"])
def test_restatement_pose_matches_reference_vectors(name):
    c = _case(name)
    (m, r), (gp, gx, gr) = _run_pose(orc.pose_transform, c, "cpu")
    assert _rel(m, c["means3D"]) < 1e-6 and _rel(r, c["rotations"]) < 1e-6
    assert _rel(gp, c["g_pose"]) < 1e-5 and _rel(gx, c["g_xyz"]) < 1e-5 and _rel(gr, c["g_rot"]) < 1e-5


@pytest.mark.parametrize("name", ["l3", "l16"])
def test_restatement_masked_l1_matches_reference_vectors(name):
    c = _case(name)
    lf = c["lf"].clone().requires_grad_(True)
    loss = orc.masked_l1(lf, c["gt"], c["mask"])
    (loss * 1.7).backward()
    assert abs(float(loss) - float(c["loss"])) < 1e-7 and torch.equal(lf.grad, c["g_lf"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b"])
def test_cuda_pose_matches_reference_vectors(name):
    from lsx_b200.render_utils import pose_transform
    c = _case(name)
    (m, r), (gp, gx, gr) = _run_pose(pose_transform, c, "cuda:0")
    assert _rel(m.cpu(), c["means3D"]) < 1e-5 and _rel(r.cpu(), c["rotations"]) < 1e-5
    assert _rel(gp.cpu(), c["g_pose"]) < 1e-4 and _rel(gx.cpu(), c["g_xyz"]) < 1e-4 and _rel(gr.cpu(), c["g_rot"]) < 1e-4


@pytest.mark.gpu
def test_cuda_pose_1M_against_float64_restatement_and_determinism():
    from lsx_b200.render_utils import pose_transform
    P = 1_000_000
    g = torch.Generator().manual_seed(5)
    c = {"pose": torch.cat([torch.randn(4, generator=g), torch.randn(3, generator=g)]), "xyz": torch.randn(P, 3, generator=g) * 3,
         "rot": torch.randn(P, 4, generator=g), "up_means3D": torch.randn(P, 3, generator=g) / P,
         "up_rotations": torch.randn(P, 4, generator=g) / P}
    (m64, r64), (gp64, gx64, gr64) = _run_pose(orc.pose_transform, c, "cpu", torch.float64)
    (m, r), (gp, gx, gr) = _run_pose(pose_transform, c, "cuda:0")
    assert _rel(m.cpu(), m64) < 1e-5 and _rel(r.cpu(), r64) < 1e-5
    assert _rel(gx.cpu(), gx64) < 1e-4 and _rel(gr.cpu(), gr64) < 1e-4
    assert _rel(gp.cpu(), gp64) < 1e-4          # 1 M-term reduction: fp32 partials per thread, blocks added in double
    (_, _), (gp2, _, _) = _run_pose(pose_transform, c, "cuda:0")
    assert torch.equal(gp, gp2)                 # fixed reduction order
    # xyz only (no rotation): the rotation output / gradient are absent
    pose = c["pose"].cuda().requires_grad_(True)
    m2, none = pose_transform(pose, c["xyz"].cuda(), None)
    assert none is None and torch.equal(m2, m)
    with pytest.raises(RuntimeError):
        pose_transform(c["pose"], c["xyz"], c["rot"])      # CPU tensors


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["l3", "l16"])
def test_cuda_masked_l1_matches_reference_vectors(name):
    from lsx_b200.loss import masked_l1_loss
    c = _case(name)
    lf = c["lf"].cuda().requires_grad_(True)
    loss = masked_l1_loss(lf, c["gt"].cuda(), c["mask"].cuda())
    (loss * 1.7).backward()
    assert abs(float(loss) - float(c["loss"])) < 1e-5 * max(1.0, abs(float(c["loss"])))
    assert _rel(lf.grad.cpu(), c["g_lf"]) < 1e-6 and torch.equal(lf.grad.cpu() == 0, c["g_lf"] == 0)


@pytest.mark.gpu
def test_cuda_masked_l1_1080p_against_restatement():
    from lsx_b200.loss import masked_l1_loss
    C, H, W = 16, 1080, 1920
    g = torch.Generator().manual_seed(9)
    lf, gt = torch.randn(C, H, W, generator=g), torch.randn(C, H, W, generator=g)
    mask = torch.rand(H, W, generator=g) > 0.3
    for mk in (mask, None, torch.rand(C, H, W, generator=g)):
        a = lf.cuda().requires_grad_(True)
        loss = masked_l1_loss(a, gt.cuda(), None if mk is None else mk.cuda())
        loss.backward()
        a64 = lf.double().requires_grad_(True)
        want = orc.masked_l1(a64, gt.double(), None if mk is None else mk.double())
        want.backward()
        assert abs(float(loss) - float(want)) < 1e-5 * float(want)
        assert _rel(a.grad.cpu(), a64.grad) < 1e-6
    with pytest.raises(RuntimeError):
        masked_l1_loss(lf, gt, mask)                        # CPU tensors
