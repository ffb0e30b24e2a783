"""Pose transform (fwd+bwd incl. the pose-gradient reduction) and masked language L1 (fwd+bwd): fused kernels vs torch ops."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.loss import masked_l1_loss  # noqa: E402
from lsx_b200.render_utils import pose_transform  # noqa: E402


def timeit(fn, n=30):
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


def torch_pose(pose, xyz, rot):
    """the reference's formulation (gaussian_renderer/__init__.py:79-87, pose_utils.py:13-107) with torch ops"""
    q = pose[None, :4]
    q = q / torch.sqrt((q * q).sum(-1, keepdim=True))
    r, x, y, z = q[0].unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z),
                     1 - 2 * (x * x + z * z), 2 * (y * z - r * x), 2 * (x * z - r * y), 2 * (y * z + r * x),
                     1 - 2 * (x * x + y * y)]).view(3, 3)
    w2c = torch.eye(4, device=xyz.device)
    w2c[:3, :3] = R
    w2c[:3, 3] = pose[4:]
    homo = torch.cat((xyz.clone(), torch.ones(xyz.shape[0], 1, device=xyz.device)), dim=1)
    m = (w2c @ homo.T).T[:, :3]
    w1, x1, y1, z1 = pose[:4].unbind(-1)
    w2, x2, y2, z2 = rot.clone().unbind(-1)
    o = torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)
    return m, o


for P in (500_000, 1_000_000, 5_000_000):
    pose = torch.randn(7, device="cuda").requires_grad_(True)
    xyz, rot = torch.randn(P, 3, device="cuda").requires_grad_(True), torch.randn(P, 4, device="cuda").requires_grad_(True)
    ux, ur = torch.randn(P, 3, device="cuda"), torch.randn(P, 4, device="cuda")

    def run(fn):
        def f():
            pose.grad = xyz.grad = rot.grad = None
            torch.autograd.backward(fn(pose, xyz, rot), (ux, ur))
        return f

    tf, tt = timeit(run(pose_transform)), timeit(run(torch_pose))
    print(json.dumps({"op": "pose_transform fwd+bwd", "P": P, "fused_ms": round(tf, 4), "torch_ops_ms": round(tt, 4),
                      "speedup": round(tt / tf, 2), "algorithmic_GBps": round(P * (56 + 56 + 28) / tf / 1e6, 1)}), flush=True)

for C, H, W in ((3, 480, 720), (16, 1080, 1920)):
    a = torch.randn(C, H, W, device="cuda").requires_grad_(True)
    b = torch.randn(C, H, W, device="cuda")
    mask = torch.rand(H, W, device="cuda") > 0.3

    def fused():
        a.grad = None
        masked_l1_loss(a, b, mask).backward()

    def ref():
        a.grad = None
        torch.abs(a * mask - b * mask).mean().backward()

    tf, tt = timeit(fused), timeit(ref)
    n = C * H * W
    print(json.dumps({"op": "masked_l1 fwd+bwd", "shape": [C, H, W], "fused_ms": round(tf, 4), "torch_ops_ms": round(tt, 4),
                      "speedup": round(tt / tf, 2), "algorithmic_GBps": round((n * 8 + n * 12 + 2 * H * W * 4) / tf / 1e6, 1)}),
          flush=True)
