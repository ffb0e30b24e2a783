"""CPU restatement (torch, fp32 or fp64) of the reference's depth -> normal path.

TEST INFRASTRUCTURE ONLY: imported by tests/ only, never by langscene-x_b200/.
Follows field_construction/utils/graphics_utils.py:
  :16-22  ndc_2_cam            cam = [x_ndc (W-1) z, y_ndc (H-1) z, z] @ inverse(K^T)
  :24-35  depth2point_cam      x_ndc = x / (W-1), y_ndc = y / (H-1)          => cam = z ((x-cx)/fx, (y-cy)/fy, 1)
  :42-63  depth_pcd2normal     offset == None: cross(right - left, top - bottom), F.normalize (eps 1e-12), zero pad 1 px
  :65-75  normal_from_depth_image, and gaussian_renderer/__init__.py:28-40 render_normal -> permute to (3, H, W).
PINNING: checked against the reference's own functions (imported on CPU from /root/reference by
oracle/make_golden_depth_normal.py) through tests/golden/depth_normal.npz, forward and autograd backward.
"""
import torch


def depth_to_normal(depth, fx, fy, cx, cy, alpha=None):
    H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=depth.dtype, device=depth.device),
                            torch.arange(W, dtype=depth.dtype, device=depth.device), indexing="ij")
    P = torch.stack([(xs - cx) / fx * depth, (ys - cy) / fy * depth, depth], dim=-1)
    l2r = P[1:-1, 2:] - P[1:-1, :-2]
    b2t = P[:-2, 1:-1] - P[2:, 1:-1]
    n = torch.nn.functional.normalize(torch.cross(l2r, b2t, dim=-1), p=2, dim=-1)
    n = torch.nn.functional.pad(n.permute(2, 0, 1), (1, 1, 1, 1), mode="constant")
    if alpha is not None:
        n = n * alpha.detach().reshape(1, H, W)
    return n


def depth_to_normal_backward(depth, fx, fy, cx, cy, grad_out, alpha=None):
    d = depth.detach().clone().requires_grad_(True)
    depth_to_normal(d, fx, fy, cx, cy, alpha).backward(grad_out)
    return d.grad
