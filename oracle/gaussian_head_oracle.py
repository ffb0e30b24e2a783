"""CPU restatement (torch) of the per-Gaussian half of the reference's render wrapper.  TEST INFRASTRUCTURE ONLY.

Follows field_construction/scene/gaussian_model.py:53-61 (activations), :193-213 (get_scaling / get_rotation / get_opacity),
:225-236 (get_smallest_axis, get_normal), :257-258 (get_rotation_matrix -> pytorch3d.transforms.quaternion_to_matrix) and
field_construction/gaussian_renderer/__init__.py:188-196 (all_map).
THIRD-PARTY: pytorch3d is not vendored in the reference (requirements pin `pytorch3d`, no version); quaternion_to_matrix is
restated from its published implementation (pytorch3d/transforms/rotation_conversions.py: two_s = 2 / sum(q^2), real part first).
PINNING: tests/golden/gaussian_head.npz is recorded by oracle/make_golden_gaussian_head.py from the reference's OWN
GaussianModel methods executed on CPU (pytorch3d stubbed with the restatement above) + autograd."""
import torch


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def gaussian_head(xyz, scaling_raw, rotation_raw, opacity_raw, viewmatrix, campos):
    scales = torch.exp(scaling_raw)
    rotations = torch.nn.functional.normalize(rotation_raw)
    opacity = torch.sigmoid(opacity_raw)
    R = quaternion_to_matrix(rotations)
    idx = scales.min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    normal = R.gather(2, idx).squeeze(dim=2)
    neg = (normal * (campos - xyz)).sum(-1) < 0.0
    normal = torch.where(neg[:, None], -normal, normal)
    local_normal = normal @ viewmatrix[:3, :3]
    pts = xyz @ viewmatrix[:3, :3] + viewmatrix[3, :3]
    dist = (local_normal * pts).sum(-1).abs()
    all_map = torch.cat([local_normal, torch.ones_like(dist)[:, None], dist[:, None]], dim=1)
    return scales, rotations, opacity, all_map
