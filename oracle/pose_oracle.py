"""CPU restatement (torch) of the pose half of the reference's render wrapper and of its masked language L1.
TEST INFRASTRUCTURE ONLY (never imported by the product).

Follows field_construction/gaussian_renderer/__init__.py:79-87 (camera_pose branch), field_construction/utils/pose_utils.py:
13-58 (quad2rotation), :60-87 (get_camera_from_tensor), :89-107 (quadmultiply); field_construction/gaussian_field.py:450-451 and
field_construction/utils/loss_utils.py:20-21 (l1_loss).
PINNING: tests/golden/pose_l1.npz is recorded by oracle/make_golden_pose.py from the reference's OWN pose_utils / loss_utils
functions executed on CPU (both modules are pure torch) + autograd."""
import torch


def quad2rotation(q):
    q = q / torch.sqrt((q * q).sum(-1, keepdim=True))
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(q.shape[:-1] + (3, 3))


def pose_transform(pose, xyz, rotation_raw):
    R = quad2rotation(pose[None, :4])[0]
    means3D = xyz @ R.T + pose[4:]
    w1, x1, y1, z1 = pose[:4].unbind(-1)
    w2, x2, y2, z2 = rotation_raw.unbind(-1)
    rot = torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                       w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)
    return means3D, rot


def masked_l1(a, b, mask=None):
    if mask is None:
        return (a - b).abs().mean()
    m = mask.to(a.dtype)
    return (a * m - b * m).abs().mean()
