"""Seeded synthetic Gaussian scenes and cameras (the generator specified in SURVEY.md §8d).

Everything is generated on the CPU with a seeded torch.Generator, so the same seed gives the same scene
on every machine; `.to(device)` afterwards.  Camera conventions follow the reference:
  world_view_transform = W2C transposed, full_proj_transform = world_view_transform @ P^T
  (field_construction/scene/cameras.py:109-112), P from getProjectionMatrix
  (field_construction/utils/graphics_utils.py:135-155) with znear=0.01, zfar=100.
"""
import math
from dataclasses import dataclass
from typing import Optional

import torch

# name -> (P, W, H, F_lang, n_views, s_med (None = 10/fx))
CONFIGS = {
    "C1": dict(P=10_000, W=256, H=256, F=3, views=1, s_med=None),
    "C2": dict(P=100_000, W=800, H=800, F=3, views=1, s_med=None),
    "C3": dict(P=1_000_000, W=1920, H=1080, F=16, views=1, s_med=None),
    "C4": dict(P=500_000, W=720, H=480, F=3, views=49, s_med=0.006),
    "C5": dict(P=5_000_000, W=1920, H=1080, F=16, views=64, s_med=0.006),
}


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # (4,4) W2C transposed
    projmatrix: torch.Tensor   # (4,4) viewmatrix @ P^T
    campos: torch.Tensor       # (3,)

    def to(self, device):
        return Camera(self.W, self.H, self.tanfovx, self.tanfovy, self.viewmatrix.to(device),
                      self.projmatrix.to(device), self.campos.to(device))


@dataclass
class Scene:
    means3D: torch.Tensor       # (P,3)
    scales: torch.Tensor        # (P,3)  (already exp-activated)
    rotations: torch.Tensor     # (P,4)  unit quaternions (w,x,y,z)
    opacities: torch.Tensor     # (P,1)  (already sigmoid-activated)
    shs: torch.Tensor           # (P,16,3)
    language_feature: torch.Tensor   # (P,F)
    instance_feature: torch.Tensor   # (P,3)
    normals: torch.Tensor       # (P,3) unit world-space normals used to build all_map

    def to(self, device):
        return Scene(*[getattr(self, f).to(device) for f in self.__dataclass_fields__])


def projection_matrix(znear, zfar, tanfovx, tanfovy):
    top, right = tanfovy * znear, tanfovx * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W, H, yaw_deg=0.0, fovx_deg=60.0, centre=(0.0, 0.0, 4.0), radius=4.0):
    """Camera on an arc of `radius` around `centre`, looking at it; yaw 0 = at the origin looking down +z."""
    tanfovx = math.tan(math.radians(fovx_deg) * 0.5)
    tanfovy = tanfovx * H / W  # square pixels
    yaw = math.radians(yaw_deg)
    c, s = math.cos(yaw), math.sin(yaw)
    # camera-to-world rotation: yaw about the world y axis
    R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)
    ctr = torch.tensor(centre, dtype=torch.float64)
    cam_pos = ctr - R @ torch.tensor([0.0, 0.0, radius], dtype=torch.float64)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = R.t()
    w2c[:3, 3] = -(R.t() @ cam_pos)
    viewmatrix = w2c.t().float().contiguous()
    proj = projection_matrix(0.01, 100.0, tanfovx, tanfovy).t()
    full = (viewmatrix.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = viewmatrix.inverse()[3, :3].contiguous()
    return Camera(W, H, tanfovx, tanfovy, viewmatrix, full, campos)


def make_cameras(W, H, n_views):
    if n_views == 1:
        return [make_camera(W, H, 0.0)]
    return [make_camera(W, H, -15.0 + 30.0 * v / (n_views - 1)) for v in range(n_views)]


def make_scene(P, W, H, F=3, seed=0, s_med: Optional[float] = None, fovx_deg=60.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    tanfovx = math.tan(math.radians(fovx_deg) * 0.5)
    tanfovy = tanfovx * H / W
    fx = W / (2.0 * tanfovx)
    if s_med is None:
        s_med = 10.0 / fx
    hx, hy = 1.04 * 4.0 * tanfovx, 1.04 * 4.0 * tanfovy
    u = torch.rand(P, 3, generator=g)
    xyz = torch.empty(P, 3)
    xyz[:, 0] = (u[:, 0] * 2 - 1) * hx
    xyz[:, 1] = (u[:, 1] * 2 - 1) * hy
    xyz[:, 2] = 2.0 + 4.0 * u[:, 2]
    near = torch.rand(P, generator=g) < 0.05          # 5 % of the points exercise the near cull
    znear = -0.5 + 0.7 * torch.rand(P, generator=g)
    xyz[:, 2] = torch.where(near, znear, xyz[:, 2])
    log_scale = math.log(s_med) + 0.6 * torch.randn(P, 3, generator=g)
    scales = torch.exp(log_scale)
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g))
    shs = torch.empty(P, 16, 3)
    shs[:, 0, :] = 0.5 * torch.randn(P, 3, generator=g)
    shs[:, 1:, :] = 0.1 * torch.randn(P, 15, 3, generator=g)
    lang = torch.randn(P, F, generator=g)
    inst = torch.randn(P, 3, generator=g)
    n = torch.randn(P, 3, generator=g)
    normals = n / n.norm(dim=1, keepdim=True)
    return Scene(xyz.contiguous(), scales.contiguous(), rotations.contiguous(), opacities.contiguous(),
                 shs.contiguous(), lang.contiguous(), inst.contiguous(), normals.contiguous())


def make_all_map(scene: Scene, cam: Camera):
    """[local normal(3), 1, |n . p_cam|] as built by field_construction/gaussian_renderer/__init__.py:188-196."""
    Rw = cam.viewmatrix[:3, :3]
    # normals face the camera, as GaussianModel.get_normal flips them in the reference
    away = ((scene.means3D - cam.campos) * scene.normals).sum(-1, keepdim=True) > 0
    normals = torch.where(away, -scene.normals, scene.normals)
    local_normal = normals @ Rw
    pts_cam = scene.means3D @ Rw + cam.viewmatrix[3, :3]
    dist = (local_normal * pts_cam).sum(-1).abs()
    am = torch.zeros(scene.means3D.shape[0], 5, dtype=torch.float32, device=scene.means3D.device)
    am[:, :3] = local_normal
    am[:, 3] = 1.0
    am[:, 4] = dist
    return am


def make_upstream_grads(W, H, F, Fi=3, seed=1, device="cpu"):
    """dL/d(out) ~ N(0,1)/(W*H) for the 5 differentiable outputs."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    s = 1.0 / (W * H)
    mk = lambda c: (torch.randn(c, H, W, generator=g) * s).to(device)
    return dict(color=mk(3), language_feature=mk(F), instance_feature=mk(Fi), all_map=mk(5), plane_depth=mk(1))
