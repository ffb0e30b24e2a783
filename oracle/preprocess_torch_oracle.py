"""CPU restatement (torch) of the reference's PYTHON preprocess path — TEST / BASELINE INFRASTRUCTURE ONLY, never imported
by langscene-x_b200/.

The reference's render() can do the per-Gaussian colour and covariance work in PyTorch instead of in the CUDA rasterizer
(pipe.convert_SHs_python / pipe.compute_cov3D_python, field_construction/gaussian_renderer/__init__.py:101-121): that is the
part of the hot path the reference can execute on a CPU.  Restated here:
  sh_to_rgb            gaussian_renderer/__init__.py:114-119 + utils/sh_utils.py:57-113 (eval_sh, degrees 0..3):
                       directions = normalise(xyz - camera_center), colour = clamp_min(eval_sh(...) + 0.5, 0)
  world_covariance     scene/gaussian_model.py:47-51 (build_covariance_from_scaling_rotation) with
                       utils/general_utils.py:66-119 (strip_symmetric, build_rotation, build_scaling_rotation):
                       L = R(q / |q|) diag(mod * s),  Sigma = L L^T, six upper-triangular entries
PINNING: tests/golden/preprocess_torch.npz is recorded by oracle/make_golden_preprocess_torch.py from the reference's OWN functions
executed on CPU (eval_sh is device-agnostic; general_utils' hard-coded device="cuda" is redirected to the CPU there).
bench.py's `cpu_baseline` times these two functions with all host threads, followed by the C compositing oracle."""
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)


def eval_sh(deg, sh, dirs):
    """sh: (P, 3, (deg+1)^2 or more), dirs: (P, 3) unit vectors -> (P, 3)"""
    res = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        res = res - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            res = (res + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] +
                   C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] +
                       C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
                       C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14] +
                       C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return res


def sh_to_rgb(deg, shs, xyz, camera_center):
    """shs: (P, M, 3) as the rasterizer takes them -> colors_precomp (P, 3)"""
    shs_view = shs.transpose(1, 2)                                   # (P, 3, M), the reference's shs_view
    d = xyz - camera_center[None, :]
    d = d / d.norm(dim=1, keepdim=True)
    return torch.clamp_min(eval_sh(deg, shs_view, d) + 0.5, 0.0)


def world_covariance(scaling, scaling_modifier, rotation):
    """(P, 3) activated scales, (P, 4) quaternions -> (P, 6) = cov3D_precomp"""
    q = rotation / torch.sqrt((rotation * rotation).sum(dim=1))[:, None]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)
    L = R * (scaling_modifier * scaling)[:, None, :]                 # R @ diag(s)
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1)
