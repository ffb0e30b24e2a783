"""Fused pieces of LangScene-X's render wrapper on the C ABI.

`render_normal(viewpoint_cam, depth, ...)` mirrors field_construction/gaussian_renderer/__init__.py:28-40
(-> field_construction/utils/graphics_utils.py:65-75 normal_from_depth_image -> :42-63 depth_pcd2normal): the normal of
the surface seen in a (H, W) depth image, (3, H, W), zero on the one-pixel border.  `alpha=` additionally fuses the
`* rendered_alpha.detach()` of the call site (gaussian_renderer/__init__.py:233-235).  One CUDA kernel each way instead of
~15 torch kernels and their autograd graph.  No CPU path.
"""
import ctypes

import torch

from . import _lib


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _DepthToNormal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, alpha, fx, fy, cx, cy):
        if not depth.is_cuda:
            raise RuntimeError("depth must be a CUDA tensor (this operator has no CPU path)")
        if depth.dim() != 2 or depth.dtype != torch.float32:
            raise RuntimeError("depth must be a float32 tensor of shape (H, W)")
        depth = depth.contiguous()
        H, W = depth.shape
        if alpha is not None:
            alpha = alpha.detach().reshape(H, W).to(torch.float32).contiguous()
        out = torch.empty((3, H, W), dtype=torch.float32, device=depth.device)
        with torch.cuda.device(depth.device):
            _lib.check(_lib.load().lsx_depth_normal_forward(W, H, fx, fy, cx, cy, depth.data_ptr(),
                                                            alpha.data_ptr() if alpha is not None else None, out.data_ptr(),
                                                            _stream(depth.device)), "depth_normal")
        ctx.save_for_backward(depth, alpha if alpha is not None else torch.empty(0, device=depth.device))
        ctx.intr = (fx, fy, cx, cy)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        depth, alpha = ctx.saved_tensors
        fx, fy, cx, cy = ctx.intr
        H, W = depth.shape
        grad_out = grad_out.to(torch.float32).contiguous()
        g_depth = torch.empty_like(depth)
        with torch.cuda.device(depth.device):
            _lib.check(_lib.load().lsx_depth_normal_backward(W, H, fx, fy, cx, cy, depth.data_ptr(),
                                                             alpha.data_ptr() if alpha.numel() else None, grad_out.data_ptr(),
                                                             g_depth.data_ptr(), _stream(depth.device)), "depth_normal backward")
        return g_depth, None, None, None, None, None


def depth_to_normal(depth, fx, fy, cx, cy, alpha=None):
    """normal (3, H, W) of the depth image `depth` (H, W) under K = [[fx,0,cx],[0,fy,cy],[0,0,1]], times `alpha` if given."""
    return _DepthToNormal.apply(depth, alpha, float(fx), float(fy), float(cx), float(cy))


def render_normal(viewpoint_cam, depth, offset=None, normal=None, scale=1, alpha=None):
    """Drop-in for the reference's render_normal (same positional arguments).  `viewpoint_cam` needs Fx, Fy, Cx, Cy (as used
    by Camera.get_calib_matrix_nerf, field_construction/scene/cameras.py:153-156).  The reference's optional sampling
    `offset` and the sub-sampling `scale` have no call site in the training loop and are not supported here."""
    if offset is not None or scale != 1:
        raise NotImplementedError("render_normal: offset / scale != 1 are not supported by the fused kernel")
    return depth_to_normal(depth, viewpoint_cam.Fx / scale, viewpoint_cam.Fy / scale, viewpoint_cam.Cx / scale,
                           viewpoint_cam.Cy / scale, alpha=alpha)


class _GaussianHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scaling_raw, rotation_raw, opacity_raw, viewmatrix, campos):
        for name, t in (("xyz", xyz), ("scaling", scaling_raw), ("rotation", rotation_raw), ("opacity", opacity_raw)):
            if not t.is_cuda:
                raise RuntimeError(f"{name} must be a CUDA tensor (this operator has no CPU path)")
            if t.dtype != torch.float32:
                raise RuntimeError(f"{name} must be float32")
        P = int(xyz.shape[0])
        if xyz.shape != (P, 3) or scaling_raw.shape != (P, 3) or rotation_raw.shape != (P, 4) or opacity_raw.numel() != P:
            raise RuntimeError("expected xyz (P,3), scaling (P,3), rotation (P,4), opacity (P,1)")
        xyz, scaling_raw, rotation_raw, opacity_raw = (t.contiguous() for t in (xyz, scaling_raw, rotation_raw, opacity_raw))
        dev = xyz.device
        view = viewmatrix.detach().to(device=dev, dtype=torch.float32).contiguous()   # stays on the device: no host sync
        cam = campos.detach().to(device=dev, dtype=torch.float32).contiguous()
        opts = dict(dtype=torch.float32, device=dev)
        scales, rotations = torch.empty((P, 3), **opts), torch.empty((P, 4), **opts)
        opacity, all_map = torch.empty((P, 1), **opts), torch.empty((P, 5), **opts)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().lsx_gaussian_head_forward(P, view.data_ptr(), cam.data_ptr(), xyz.data_ptr(), scaling_raw.data_ptr(),
                                                             rotation_raw.data_ptr(), opacity_raw.data_ptr(), scales.data_ptr(),
                                                             rotations.data_ptr(), opacity.data_ptr(), all_map.data_ptr(),
                                                             _stream(dev)), "gaussian_head")
        ctx.save_for_backward(xyz, scaling_raw, rotation_raw, opacity_raw)
        ctx.cam = (view, cam)
        return scales, rotations, opacity, all_map

    @staticmethod
    def backward(ctx, g_scales, g_rotations, g_opacity, g_all_map):
        xyz, scaling_raw, rotation_raw, opacity_raw = ctx.saved_tensors
        view, cam = ctx.cam
        P, dev = int(xyz.shape[0]), xyz.device
        ptr = lambda t: None if t is None else t.to(torch.float32).contiguous()
        g_scales, g_rotations, g_opacity, g_all_map = (ptr(t) for t in (g_scales, g_rotations, g_opacity, g_all_map))
        dp = lambda t: None if t is None else t.data_ptr()
        opts = dict(dtype=torch.float32, device=dev)
        g_xyz, g_s, g_r = torch.empty((P, 3), **opts), torch.empty((P, 3), **opts), torch.empty((P, 4), **opts)
        g_o = torch.empty(opacity_raw.shape, **opts)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().lsx_gaussian_head_backward(P, view.data_ptr(), cam.data_ptr(), xyz.data_ptr(), scaling_raw.data_ptr(),
                                                              rotation_raw.data_ptr(), opacity_raw.data_ptr(), dp(g_scales),
                                                              dp(g_rotations), dp(g_opacity), dp(g_all_map), None,
                                                              g_xyz.data_ptr(), g_s.data_ptr(), g_r.data_ptr(), g_o.data_ptr(),
                                                              _stream(dev)), "gaussian_head backward")
        return g_xyz, g_s, g_r, g_o, None, None


def gaussian_head(xyz, scaling_raw, rotation_raw, opacity_raw, viewmatrix, campos):
    """(scales, rotations, opacity, all_map) for the rasterizer from the raw Gaussian parameters and one camera: the fused
    form of get_scaling / get_rotation / get_opacity / get_normal and of the all_map lines of render()
    (field_construction/scene/gaussian_model.py:193-236, field_construction/gaussian_renderer/__init__.py:188-196).
    The position gradient returned for `xyz` is only the part that flows through all_map; the rasterizer's own
    dL/dmeans3D reaches `xyz` through autograd as usual."""
    return _GaussianHead.apply(xyz, scaling_raw, rotation_raw, opacity_raw, viewmatrix, campos)


class _PoseTransform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, xyz, rotation_raw):
        for name, t in (("pose", pose), ("xyz", xyz), ("rotation", rotation_raw)):
            if t is None:
                continue
            if not t.is_cuda:
                raise RuntimeError(f"{name} must be a CUDA tensor (this operator has no CPU path)")
            if t.dtype != torch.float32:
                raise RuntimeError(f"{name} must be float32")
        P = int(xyz.shape[0])
        if pose.numel() != 7 or xyz.shape != (P, 3) or (rotation_raw is not None and rotation_raw.shape != (P, 4)):
            raise RuntimeError("expected pose (7,) = [quaternion | translation], xyz (P,3), rotation (P,4)")
        ctx.pose_shape = pose.shape          # a caller may pass a (1, 7) row slice such as P[idx:idx + 1]
        pose, xyz = pose.reshape(7).contiguous(), xyz.contiguous()
        rot = None if rotation_raw is None else rotation_raw.contiguous()
        dev = xyz.device
        out_xyz = torch.empty_like(xyz)
        out_rot = None if rot is None else torch.empty_like(rot)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().lsx_pose_transform_forward(P, pose.data_ptr(), xyz.data_ptr(), None if rot is None else rot.data_ptr(),
                                                              out_xyz.data_ptr(), None if rot is None else out_rot.data_ptr(),
                                                              _stream(dev)), "pose_transform")
        ctx.save_for_backward(pose, xyz, rot if rot is not None else torch.empty(0, device=dev))
        if rot is None:
            return out_xyz, torch.empty(0, device=dev)
        return out_xyz, out_rot

    @staticmethod
    def backward(ctx, g_xyz, g_rot):
        pose, xyz, rot = ctx.saved_tensors
        P, dev = int(xyz.shape[0]), xyz.device
        has_rot = rot.numel() > 0
        c = lambda t: None if t is None else t.to(torch.float32).contiguous()
        g_xyz, g_rot = c(g_xyz), (c(g_rot) if has_rot else None)
        lib = _lib.load()
        d_xyz = torch.empty_like(xyz)
        d_rot = torch.empty_like(rot) if has_rot else None
        d_pose = torch.empty(7, dtype=torch.float32, device=dev)
        partials = torch.empty(int(lib.lsx_pose_num_partials()), dtype=torch.float32, device=dev)
        dp = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(lib.lsx_pose_transform_backward(P, pose.data_ptr(), xyz.data_ptr(), dp(rot) if has_rot else None, dp(g_xyz),
                                                       dp(g_rot), d_xyz.data_ptr(), dp(d_rot), d_pose.data_ptr(),
                                                       partials.data_ptr(), _stream(dev)), "pose_transform backward")
        return d_pose.view(ctx.pose_shape), d_xyz, d_rot


def pose_transform(camera_pose, xyz, rotation_raw=None):
    """(means3D, rotations) of render(..., camera_pose=pose) (field_construction/gaussian_renderer/__init__.py:79-87):
    means3D = (get_camera_from_tensor(pose) @ [xyz 1]^T)^T[:, :3], rotations = quadmultiply(pose[:4], rotation_raw)
    (field_construction/utils/pose_utils.py:13-107).  camera_pose: (7,) = [quaternion (w,x,y,z) | translation], a row of
    GaussianModel.P; gradients flow to the pose (reduced over all Gaussians inside the backward kernel), xyz and rotation."""
    out_xyz, out_rot = _PoseTransform.apply(camera_pose, xyz, rotation_raw)
    return (out_xyz, out_rot) if rotation_raw is not None else (out_xyz, None)
