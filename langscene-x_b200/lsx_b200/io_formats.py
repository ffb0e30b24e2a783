"""On-disk formats of LangScene-X scenes for the arena-based loop.

  save_ply / load_ply            GaussianModel.save_ply / load_ply   field_construction/scene/gaussian_model.py:400-441,448-504
  capture / restore              GaussianModel.capture / restore     gaussian_model.py:90-191  (the `(params, iteration)` checkpoint tuple
                                                                     torch.save'd at field_construction/gaussian_field.py:546-549)
  read_camera_npz                read_camera_npz                     field_construction/scene/dataset_readers.py:234-296
  load_language_feature          Camera.get_language_feature         field_construction/scene/cameras.py:137-151

The PLY body is ONE row of float32 per Gaussian: x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_* [language_feature_*
instance_feature_*], SH coefficients channel-major (`transpose(1, 2).flatten`).  The reference builds it on the host with a
Python tuple per row; here the rows are packed / unpacked on the device by one gather / scatter over the arena
(lsx_rows_pack / lsx_rows_unpack, csrc/pose.cu) and cross PCIe as a single copy.  The header is the one plyfile writes for an
all-'f4' vertex element (plyfile is not vendored in the reference and absent here: its binary_little_endian layout is restated;
the reader accepts any PLY whose vertex element holds scalar properties, by name, like load_ply does).
File I/O and header parsing are host code (numpy); the packing has no CPU path.
"""
import ctypes
import os
import re
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .densify import ParamArena, arena_rows, arena_widths
from .multiview import DensifyStats, GradArena

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def ply_attributes(n_dc: int, n_rest: int, F: int, Fi: int, include_feature: bool) -> List[str]:
    """construct_list_of_attributes (gaussian_model.py:400-413)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    names += [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    if include_feature:
        names += [f"language_feature_{i}" for i in range(F)] + [f"instance_feature_{i}" for i in range(Fi)]
    return names


def _column_map(arena: GradArena, names: List[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(col_begin int64, col_stride int32) device arrays: PLY column -> element of the arena row, for either naming scheme
    (the reference's optimizer groups xyz / f_dc / f_rest / ... or lsx_b200.multiview's means3D / sh / ...)."""
    w = arena_widths(arena)
    off = {n: o for n, (o, _) in arena.offsets.items()}
    pick = lambda *cands: next((c for c in cands if w.get(c, 0) > 0), None)
    g_xyz, g_op, g_sc, g_rot = pick("xyz", "means3D"), pick("opacity"), pick("scaling", "scales"), pick("rotation", "rotations")
    g_lang, g_inst = pick("language_feature"), pick("instance_feature")
    if None in (g_xyz, g_op, g_sc, g_rot):
        raise RuntimeError(f"arena lacks position / opacity / scale / rotation groups (has {list(w)})")
    if pick("sh") and not pick("f_dc"):
        M = w["sh"] // 3                                              # (P, M, 3): coefficient-major, channel-minor
        sh_of = lambda k, c: ("sh", 3 * k + c)
        n_dc, n_rest = 3, 3 * (M - 1)
    else:
        K = w.get("f_rest", 0) // 3
        sh_of = lambda k, c: ("f_dc", c) if k == 0 else ("f_rest", 3 * (k - 1) + c)
        n_dc, n_rest = w["f_dc"], w.get("f_rest", 0)
    begin, stride = [], []
    for name in names:
        grp, j = None, 0
        if name in ("x", "y", "z"):
            grp, j = g_xyz, "xyz".index(name)
        elif name in ("nx", "ny", "nz"):
            grp = None
        elif name == "opacity":
            grp = g_op
        else:
            base, idx = name.rsplit("_", 1)
            idx = int(idx)
            if base == "f_dc":
                grp, j = sh_of(0, idx)
            elif base == "f_rest":                                    # f_rest_{c * K + (k - 1)} = features_rest[k - 1][c]
                K = n_rest // 3
                grp, j = sh_of(1 + idx % K, idx // K)
            elif base == "scale":
                grp, j = g_sc, idx
            elif base == "rot":
                grp, j = g_rot, idx
            elif base == "language_feature":
                grp, j = g_lang, idx
            elif base == "instance_feature":
                grp, j = g_inst, idx
            else:
                raise RuntimeError(f"unknown PLY attribute {name}")
            if grp is None or j >= w[grp]:
                raise RuntimeError(f"PLY attribute {name} has no source in the arena")
        begin.append(-1 if grp is None else off[grp] + j)
        stride.append(0 if grp is None else w[grp])
    dev = arena.flat.device
    return (torch.tensor(begin, dtype=torch.int64, device=dev), torch.tensor(stride, dtype=torch.int32, device=dev))


def _sh_counts(arena: GradArena) -> Tuple[int, int]:
    w = arena_widths(arena)
    if w.get("sh", 0) > 0 and w.get("f_dc", 0) == 0:
        return 3, w["sh"] - 3
    return w["f_dc"], w.get("f_rest", 0)


def pack_rows(arena: GradArena, names: List[str]) -> torch.Tensor:
    """(P, len(names)) float32 CUDA tensor: the PLY body of the arena."""
    if not arena.flat.is_cuda:
        raise RuntimeError("pack_rows needs a CUDA arena (this operator has no CPU path)")
    P, dev = arena_rows(arena), arena.flat.device
    begin, stride = _column_map(arena, names)
    rows = torch.empty((P, len(names)), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsx_rows_pack(P, len(names), begin.data_ptr(), stride.data_ptr(), arena.flat.data_ptr(),
                                             rows.data_ptr(), _stream(dev)), "rows_pack")
    return rows


def unpack_rows(rows: torch.Tensor, names: List[str], arena: GradArena) -> GradArena:
    if not (arena.flat.is_cuda and rows.is_cuda):
        raise RuntimeError("unpack_rows needs CUDA tensors (this operator has no CPU path)")
    rows = rows.to(torch.float32).contiguous()
    P, dev = arena_rows(arena), arena.flat.device
    if rows.shape != (P, len(names)):
        raise RuntimeError("rows must have shape (P, len(names))")
    begin, stride = _column_map(arena, names)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().lsx_rows_unpack(P, len(names), begin.data_ptr(), stride.data_ptr(), rows.data_ptr(),
                                               arena.flat.data_ptr(), _stream(dev)), "rows_unpack")
    return arena


def ply_header(n: int, names: List[str]) -> bytes:
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + [f"property float {a}" for a in names]
    return ("\n".join(lines) + "\nend_header\n").encode("ascii")


def save_ply(path: str, arena: GradArena, include_feature: bool = False):
    """GaussianModel.save_ply for an arena: same attribute list, order and byte layout."""
    w = arena_widths(arena)
    n_dc, n_rest = _sh_counts(arena)
    names = ply_attributes(n_dc, n_rest, w.get("language_feature", 0), w.get("instance_feature", 0), include_feature)
    rows = pack_rows(arena, names)
    host = torch.empty(rows.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(rows, non_blocking=False)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(ply_header(rows.shape[0], names))
        host.numpy().astype("<f4", copy=False).tofile(f)
    return names


def read_ply_vertices(path: str) -> Tuple[List[str], np.ndarray]:
    """(property names, (N, n_props) float32 array) of the vertex element of a binary little-endian or ASCII PLY."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError(f"{path}: the first element must be `vertex`")
        el = elements[0]
        names = [n for n, _ in el["props"]]
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, max_rows=el["count"], ndmin=2).astype(np.float32)
        elif fmt in ("binary_little_endian", "binary_big_endian"):
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(n, end + t) for n, t in el["props"]])
            rec = np.fromfile(f, dtype=dt, count=el["count"])
            if rec.shape[0] != el["count"]:
                raise ValueError(f"{path}: truncated PLY body")
            if all(t == "f4" for _, t in el["props"]) and end == "<":
                data = rec.view("<f4").reshape(el["count"], len(names))
            else:
                data = np.stack([rec[n].astype(np.float32) for n in names], axis=1) if names else np.zeros((el["count"], 0), np.float32)
        else:
            raise ValueError(f"{path}: unknown PLY format {fmt}")
    return names, np.ascontiguousarray(data, dtype=np.float32)


def load_ply(path: str, device, max_sh_degree: int = 3) -> ParamArena:
    """GaussianModel.load_ply into a ParamArena with the reference's groups (xyz, f_dc, f_rest, opacity, scaling, rotation,
    language_feature, instance_feature); columns are looked up by name, sorted by their trailing integer, like the reference."""
    names, data = read_ply_vertices(path)
    count = lambda prefix: len([n for n in names if re.fullmatch(prefix + r"_\d+", n)])
    n_rest = count("f_rest")
    if n_rest != 3 * (max_sh_degree + 1) ** 2 - 3:                    # gaussian_model.py:464
        raise AssertionError(f"{path}: {n_rest} f_rest columns, expected {3 * (max_sh_degree + 1) ** 2 - 3}")
    widths = OrderedDict([("xyz", 3), ("f_dc", 3), ("f_rest", n_rest), ("opacity", 1), ("scaling", count("scale")),
                          ("rotation", count("rot")), ("language_feature", count("language_feature")),
                          ("instance_feature", count("instance_feature"))])
    dev = torch.device(device)
    arena = ParamArena.allocate(data.shape[0], {k: v for k, v in widths.items()}, dev)
    rows = torch.from_numpy(data).pin_memory().to(dev, non_blocking=False)
    unpack_rows(rows, names, arena)
    return arena


# ---------------------------------------------------------------------------------------------------------------------------
# checkpoint tuple
# ---------------------------------------------------------------------------------------------------------------------------
_REF_GROUPS = ("xyz", "knn_f", "f_dc", "f_rest", "opacity", "scaling", "rotation", "language_feature", "instance_feature")
_REF_SHAPES = {"f_dc": lambda t: t.reshape(t.shape[0], -1, 3), "f_rest": lambda t: t.reshape(t.shape[0], -1, 3)}


def capture(params: GradArena, exp_avg: Optional[GradArena], exp_avg_sq: Optional[GradArena], stats: DensifyStats,
            active_sh_degree: int, step: int, lrs: Dict[str, float], spatial_lr_scale: float, poses: torch.Tensor,
            cam_optimizer_state: Optional[dict] = None, include_feature: bool = True, max_weight: Optional[torch.Tensor] = None):
    """The tuple GaussianModel.capture returns (20 entries with features, 18 without), built from arenas with the reference's
    group names, so that `torch.save((capture(...), iteration), path)` is a checkpoint the reference's restore() reads.
    Optimizer state uses torch.optim.Adam's state_dict layout (one parameter per group, in the reference's group order)."""
    w = arena_widths(params)
    P = arena_rows(params)
    dev = params.flat.device

    def grp(arena, name):
        if w.get(name, 0) > 0:
            t = arena.views[name].detach().clone()
        else:
            t = torch.zeros((P, 6 if name == "knn_f" else 0), device=dev)          # knn_f: unused 6-d parameter (:285-287)
        return _REF_SHAPES.get(name, lambda x: x)(t)

    # The reference pickles nn.Parameters and restore() assigns them straight to self._xyz etc. (gaussian_model.py:139-191),
    # so resumed training only gets gradients if the entries ARE Parameters with the reference's requires_grad flags
    # (instance_feature starts frozen, gaussian_model.py:301; change_reqiures_grad switches the sets later).
    tensors = {n: torch.nn.Parameter(grp(params, n), requires_grad=(n != "instance_feature")) for n in _REF_GROUPS}
    # param_groups come from a real torch.optim.Adam over placeholders, so that the dict has exactly the keys this torch
    # version's Adam.load_state_dict expects (the reference builds its optimizer the same way, gaussian_model.py:313-328)
    template = torch.optim.Adam([{"params": [torch.nn.Parameter(torch.empty(0))], "lr": float(lrs.get(n, 0.0)), "name": n}
                                 for n in _REF_GROUPS], lr=0.0, eps=1e-15)
    groups = template.state_dict()["param_groups"]
    state = {}
    for i, n in enumerate(_REF_GROUPS):
        if exp_avg is not None and w.get(n, 0) > 0:
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": grp(exp_avg, n), "exp_avg_sq": grp(exp_avg_sq, n)}
    opt_dict = {"state": state, "param_groups": groups}
    if cam_optimizer_state is not None:
        cam_dict = cam_optimizer_state
    else:
        # the reference's cam_optimizer has ONE group ("pose", lr = rotation_lr * 0.1, gaussian_model.py:326-330);
        # load_state_dict rejects a state dict with a different number of groups, so the default is a fresh Adam over P
        cam_tmpl = torch.optim.Adam([{"params": [torch.nn.Parameter(torch.empty(0))],
                                      "lr": float(lrs.get("pose", 0.1 * float(lrs.get("rotation", 0.0)))), "name": "pose"}],
                                    lr=0.0, eps=1e-15)
        cam_dict = cam_tmpl.state_dict()
    col = lambda t: t.detach().clone().reshape(P, 1)
    head = (active_sh_degree, tensors["xyz"], tensors["knn_f"], tensors["f_dc"], tensors["f_rest"], tensors["scaling"],
            tensors["rotation"], tensors["opacity"])
    feat = (tensors["language_feature"], tensors["instance_feature"]) if include_feature else ()
    mw = max_weight if max_weight is not None else torch.zeros(P, device=dev)
    tail = (stats.max_radii2D.detach().clone(), mw, col(stats.grad_accum), col(stats.grad_accum_abs), col(stats.denom),
            col(stats.denom), opt_dict, cam_dict, spatial_lr_scale, poses)
    return head + feat + tail


def restore(model_args, device) -> dict:
    """Inverse of capture for a tuple written by either side (GaussianModel.restore, gaussian_model.py:139-191): returns
    dict(params, exp_avg, exp_avg_sq (ParamArenas, moments None if the optimizer state is empty), stats, active_sh_degree, step,
    lrs, spatial_lr_scale, poses, cam_optimizer_state, max_weight)."""
    if len(model_args) == 20:
        (sh_deg, xyz, knn_f, f_dc, f_rest, scaling, rotation, opacity, lang, inst, max_radii2D, max_weight, acc, acc_abs, denom,
         _denom_abs, opt_dict, cam_dict, spatial_lr_scale, poses) = model_args
    elif len(model_args) == 18:
        (sh_deg, xyz, knn_f, f_dc, f_rest, scaling, rotation, opacity, max_radii2D, max_weight, acc, acc_abs, denom, _denom_abs,
         opt_dict, cam_dict, spatial_lr_scale, poses) = model_args
        lang = inst = None
    else:
        raise ValueError(f"checkpoint tuple has {len(model_args)} entries, expected 18 or 20")
    dev = torch.device(device)
    P = xyz.shape[0]
    src = {"xyz": xyz, "knn_f": knn_f, "f_dc": f_dc, "f_rest": f_rest, "opacity": opacity, "scaling": scaling,
           "rotation": rotation, "language_feature": lang, "instance_feature": inst}
    flat2 = lambda t: t.detach().reshape(P, -1).to(device=dev, dtype=torch.float32)
    widths = OrderedDict((n, flat2(src[n]).shape[1]) for n in _REF_GROUPS if src[n] is not None and src[n].numel() > 0)
    params = ParamArena.allocate(P, widths, dev)
    for n in widths:
        params.views[n].copy_(flat2(src[n]))
    m = v = None
    step, lrs = 0, {}
    names = [g.get("name", _REF_GROUPS[i] if i < len(_REF_GROUPS) else str(i)) for i, g in enumerate(opt_dict.get("param_groups", []))]
    for g, n in zip(opt_dict.get("param_groups", []), names):
        lrs[n] = float(g.get("lr", 0.0))
    if opt_dict.get("state"):
        m, v = ParamArena.allocate(P, widths, dev), ParamArena.allocate(P, widths, dev)
        for g, n in zip(opt_dict["param_groups"], names):
            st = opt_dict["state"].get(g["params"][0])
            if st is None or n not in widths:
                continue
            m.views[n].copy_(flat2(st["exp_avg"]))
            v.views[n].copy_(flat2(st["exp_avg_sq"]))
            step = max(step, int(float(st["step"])))
    f1 = lambda t: t.detach().reshape(-1).to(device=dev, dtype=torch.float32).clone()
    stats = DensifyStats(f1(acc), f1(acc_abs), f1(denom), f1(max_radii2D))
    return {"params": params, "exp_avg": m, "exp_avg_sq": v, "stats": stats, "active_sh_degree": int(sh_deg), "step": step,
            "lrs": lrs, "spatial_lr_scale": spatial_lr_scale, "poses": poses, "cam_optimizer_state": cam_dict,
            "max_weight": max_weight}


# ---------------------------------------------------------------------------------------------------------------------------
# cameras and language features
# ---------------------------------------------------------------------------------------------------------------------------
def _rotmat_to_qvec(R: np.ndarray) -> np.ndarray:
    """(w, x, y, z) of a rotation matrix with scipy's convention (the reference calls scipy's Rotation.from_matrix(...).as_quat()
    and reorders x,y,z,w -> w,x,y,z, dataset_readers.py:251-253)."""
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(R).as_quat()
    return np.array([q[3], q[0], q[1], q[2]])


def read_camera_npz(camera_dir: str) -> List[dict]:
    """read_camera_npz: every `*.npz` holds `pose` (camera-to-world 4x4) and `intrinsics` (3x3).  Returns, sorted by file name,
    dict(id, name, qvec (w,x,y,z of world-to-camera), tvec, R (world-to-camera), fx, fy, cx, cy, width, height)."""
    out = []
    for file_name in sorted(os.listdir(camera_dir)):
        if not file_name.endswith(".npz"):
            continue
        data = np.load(os.path.join(camera_dir, file_name))
        pose, K = data["pose"], data["intrinsics"]
        R_w2c = pose[:3, :3].T
        t_w2c = -R_w2c @ pose[:3, 3]
        stem = os.path.splitext(file_name)[0]
        try:
            image_id = int(stem)
        except ValueError:
            image_id = int(os.path.splitext(file_name.split("_")[1])[0])
        cx, cy = K[0, 2], K[1, 2]
        out.append({"id": image_id, "name": stem + ".png", "qvec": _rotmat_to_qvec(R_w2c), "tvec": t_w2c, "R": R_w2c,
                    "fx": float(K[0, 0]), "fy": float(K[1, 1]), "cx": float(cx), "cy": float(cy), "width": int(cx * 2),
                    "height": int(cy * 2)})
    return out


def load_language_feature(language_feature_dir: str, image_name: str, image_height: int, image_width: int, device):
    """Camera.get_language_feature: `<name>_f.npy` (feature map, bilinearly resized to the image) and `<name>_s.npy`
    (segment ids, -1 = unlabelled).  Returns (feature (F, H, W), mask (H', W') bool, seg (H', W') int64)."""
    base = os.path.join(language_feature_dir, image_name)
    fmap = torch.from_numpy(np.load(base + "_f.npy")).to(device)
    if fmap.dim() < 4:
        fmap = fmap[None]
    feat = torch.nn.functional.interpolate(fmap, (image_height, image_width), mode="bilinear", align_corners=False).squeeze(0)
    seg = torch.from_numpy(np.load(base + "_s.npy")).to(device).long()
    return feat, seg != -1, seg
