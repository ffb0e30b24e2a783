"""Next row: on-disk formats.
CPU tier (host code): camera / language-feature readers and the checkpoint tuple against outputs of the REFERENCE's own
read_camera_npz / Camera.get_language_feature / GaussianModel.capture executed on CPU (oracle/make_golden_formats.py);
PLY header + reader against a restated byte-level known-answer file (plyfile, which the reference uses, is absent here).
GPU tier: device row pack / unpack (lsx_rows_pack / lsx_rows_unpack) — save_ply bytes against a numpy restatement of
save_ply's column order, load_ply(save_ply(x)) == x bit-exact for both arena naming schemes."""
import os
import struct

import numpy as np
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GROUPS = ("xyz", "knn_f", "f_dc", "f_rest", "opacity", "scaling", "rotation", "language_feature", "instance_feature")


def _gold():
    return np.load(os.path.join(GOLD_DIR, "formats.npz"))


def test_read_camera_npz_matches_reference(tmp_path):
    from lsx_b200.io_formats import read_camera_npz
    z = _gold()
    for i, fn in enumerate(z["cam_file_names"]):
        np.savez(tmp_path / str(fn), pose=z[f"cam_in_{i}_pose"], intrinsics=z[f"cam_in_{i}_K"])
    (tmp_path / "notes.txt").write_text("ignored")
    cams = sorted(read_camera_npz(str(tmp_path)), key=lambda c: c["id"])
    assert [c["id"] for c in cams] == list(z["cam_ids"]) and [c["name"] for c in cams] == list(z["cam_names"])
    for k, c in enumerate(cams):
        assert np.allclose(c["qvec"], z["cam_qvec"][k], atol=1e-12) and np.allclose(c["tvec"], z["cam_tvec"][k], atol=1e-12)
        assert np.array_equal([c["fx"], c["fy"], c["cx"], c["cy"]], z["cam_params"][k])
        assert [c["width"], c["height"]] == list(z["cam_wh"][k])


def test_load_language_feature_matches_reference(tmp_path):
    from lsx_b200.io_formats import load_language_feature
    z = _gold()
    np.save(tmp_path / "00003_f.npy", z["lf_in_f"])
    np.save(tmp_path / "00003_s.npy", z["lf_in_s"])
    feat, mask, seg = load_language_feature(str(tmp_path), "00003", 48, 72, "cpu")
    assert np.array_equal(feat.numpy(), z["lf_feat"]) and np.array_equal(mask.numpy(), z["lf_mask"])
    assert np.array_equal(seg.numpy(), z["lf_seg"]) and seg.dtype == torch.int64


def test_restore_reads_reference_checkpoint_and_capture_mirrors_it():
    from lsx_b200.io_formats import capture, restore
    z = _gold()
    ref_tuple, iteration = torch.load(os.path.join(GOLD_DIR, "checkpoint_ref.pth"), weights_only=False)
    assert iteration == 1234 and len(ref_tuple) == 20
    st = restore(ref_tuple, "cpu")
    for n in GROUPS:
        assert np.array_equal(st["params"].views[n].numpy(), z[f"ck_{n}"]), n
        assert np.array_equal(st["exp_avg"].views[n].numpy(), z[f"ck_m_{n}"]) and \
            np.array_equal(st["exp_avg_sq"].views[n].numpy(), z[f"ck_v_{n}"]), n
    assert st["step"] == 3 and st["active_sh_degree"] == 2 and st["spatial_lr_scale"] == 2.5
    assert np.array_equal(st["stats"].grad_accum.numpy(), z["ck_grad_accum"]) and np.array_equal(st["stats"].denom.numpy(), z["ck_denom"])
    assert np.array_equal(st["stats"].max_radii2D.numpy(), z["ck_max_radii2D"]) and abs(st["lrs"]["f_dc"] - 3e-3) < 1e-12
    mine = capture(st["params"], st["exp_avg"], st["exp_avg_sq"], st["stats"], st["active_sh_degree"], st["step"], st["lrs"],
                   st["spatial_lr_scale"], st["poses"], st["cam_optimizer_state"], include_feature=True)
    assert len(mine) == len(ref_tuple)
    for a, b in zip(mine, ref_tuple):                       # same positions, shapes and values as GaussianModel.capture
        if isinstance(b, torch.Tensor):
            assert tuple(a.shape) == tuple(b.shape) and torch.equal(a.detach(), b.detach())
    mo, ro = mine[16], ref_tuple[16]
    assert set(mo["state"]) == set(ro["state"]) and [g["name"] for g in mo["param_groups"]] == [g["name"] for g in ro["param_groups"]]
    for i in ro["state"]:
        assert set(ro["state"][i]) == set(mo["state"][i]) and float(mo["state"][i]["step"]) == float(ro["state"][i]["step"])
        assert torch.equal(mo["state"][i]["exp_avg_sq"], ro["state"][i]["exp_avg_sq"])
    for gm_, gr_ in zip(mo["param_groups"], ro["param_groups"]):
        assert set(gr_) <= set(gm_) and gm_["params"] == gr_["params"] and abs(gm_["lr"] - gr_["lr"]) < 1e-12
    short = capture(st["params"], None, None, st["stats"], 2, 0, {}, 1.0, st["poses"], include_feature=False)
    assert len(short) == 18 and restore(short, "cpu")["exp_avg"] is None
    torch.optim.Adam([torch.nn.Parameter(t.clone()) for t in (mine[1], mine[2])], lr=0.0)  # tensors are usable as parameters


def test_capture_round_trips_through_the_reference_restore_sequence():
    """the tuple must survive GaussianModel.restore(mode='train') (gaussian_model.py:139-191): the entries are
    assigned to self._xyz etc. AS THEY ARE (so they must be Parameters that require grad), then training_setup builds a
    9-group Adam and a 1-group camera Adam and both load_state_dict calls must accept the captured dicts."""
    from lsx_b200.io_formats import capture, restore
    ref_tuple, _ = torch.load(os.path.join(GOLD_DIR, "checkpoint_ref.pth"), weights_only=False)
    st = restore(ref_tuple, "cpu")
    for cam_state in (st["cam_optimizer_state"], None):                       # the reference's dict, and our default
        tup = capture(st["params"], st["exp_avg"], st["exp_avg_sq"], st["stats"], st["active_sh_degree"], st["step"], st["lrs"],
                      st["spatial_lr_scale"], st["poses"], cam_state, include_feature=True)
        (_deg, xyz, knn_f, f_dc, f_rest, scaling, rotation, opacity, lang, inst, _mr, _mw, _a, _aa, _d, _da, opt_dict, cam_dict,
         _sls, P) = tup
        named = [("xyz", xyz), ("knn_f", knn_f), ("f_dc", f_dc), ("f_rest", f_rest), ("opacity", opacity), ("scaling", scaling),
                 ("rotation", rotation), ("language_feature", lang), ("instance_feature", inst)]
        for n, t in named:                                                    # restore(): self._xyz = <tuple entry>
            assert isinstance(t, torch.nn.Parameter), n
            assert t.requires_grad == (n != "instance_feature"), n
        optimizer = torch.optim.Adam([{"params": [t], "lr": 1e-3, "name": n} for n, t in named], lr=0.0, eps=1e-15)
        poses = P.detach().clone().requires_grad_(True)
        cam_optimizer = torch.optim.Adam([{"params": [poses], "lr": 1e-4, "name": "pose"}], lr=0.0, eps=1e-15)
        optimizer.load_state_dict(opt_dict)                                   # training_setup + load_state_dict (:186-191)
        cam_optimizer.load_state_dict(cam_dict)
        before = xyz.detach().clone()
        loss = sum((t ** 2).sum() for n, t in named if t.requires_grad and t.numel()) + (poses ** 2).sum()
        loss.backward()
        assert xyz.grad is not None and inst.grad is None
        optimizer.step()
        cam_optimizer.step()
        assert not torch.equal(xyz.detach(), before)                          # resumed training really trains
        assert float(optimizer.state[xyz]["step"]) == st["step"] + 1          # moments were picked up, not re-created


def test_ply_header_and_reader_known_answer(tmp_path):
    """Byte-level KAT of the layout plyfile produces for an all-'f4' vertex element (restated: plyfile is absent)."""
    from lsx_b200.io_formats import ply_attributes, ply_header, read_ply_vertices
    names = ply_attributes(3, 45, 3, 3, True)
    assert names[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] and names[9] == "f_rest_0"
    assert names[54:62] == ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert names[62:] == [f"language_feature_{i}" for i in range(3)] + [f"instance_feature_{i}" for i in range(3)]
    assert len(ply_attributes(3, 45, 3, 3, False)) == 62
    hdr = ply_header(2, ["x", "y", "opacity"])
    assert hdr == b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\n" \
                  b"property float opacity\nend_header\n"
    body = struct.pack("<6f", 1.0, 2.0, 0.5, -1.0, -2.0, 0.25)
    (tmp_path / "a.ply").write_bytes(hdr + body)
    got_names, data = read_ply_vertices(str(tmp_path / "a.ply"))
    assert got_names == ["x", "y", "opacity"] and np.array_equal(data, np.array([[1, 2, .5], [-1, -2, .25]], np.float32))
    # mixed types + comment + a second element, big endian
    hdr2 = b"ply\nformat binary_big_endian 1.0\ncomment hi\nelement vertex 1\nproperty double x\nproperty uchar red\n" \
           b"element face 0\nproperty list uchar int vertex_indices\nend_header\n"
    with pytest.raises(ValueError):
        (tmp_path / "c.ply").write_bytes(hdr2 + struct.pack(">dB", 3.5, 7))
        read_ply_vertices(str(tmp_path / "c.ply"))          # list property in the header -> rejected loudly
    hdr3 = b"ply\nformat binary_big_endian 1.0\ncomment hi\nelement vertex 1\nproperty double x\nproperty uchar red\nend_header\n"
    (tmp_path / "b.ply").write_bytes(hdr3 + struct.pack(">dB", 3.5, 7))
    n3, d3 = read_ply_vertices(str(tmp_path / "b.ply"))
    assert n3 == ["x", "red"] and np.array_equal(d3, np.array([[3.5, 7.0]], np.float32))
    (tmp_path / "t.ply").write_bytes(hdr[:-1] + b"\n" + body[:-4])
    with pytest.raises(ValueError):
        read_ply_vertices(str(tmp_path / "t.ply"))          # truncated body


def _reference_rows(t, include_feature):
    """save_ply's attribute matrix (gaussian_model.py:418-436) restated with numpy on host tensors."""
    P = t["xyz"].shape[0]
    f_dc = t["f_dc"].reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    f_rest = t["f_rest"].reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    cols = [t["xyz"], np.zeros_like(t["xyz"]), f_dc, f_rest, t["opacity"], t["scaling"], t["rotation"]]
    if include_feature:
        cols += [t["language_feature"], t["instance_feature"]]
    return np.concatenate(cols, axis=1).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("include_feature", [True, False])
def test_cuda_save_ply_bytes_and_round_trip(tmp_path, include_feature):
    from lsx_b200.densify import ParamArena
    from lsx_b200.io_formats import load_ply, ply_header, save_ply
    dev = torch.device("cuda:0")
    P, F = 10_007, 16
    widths = {"xyz": 3, "f_dc": 3, "f_rest": 45, "opacity": 1, "scaling": 3, "rotation": 4, "language_feature": F, "instance_feature": 3}
    arena = ParamArena.allocate(P, widths, dev)
    arena.flat.normal_()
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")
    names = save_ply(path, arena, include_feature=include_feature)
    host = {n: v.cpu().numpy() for n, v in arena.views.items()}
    want = ply_header(P, names) + _reference_rows(host, include_feature).tobytes()
    assert open(path, "rb").read() == want
    back = load_ply(path, dev)
    for n in widths:
        if n in ("language_feature", "instance_feature") and not include_feature:
            assert back.views[n].shape == (P, 0)
        else:
            assert torch.equal(back.views[n], arena.views[n]), n
    with pytest.raises(AssertionError):
        load_ply(path, dev, max_sh_degree=2)


@pytest.mark.gpu
def test_cuda_save_ply_from_rasterizer_named_arena(tmp_path):
    """the multi-view arena (means3D / sh / scales / rotations ...) writes the same file as the reference-named one"""
    from lsx_b200.densify import ParamArena
    from lsx_b200.io_formats import save_ply
    from lsx_b200.multiview import GradArena
    dev = torch.device("cuda:0")
    P, F = 3001, 3
    a = GradArena.allocate(P, 16, F, 3, dev)
    a.flat.normal_()
    b = ParamArena.allocate(P, {"xyz": 3, "f_dc": 3, "f_rest": 45, "opacity": 1, "scaling": 3, "rotation": 4, "language_feature": F,
                                "instance_feature": 3}, dev)
    b.views["xyz"].copy_(a.views["means3D"])
    b.views["f_dc"].copy_(a.views["sh"][:, :3])
    b.views["f_rest"].copy_(a.views["sh"][:, 3:])
    for x, y in (("opacity", "opacity"), ("scaling", "scales"), ("rotation", "rotations"), ("language_feature", "language_feature"),
                 ("instance_feature", "instance_feature")):
        b.views[x].copy_(a.views[y])
    save_ply(str(tmp_path / "a.ply"), a, include_feature=True)
    save_ply(str(tmp_path / "b.ply"), b, include_feature=True)
    assert open(tmp_path / "a.ply", "rb").read() == open(tmp_path / "b.ply", "rb").read()
    with pytest.raises(RuntimeError):
        save_ply(str(tmp_path / "c.ply"), ParamArena.allocate(4, {"xyz": 3, "f_dc": 3, "f_rest": 45, "opacity": 1, "scaling": 3,
                                                                   "rotation": 4}, "cpu"))
