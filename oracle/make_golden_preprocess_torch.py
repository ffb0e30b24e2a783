"""Records tests/golden/preprocess_torch.npz from the reference's OWN Python preprocess functions executed on CPU:
eval_sh (field_construction/utils/sh_utils.py, device-agnostic) composed as in gaussian_renderer/__init__.py:114-119, and
build_scaling_rotation / strip_symmetric (utils/general_utils.py) composed as gaussian_model.py:47-51 — their hard-coded
device="cuda" is redirected to the CPU for the duration of the call; nothing else is changed.
    python oracle/make_golden_preprocess_torch.py      (needs /root/reference; run in the build container)"""
import importlib.util
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LSX_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "langscene-x_b200"))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from lsx_b200.synthetic import make_camera, make_scene
    sh_utils = load(os.path.join(REF, "field_construction", "utils", "sh_utils.py"), "ref_sh_utils")
    src = open(os.path.join(REF, "field_construction", "utils", "general_utils.py")).read()
    # general_utils imports PIL / numpy helpers at module level that are irrelevant here: execute only the three functions
    ns = {"torch": torch}
    start = src.index("def strip_lowerdiag")
    end = src.index("def safe_state")
    code = src[start:end].replace('device="cuda"', 'device="cpu"').replace("device='cuda'", "device='cpu'")
    exec(compile(code, "general_utils.py[strip_lowerdiag..build_scaling_rotation]", "exec"), ns)
    out = {}
    for tag, P, seed, deg in (("a", 600, 3, 3), ("b", 300, 4, 1), ("c", 200, 5, 0), ("d", 300, 6, 2)):
        scene = make_scene(P, 320, 240, F=3, seed=seed)
        cam = make_camera(320, 240, yaw_deg=7.0)
        shs_view = scene.shs.transpose(1, 2).view(-1, 3, 16)
        d = scene.means3D - cam.campos.repeat(P, 1)
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(sh_utils.eval_sh(deg, shs_view, d) + 0.5, 0.0)         # gaussian_renderer/__init__.py:114-119
        rot = scene.rotations * (0.5 + torch.rand(P, 1, generator=torch.Generator().manual_seed(seed)))   # un-normalised
        mod = 1.0 if tag != "d" else 0.7
        L = ns["build_scaling_rotation"](mod * scene.scales, rot)                    # gaussian_model.py:47-51
        cov = ns["strip_symmetric"](L @ L.transpose(1, 2))
        out.update({f"{tag}_deg": deg, f"{tag}_mod": mod, f"{tag}_shs": scene.shs.numpy(), f"{tag}_xyz": scene.means3D.numpy(),
                    f"{tag}_campos": cam.campos.numpy(), f"{tag}_scales": scene.scales.numpy(), f"{tag}_rot": rot.numpy(),
                    f"{tag}_rgb": rgb.numpy(), f"{tag}_cov": cov.numpy()})
    path = os.path.join(REPO, "tests", "golden", "preprocess_torch.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
