"""Next row (per-Gaussian half): activations + plane normal + all_map.
CPU tier: the torch restatement against vectors recorded from the REFERENCE's own GaussianModel methods + autograd
(oracle/make_golden_gaussian_head.py).  GPU tier: the fused CUDA kernels (C ABI) against the same vectors and, at 1 M
Gaussians, against the restatement.  Tolerances: forward 1e-5, gradient 1e-4 (tensor-scale relative)."""
import os

import numpy as np
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)
from oracle import gaussian_head_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gaussian_head.npz")
OUT = ("scales", "rotations", "opac", "all_map")
RAW = ("xyz", "scaling", "rotation", "opacity")


def _cases():
    z = np.load(GOLD)
    for name in ("a", "b"):
        yield name, {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "_")}


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


def _run(fn, c, device):
    raw = [c[k].to(device).clone().requires_grad_(True) for k in RAW]
    outs = fn(*raw, c["view"].to(device), c["campos"].to(device))
    loss = sum((o * c["up_" + k].to(device)).sum() for o, k in zip(outs, OUT))
    loss.backward()
    return outs, [r.grad for r in raw]


@pytest.mark.parametrize("name,c", list(_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_restatement_matches_reference_vectors(name, c):
    outs, grads = _run(orc.gaussian_head, c, "cpu")
    for o, k in zip(outs, OUT):
        assert _rel(o.detach(), c[k]) < 1e-6, k
    for g, k in zip(grads, RAW):
        assert _rel(g, c["g_" + k]) < 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize("name,c", list(_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_cuda_matches_reference_vectors(name, c):
    from lsx_b200.render_utils import gaussian_head
    outs, grads = _run(gaussian_head, c, "cuda:0")
    for o, k in zip(outs, OUT):
        assert _rel(o.detach().cpu(), c[k]) < 1e-5, k
    for g, k in zip(grads, RAW):
        assert _rel(g.cpu(), c["g_" + k]) < 1e-4, k


@pytest.mark.gpu
def test_cuda_matches_restatement_at_1M():
    from lsx_b200.render_utils import gaussian_head
    from lsx_b200.synthetic import make_camera
    P = 1_000_000
    g = torch.Generator().manual_seed(2)
    c = {"xyz": torch.randn(P, 3, generator=g) * 3 + torch.tensor([0.0, 0.0, 4.0]), "scaling": torch.randn(P, 3, generator=g) - 4,
         "rotation": torch.randn(P, 4, generator=g), "opacity": torch.randn(P, 1, generator=g)}
    cam = make_camera(1920, 1080, yaw_deg=7.0)
    c["view"], c["campos"] = cam.viewmatrix, cam.campos
    for k, shp in zip(OUT, ((P, 3), (P, 4), (P, 1), (P, 5))):
        c["up_" + k] = torch.randn(*shp, generator=g)
    want_o, want_g = _run(orc.gaussian_head, {k: (v.double() if v.dtype == torch.float32 else v) for k, v in c.items()}, "cpu")
    got_o, got_g = _run(gaussian_head, c, "cuda:0")
    for a, b, k in zip(got_o, want_o, OUT):
        assert _rel(a.detach().cpu(), b.detach()) < 1e-5, k
    for a, b, k in zip(got_g, want_g, RAW):
        assert _rel(a.cpu(), b) < 1e-4, k
    with pytest.raises(RuntimeError):
        gaussian_head(c["xyz"], c["scaling"], c["rotation"], c["opacity"], c["view"], c["campos"])   # CPU tensors
