"""3-D neighbourhood regulariser (loss_cls_3d, field_construction/utils/loss_utils.py:158-186).

Golden vectors: tests/golden/cls3d.npz, produced by the reference's OWN loss_cls_3d + autograd on CPU
(oracle/make_golden_cls3d.py).  CPU tier: the restatement (oracle/cls3d_oracle.py) against them.  GPU tier: the CUDA
operator through the C ABI against them (same seeds -> same torch.randperm draws -> same rows) and, at a larger size,
against the restatement.  Tolerances: loss 1e-5 relative, neighbour sets exact, gradient 1e-4 of the tensor's scale PLUS,
element by element, 3x the reference's own float32 error against float64 (the restatement run in double): when the minimum
element is somebody's neighbour, log(0 + 1e-10) puts terms of ~1e7 on the two gradient paths of that element (directly and
through `predictions.min()`), which cancel; float32 autograd leaves the rounding noise of that cancellation there (case c:
0.0 instead of 0.0145), the CUDA path cancels in double.
"""
import os

import numpy as np
import pytest
import torch

import harness as hz  # noqa: F401  (sys.path)
from oracle import cls3d_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cls3d.npz")
CASES = ["a", "b", "c", "d"]
LOSS_TOL, GRAD_TOL = 1e-5, 1e-4


def _case(name):
    z = np.load(GOLD)
    c = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "_")}
    k, lam, max_points, sample_size, seed, up = c["args"]
    c.update(k=int(k), lam=float(lam), max_points=int(max_points), sample_size=int(sample_size), seed=int(seed), up=float(up))
    return c


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _grad_f64(c):
    """the restatement in double on the case's rows: what the reference's float32 gradient approximates"""
    pred = torch.from_numpy(c["pred"]).double().requires_grad_(True)
    down = torch.from_numpy(c["down"]) if c["down"].size else None
    loss, _ = orc.loss_cls_3d(torch.from_numpy(c["xyz"]).double(), pred, c["k"], c["lam"], torch.from_numpy(c["samples"]), down)
    (loss * c["up"]).backward()
    return pred.grad


def _assert_grad(got, g_ref, g64):
    """|got - ref| <= 1e-4 max|ref| + 3 |ref - f64|, element by element (median-scale reference: the extremum rows can be huge)"""
    if float(g_ref.abs().max()) == 0:
        assert float(got.abs().max()) == 0.0
        return
    scale = float(g_ref.abs().max())
    bound = GRAD_TOL * scale + 3 * (g_ref.double() - g64).abs()
    bad = (got.double() - g_ref.double()).abs() > bound
    assert not bool(bad.any()), (int(bad.sum()), got[bad][:4], g_ref[bad][:4])


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_vectors(name):
    c = _case(name)
    pred = torch.from_numpy(c["pred"]).requires_grad_(True)
    down = torch.from_numpy(c["down"]) if c["down"].size else None
    loss, nbr = orc.loss_cls_3d(torch.from_numpy(c["xyz"]), pred, c["k"], c["lam"], torch.from_numpy(c["samples"]), down)
    (loss * c["up"]).backward()
    assert torch.equal(nbr.sort(dim=1).values, torch.from_numpy(c["nbr"]).sort(dim=1).values)
    assert abs(float(loss.detach()) - float(c["loss"])) <= LOSS_TOL * max(abs(float(c["loss"])), 1e-6)
    g_ref = torch.from_numpy(c["g_pred"])
    if float(g_ref.abs().max()) > 0:
        assert _rel(pred.grad, g_ref) < GRAD_TOL
    else:
        assert float(pred.grad.abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_vectors(name):
    from lsx_b200.loss import loss_cls_3d
    c = _case(name)
    xyz = torch.from_numpy(c["xyz"]).cuda()
    pred = torch.from_numpy(c["pred"]).cuda().requires_grad_(True)
    torch.manual_seed(c["seed"])                       # same CPU-generator draws as the reference run that made the vectors
    loss, nbr = loss_cls_3d(xyz, pred, c["k"], c["lam"], c["max_points"], c["sample_size"], return_neighbors=True)
    (loss * c["up"]).backward()
    assert torch.equal(nbr.cpu().long().sort(dim=1).values, torch.from_numpy(c["nbr"]).sort(dim=1).values)
    assert abs(float(loss.detach()) - float(c["loss"])) <= LOSS_TOL * max(abs(float(c["loss"])), 1e-6)
    assert pred.grad.shape == pred.shape
    _assert_grad(pred.grad.cpu(), torch.from_numpy(c["g_pred"]), _grad_f64(c))


@pytest.mark.gpu
@pytest.mark.parametrize("N,C,k,S", [(100_000, 3, 5, 800), (70_001, 16, 8, 333), (9, 2, 8, 20)])
def test_cuda_against_restatement(N, C, k, S):
    from lsx_b200.loss import loss_cls_3d
    g = torch.Generator().manual_seed(N)
    xyz = torch.rand(N, 3, generator=g) * 4 - 2
    pred = torch.randn(N, C, generator=g)
    samples = torch.randperm(N, generator=g)[:S]
    nbr_want = orc.knn_exact(xyz, samples, k)                          # float32 distances, like the kernel
    p_cpu = pred.double().requires_grad_(True)
    want, _ = orc.loss_cls_3d(xyz.double(), p_cpu, k, 2.0, samples, neighbors=nbr_want)
    want.backward()
    p_gpu = pred.cuda().requires_grad_(True)
    got, nbr = loss_cls_3d(xyz.cuda(), p_gpu, k, 2.0, sample_indices=samples, return_neighbors=True)
    got.backward()
    assert torch.equal(nbr.cpu().long(), nbr_want)          # same order too: both break ties towards the lower index
    assert abs(float(got) - float(want)) <= LOSS_TOL * abs(float(want))
    assert _rel(p_gpu.grad.cpu().double(), p_cpu.grad) < GRAD_TOL


@pytest.mark.gpu
def test_cuda_errors():
    from lsx_b200.loss import loss_cls_3d
    xyz, pred = torch.rand(50, 3), torch.rand(50, 3)
    with pytest.raises(RuntimeError):
        loss_cls_3d(xyz, pred)                                              # CPU tensors
    with pytest.raises(RuntimeError):
        loss_cls_3d(xyz.cuda(), pred.cuda(), k=9)                           # k > 8
    with pytest.raises(RuntimeError):
        loss_cls_3d(xyz.cuda()[:4], pred.cuda()[:4], k=5)                   # k > N (torch.topk raises in the reference)
    with pytest.raises(RuntimeError):
        loss_cls_3d(xyz.cuda(), pred.cuda()[:10])


@pytest.mark.gpu
@pytest.mark.parametrize("N,C,k,S,dup", [(500_000, 3, 5, 800, False), (70_000, 16, 8, 300, False), (1500, 3, 5, 64, True),
                                         (20, 2, 3, 20, False), (33, 1, 1, 7, True), (40_000, 3, 5, 24, "all"),
                                         (2_000_000, 3, 5, 100, False)])
def test_tree_search_returns_the_brute_force_neighbours(N, C, k, S, dup):
    """loss_cls_3d(..., tree=knn_tree(xyz)): the neighbour search through the Morton / box hierarchy must return exactly what the
    scan of all points returns — same indices in the same order (ties towards the lower index: `dup` plants coincident points),
    hence the same loss and gradient bits — at sizes with one, two and three populated hierarchy levels, and on degenerate data
    (all points coincident) where nothing can be pruned."""
    from lsx_b200.loss import knn_tree, loss_cls_3d
    g = torch.Generator().manual_seed(N + k)
    xyz = (torch.rand(N, 3, generator=g) * 6 - 3)
    if dup == "all":
        xyz[:] = xyz[0].clone()     # every point coincides: all 1250 leaves tie with the bound, the leaf queue (1024) overflows
    elif dup:
        xyz[N // 2:] = xyz[:N - N // 2].clone()                      # every point of the second half coincides with one of the first
    xyz = xyz.cuda()
    pred = torch.randn(N, C, generator=g).cuda()
    samples = torch.randint(0, N, (S,), generator=g)
    tree = knn_tree(xyz)
    out = {}
    for name, t in (("scan", None), ("tree", tree)):
        p = pred.clone().requires_grad_(True)
        loss, nbr = loss_cls_3d(xyz, p, k, 2.0, 10_000_000, S, sample_indices=samples, return_neighbors=True, tree=t)
        loss.backward()
        out[name] = (loss.detach(), nbr, p.grad)
    assert torch.equal(out["scan"][1], out["tree"][1])
    assert torch.equal(out["scan"][0], out["tree"][0])
    if dup == "all":   # every sample shares the same k neighbour rows: the backward's atomics onto them reorder between runs
        scale = float(out["scan"][2].abs().max())
        assert float((out["scan"][2] - out["tree"][2]).abs().max()) <= 1e-5 * scale
    else:
        assert torch.equal(out["scan"][2], out["tree"][2])
    with pytest.raises(RuntimeError, match="another point count"):
        loss_cls_3d(xyz[:-1].contiguous(), pred[:-1].contiguous(), k, 2.0, 10_000_000, S, sample_indices=samples % (N - 1), tree=tree)
