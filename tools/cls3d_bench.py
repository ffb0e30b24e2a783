"""3-D neighbourhood regulariser (loss_cls_3d) fwd+bwd: fused tiled k-NN kernels vs the reference's cdist + topk formulation."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.loss import loss_cls_3d  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def torch_formulation(xyz, pred, k, lam, samples):
    """the reference's formulation (loss_utils.py:165-186) with torch ops on the GPU: full distance matrix + topk"""
    lo, hi = pred.min(), pred.max()
    if hi > lo:                                                   # host read, like the reference
        pred = (pred - lo) / (hi - lo)
    dists = torch.cdist(xyz[samples], xyz)
    nbr = dists.topk(k, largest=False).indices
    own = pred[samples].unsqueeze(1)
    kl = own * (torch.log(own + 1e-10) - torch.log(pred[nbr] + 1e-10))
    return lam * kl.abs().mean(), nbr


for N in (100_000, 500_000, 2_000_000):
    g = torch.Generator().manual_seed(N)
    xyz = (torch.rand(N, 3, generator=g) * 8 - 4).cuda()
    pred = torch.randn(N, 3, generator=g).cuda().requires_grad_(True)
    samples = torch.randperm(N, generator=g)[:800].cuda()

    def fused():
        pred.grad = None
        loss_cls_3d(xyz, pred, 5, 2.0, 2_000_000, 800, sample_indices=samples).backward()

    def ref():
        pred.grad = None
        torch_formulation(xyz, pred, 5, 2.0, samples)[0].backward()

    from lsx_b200.loss import knn_tree
    tree = knn_tree(xyz)

    def fused_tree():
        pred.grad = None
        loss_cls_3d(xyz, pred, 5, 2.0, 2_000_000, 800, sample_indices=samples, tree=tree).backward()

    t_build, t_tree = timeit(lambda: knn_tree(xyz)), timeit(fused_tree)
    tf, tt = timeit(fused), timeit(ref)
    lf, nf = loss_cls_3d(xyz, pred, 5, 2.0, 2_000_000, 800, sample_indices=samples, return_neighbors=True)
    lt, nt = torch_formulation(xyz, pred, 5, 2.0, samples)
    same = float((nf.long().sort(dim=1).values == nt.sort(dim=1).values).float().mean())
    print(json.dumps({"op": "loss_cls_3d fwd+bwd", "N": N, "samples": 800, "k": 5, "fused_ms": round(tf, 4),
                      "torch_ops_ms": round(tt, 4), "fused_with_prebuilt_tree_ms": round(t_tree, 4), "tree_build_ms": round(t_build, 4), "speedup": round(tt / tf, 2), "loss_fused": float(lf), "loss_torch_ops": float(lt),
                      "neighbour_entries_equal": round(same, 6),
                      "distance_evals_per_s": round(800 * N / tf * 1e3 / 1e12, 3), "unit": "T evals/s"}), flush=True)
