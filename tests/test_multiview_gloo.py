"""CPU tier, world_size 2 over gloo: view sharding, flat gradient arena all-reduce and densification statistics
give exactly what single-process accumulation over the same views gives."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness  # noqa: F401  (sys.path)
from lsx_b200.multiview import DensifyStats, GradArena, multiview_step, shard_views

P, M, F, FI, V = 257, 4, 16, 3, 5


def fake_view(v):
    """Deterministic stand-in for (forward, backward) of view v with the native tuple's names and shapes."""
    g = torch.Generator().manual_seed(100 + v)
    r = lambda *s: torch.randn(*s, generator=g)
    radii = (torch.rand(P, generator=g) > 0.3).int() * torch.randint(1, 40, (P,), generator=g, dtype=torch.int32)
    fwd = {"radii": radii, "out_observe": torch.randint(0, 3, (P,), generator=g, dtype=torch.int32)}
    bwd = {"means2D": r(P, 3), "means2D_abs": r(P, 3).abs(), "colors": r(P, 3), "language_feature": r(P, F),
           "instance_feature": r(P, FI), "opacity": r(P, 1), "means3D": r(P, 3), "cov3D": r(P, 6), "sh": r(P, M, 3),
           "scales": r(P, 3), "rotations": r(P, 4), "all_map": r(P, 5)}
    return fwd, bwd


def single_process_reference():
    arena, stats = GradArena.allocate(P, M, F, FI, "cpu"), DensifyStats.allocate(P, "cpu")
    n = multiview_step(fake_view, V, arena, stats)
    assert n == V
    return arena, stats


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arena, stats = GradArena.allocate(P, M, F, FI, "cpu"), DensifyStats.allocate(P, "cpu")
    n_local = multiview_step(fake_view, V, arena, stats)
    torch.save((rank, n_local, arena.flat.clone(), stats.grad_accum.clone(), stats.grad_accum_abs.clone(),
                stats.denom.clone(), stats.max_radii2D.clone()), os.path.join(q, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_views_partition():
    for n in (0, 1, 5, 49, 64):
        for w in (1, 2, 4, 8):
            parts = [shard_views(n, w, r) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_arena_layout():
    a = GradArena.allocate(P, M, F, FI, "cpu")
    assert a.views["sh"].shape == (P, 3 * M) and a.views["language_feature"].shape == (P, F)
    for v in a.views.values():
        assert v.data_ptr() % 256 == a.flat.data_ptr() % 256  # 256-B aligned groups
    a.views["opacity"].fill_(2.0)
    assert float(a.flat.sum()) == 2.0 * P


def test_two_rank_allreduce_equals_single_process_accumulation():
    ref_arena, ref_stats = single_process_reference()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    import tempfile
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as q:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
            assert p.exitcode == 0
        results = [torch.load(os.path.join(q, f"rank{r}.pt")) for r in range(2)]
    assert sorted(r[1] for r in results) == [2, 3]          # 5 views over 2 ranks
    for _, _, flat, ga, gaa, den, mr in results:
        assert torch.allclose(flat, ref_arena.flat, rtol=1e-6, atol=1e-6)   # identical on every rank
        assert torch.allclose(ga, ref_stats.grad_accum, rtol=1e-6, atol=1e-6)
        assert torch.allclose(gaa, ref_stats.grad_accum_abs, rtol=1e-6, atol=1e-6)
        assert torch.equal(den, ref_stats.denom) and torch.equal(mr, ref_stats.max_radii2D)
    assert torch.equal(results[0][2], results[1][2])          # bitwise identical across ranks
