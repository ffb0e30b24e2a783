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


STEPS = 3


def fake_view_of_step(step):
    return lambda v: fake_view(1000 * step + v)


def single_process_reference():
    arena, stats = GradArena.allocate(P, M, F, FI, "cpu"), DensifyStats.allocate(P, "cpu")
    n = multiview_step(fake_view, V, arena, stats)
    assert n == V
    return arena, stats


def single_process_steps():
    """Persistent statistics after STEPS steps of V views each, accumulated by one process (the reference's semantics)."""
    arena, stats = GradArena.allocate(P, M, F, FI, "cpu"), DensifyStats.allocate(P, "cpu")
    for s in range(STEPS):
        multiview_step(fake_view_of_step(s), V, arena, stats)
    return arena, stats


def _worker_steps(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arena, stats = GradArena.allocate(P, M, F, FI, "cpu", extra={"pose": (V, 7)}), DensifyStats.allocate(P, "cpu")
    for s in range(STEPS):
        multiview_step(fake_view_of_step(s), V, arena, stats)
    # piecewise asynchronous reduction of the same arena must give what the single collective gives
    a2 = GradArena.allocate(P, M, F, FI, "cpu", extra={"pose": (V, 7)})
    a2.flat.copy_(torch.arange(a2.flat.numel(), dtype=torch.float32) * (rank + 1))
    pend = a2.all_reduce_spans([["sh"], ["means3D"], ["opacity", "scales", "rotations"],
                                ["language_feature", "instance_feature", "pose"]])
    pend.wait()
    expect = torch.arange(a2.flat.numel(), dtype=torch.float32) * sum(r + 1 for r in range(world))
    torch.save((rank, arena.flat.clone(), stats.grad_accum.clone(), stats.grad_accum_abs.clone(), stats.denom.clone(),
                stats.max_radii2D.clone(), bool(torch.equal(a2.flat, expect))), os.path.join(q, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


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


def test_statistics_stay_cumulative_over_steps():
    """the persistent statistics are cumulative; only the per-step delta may be all-reduced.  Three steps on two
    ranks must equal three steps of single-process accumulation (they grew by world_size per step before the fix)."""
    _, ref_stats = single_process_steps()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    import tempfile
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as q:
        procs = [ctx.Process(target=_worker_steps, args=(r, 2, port, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=180)
            assert p.exitcode == 0
        results = [torch.load(os.path.join(q, f"rank{r}.pt")) for r in range(2)]
    for _, _, ga, gaa, den, mr, spans_ok in results:
        assert torch.allclose(ga, ref_stats.grad_accum, rtol=1e-5, atol=1e-6)
        assert torch.allclose(gaa, ref_stats.grad_accum_abs, rtol=1e-5, atol=1e-6)
        assert torch.equal(den, ref_stats.denom) and torch.equal(mr, ref_stats.max_radii2D)
        assert spans_ok


def test_arena_spans_and_extra_groups():
    a = GradArena.allocate(P, M, F, FI, "cpu", extra={"pose": (V, 7)})
    assert a.views["pose"].shape == (V, 7) and "all_map" not in a.views
    whole = a.span(list(a.offsets))
    assert whole.data_ptr() == a.flat.data_ptr() and whole.numel() == a.flat.numel()
    pieces = [a.span([n]) for n in a.offsets]
    assert sum(p.numel() for p in pieces) == a.flat.numel()            # the spans tile the arena, padding included
    for x, y in zip(pieces[:-1], pieces[1:]):
        assert x.data_ptr() + 4 * x.numel() == y.data_ptr()
    try:
        a.span(["means3D", "opacity"])
        raise AssertionError("non-adjacent groups must be rejected")
    except ValueError:
        pass
