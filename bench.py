#!/usr/bin/env python
"""bench.py — fwd+bwd throughput of the rasterizer hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl new|reference|reference-cpu] [--config C3|C4-loop|C5-stress]

(--config C4-loop / C5-stress: BASELINE configs 4 and 5 — the whole optimisation iteration, sharded; see bench_loop.py.)

A "step" is one optimisation step's gradient on every GPU: forward + backward of `--views-per-gpu` (default 8)
1920x1080 views of the 1M-Gaussian synthetic scene (16-d language feature + instance feature +
normal/alpha/distance map + plane depth, SH degree 3), their per-Gaussian parameter gradients summed into one
flat arena.  With N > 1 (torchrun) every rank renders its own views (weak scaling: 8 views per GPU, i.e. the
64-view batch of BASELINE config 5 at N = 8) and the arena is summed with ONE NCCL all-reduce inside the timed
region.  MPix/s = N * views * W * H / time; `ms_per_iter` = time per single-view fwd+bwd (the metric's ms/iter).

Printed JSON (rank 0, one line): see the task contract.  Extras: `stages_ms` (per-stage device times from the
library's event hooks), `stats` (P_vis, R, S), `peaks` (measured FP32 / EX2 / RED.ADD rates), `ref_cuda`
(the reference CUDA rasterizer timed in the same process when oracle/_ref is present).

--impl reference       the UNMODIFIED reference CUDA rasterizer (oracle/_ref/ref_rast_f16.so, recompiled for
                       sm_100a) driven through the same protocol — the baseline BASELINE.json's target names.
--impl reference-cpu   the CPU restatement (oracle/) on all host threads on a bounded sample.
"""
import argparse
import gc
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(REPO, "langscene-x_b200"), REPO, os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "fwd+bwd MPix/s (1M Gaussians, 1920x1080, 16-d language feature + normal + depth)"


def ncu_record():
    """dram bytes / issue-slot utilisation of the render kernels from the committed `ncu --set full` capture, stamped with
    the sha1 of the kernel sources they were captured from (profiles/ncu_render_kernels.json, written by
    tools/ncu_summary.py).  A capture of OTHER sources is reported as stale: traffic = null, never a stale number."""
    import hashlib
    path = os.path.join(REPO, "profiles", "ncu_render_kernels.json")
    if not os.path.exists(path):
        return None, "no committed capture (profiles/ncu_render_kernels.json)"
    with open(path) as f:
        rec = json.load(f)
    h = hashlib.sha1()
    for name in ("render_fwd.cu", "render_bwd.cu", "tile_stage.cuh"):
        with open(os.path.join(REPO, "langscene-x_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    if rec.get("sources_sha1") != h.hexdigest():
        return None, f"stale: captured from sources {str(rec.get('sources_sha1'))[:12]}, tree has {h.hexdigest()[:12]}"
    return rec, "profiles/" + rec.get("capture", "ncu_render_kernels.json")


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """SM clock / throttle reasons sampled every 100 ms DURING the timed regions, in-process through NVML
    (spawning `nvidia-smi -lms` costs a driver initialisation that lands inside short timed regions and was
    measured to stall kernel launches for 50-200 ms).  Started before the warm-up so that NVML is initialised
    outside the timed region."""

    def __init__(self, index):
        self.rows, self.ok, self._stop = [], False, threading.Event()
        self.call_ms, self.period = 0.0, float(os.environ.get("LSX_BENCH_SAMPLE_PERIOD", "0.1"))
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = str(e)[:100]

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                t0 = time.perf_counter()
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                t1 = time.perf_counter()
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                t2 = time.perf_counter()
                self.rows.append((sm, reasons))
                self.call_ms = max(self.call_ms, (t1 - t0) * 1e3, (t2 - t1) * 1e3)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self.thread.join(timeout=1.0)
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for _, r in self.rows:
            bits |= r
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": [n for b, n in names.items() if bits & b], "samples": len(sm),
                "nvml_call_ms_max": round(self.call_ms, 3)}


def build_views(cfg_name, device, rank=0, world=1, views=1, seed=0):
    """The shared scene plus this rank's cameras: global view v = rank + world*i of world*views views on the arc
    yaw = -15..+15 degrees; a single view overall is the yaw-0 camera."""
    from lsx_b200.synthetic import CONFIGS, make_all_map, make_camera, make_scene, make_upstream_grads
    import harness as hz
    c = CONFIGS[cfg_name]
    scene = make_scene(c["P"], c["W"], c["H"], F=c["F"], seed=seed, s_med=c["s_med"]).to(device)
    grads = make_upstream_grads(c["W"], c["H"], c["F"], device=device)
    bg = torch.zeros(3, device=device)
    total = world * views
    out = []
    for i in range(views):
        v = rank + world * i      # round-robin over the arc: every rank gets the same spread of view costs
        yaw = 0.0 if total == 1 else (-15.0 + 30.0 * v / (total - 1))
        cam = make_camera(c["W"], c["H"], yaw_deg=yaw).to(device)
        am = make_all_map(scene, cam)
        out.append({"cam": cam, "am": am, "yaw": yaw, "fargs": hz.native_forward_args(scene, cam, bg, c["F"], all_map=am)})
    return c, scene, grads, bg, out


def build_case(cfg_name, device, view_yaw=0.0, seed=0):
    """Single-view case (tools/ use this)."""
    from lsx_b200.synthetic import CONFIGS, make_all_map, make_camera, make_scene, make_upstream_grads
    import harness as hz
    c = CONFIGS[cfg_name]
    scene = make_scene(c["P"], c["W"], c["H"], F=c["F"], seed=seed, s_med=c["s_med"]).to(device)
    cam = make_camera(c["W"], c["H"], yaw_deg=view_yaw).to(device)
    grads = make_upstream_grads(c["W"], c["H"], c["F"], device=device)
    bg = torch.zeros(3, device=device)
    am = make_all_map(scene, cam)
    fargs = hz.native_forward_args(scene, cam, bg, c["F"], all_map=am)
    return c, scene, cam, grads, bg, am, fargs


def time_loop(step_fn, steps, warmup, world, detail=None):
    """W warm-ups, then exactly K steps between barrier+synchronize pairs, CUDA events, max over ranks.
    `detail` (dict) receives the per-step median from one event per step (diagnostic only)."""
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    # The step has a host round trip (num_rendered), so a stop-the-world pass of CPython's cyclic collector over the
    # torch-sized heap (measured: 100-175 ms, once every few hundred steps) would land inside a 100 ms timed region.
    # Collect now, keep the collector off while timing (reference-counted frees, i.e. all tensor frees, still happen).
    gc.collect()
    gc.disable()
    ev[0].record()
    for i in range(steps):
        step_fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    gc.enable()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[steps])], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if detail is not None:
        per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
        detail["median_ms"] = per[len(per) // 2]
        detail["max_ms"] = per[-1]
    return float(ms.item())


def time_collective(flat, world, reps=10):
    """The arena all-reduce alone (no compute around it), CUDA events, max over ranks: what the step pays when nothing
    overlaps it.  busbw = algbw * 2 (N - 1) / N, NCCL's convention."""
    if world <= 1:
        return None
    buf = torch.empty_like(flat)
    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=flat.device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    nbytes = flat.numel() * 4
    alg = nbytes / (ms * 1e-3) / 1e9
    return {"what": "ncclAllReduce(sum, fp32) of the flat gradient arena, back to back, nothing else on the GPU", "bytes": nbytes,
            "ms": ms, "algbw_GBs": alg, "busbw_GBs": alg * 2 * (world - 1) / world}


def native_stepper(mod, fargs, grads, arena=None, world=1):
    """Single-view device-resident step (tools/): raw `_C`-level forward + backward."""
    import harness as hz

    def step():
        fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*fargs)))
        bwd = dict(zip(hz.BWD_NAMES, mod.rasterize_gaussians_backward(*hz.native_backward_args(fargs, fwd, grads))))
        return fwd, bwd
    return step


def batch_stepper(mod, views, grads, arena, world):
    """Device-resident step: raw `_C`-level forward + backward of every local view, gradients summed into the
    flat arena, one all-reduce when world > 1."""
    import harness as hz
    from lsx_b200.multiview import BWD_TO_GROUP

    fused = hasattr(mod, "distCUDA2")  # lsx_b200.ops: its backward accumulates straight into the arena

    def step():
        if not fused:
            arena.zero_()
        for i, vw in enumerate(views):
            fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*vw["fargs"])))
            bargs = hz.native_backward_args(vw["fargs"], fwd, grads)
            if fused:
                # the first view of a step overwrites the arena (every row is written), the others add to it
                bwd = dict(zip(hz.BWD_NAMES, mod.rasterize_gaussians_backward(*bargs, grad_buffers=arena.grad_buffers(),
                                                                              accumulate=i > 0)))
            else:
                bwd = dict(zip(hz.BWD_NAMES, mod.rasterize_gaussians_backward(*bargs)))
                arena.accumulate({g: bwd[k] for k, g in BWD_TO_GROUP.items()})
        if world > 1:
            arena.all_reduce()
        return fwd, bwd
    return step


def _pin_views(views, grads):
    cams = [torch.cat([vw["cam"].viewmatrix.flatten(), vw["cam"].projmatrix.flatten(), vw["cam"].campos.flatten()]).cpu().pin_memory()
            for vw in views]
    gcol = grads["color"].cpu().pin_memory()
    h2d = sum(c.numel() * 4 for c in cams) + len(views) * gcol.numel() * 4
    return cams, gcol, h2d


class HostFeeder:
    """Per-view host inputs (camera matrices + colour-supervision gradient) in pinned memory, copied to the device on
    a side stream into two alternating buffers: while view i is computed, the copy of the NEXT view — cyclically, i.e. the
    first view of the next step while the last view of this one runs — is in flight.  Every copy is issued inside the timed
    region; only the very first one (before the first step) is not overlapped."""

    def __init__(self, views, grads):
        dev = grads["color"].device
        self.cams, self.gcol, self.h2d_bytes = _pin_views(views, grads)
        self.n = len(views)
        self.stream = torch.cuda.Stream(device=dev)
        self.buf = [(torch.empty_like(self.cams[0], device=dev), torch.empty_like(grads["color"])) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        for e in self.free:
            e.record()
        self.issued = 0        # copies issued so far; copy k goes to buffer k & 1 and carries view k % n
        self.consumed = 0

    def _issue(self):
        slot, i = self.issued & 1, self.issued % self.n
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])
            self.buf[slot][0].copy_(self.cams[i], non_blocking=True)
            self.buf[slot][1].copy_(self.gcol, non_blocking=True)
            self.ready[slot].record(self.stream)
        self.issued += 1

    def get(self):
        """the next view's inputs (views come in cyclic order); keeps one further copy in flight"""
        while self.issued < self.consumed + 2:
            self._issue()
        slot = self.consumed & 1
        torch.cuda.current_stream().wait_event(self.ready[slot])
        return self.buf[slot]

    def release(self):
        self.free[self.consumed & 1].record()
        self.consumed += 1


def e2e_stepper(mod, views, grads, arena, world):
    """End-to-end step through the reference-shaped `_C` functions with HOST buffers: for every view the camera and
    its colour-supervision gradient come from pinned host memory, forward + backward run, gradients are summed;
    one all-reduce when world > 1; two scalars are read back."""
    import harness as hz
    from lsx_b200.multiview import BWD_TO_GROUP
    feeder = HostFeeder(views, grads)
    g = dict(grads)

    def step():
        arena.zero_()
        acc = None
        for i, vw in enumerate(views):
            d_cam, d_gcol = feeder.get()
            fargs = list(vw["fargs"])
            fargs[11], fargs[12], fargs[19] = d_cam[:16].view(4, 4), d_cam[16:32].view(4, 4), d_cam[32:35]
            g["color"] = d_gcol
            fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*fargs)))
            bwd = dict(zip(hz.BWD_NAMES, mod.rasterize_gaussians_backward(*hz.native_backward_args(fargs, fwd, g))))
            feeder.release()
            arena.accumulate({gname: bwd[k] for k, gname in BWD_TO_GROUP.items()})
            acc = fwd["color"].sum() if acc is None else acc + fwd["color"].sum()
        if world > 1:
            arena.all_reduce()
        return torch.stack([acc, arena.views["means3D"].sum()]).to("cpu", non_blocking=False)
    return step, feeder.h2d_bytes, 8


def module_e2e_stepper(scene, views, grads, bg, F, arena, world):
    """Same, through the public nn.Module + autograd (the call a LangScene-X user makes).  The rasterizer's inputs are the
    leaves here, so the module's multi-view extension hands the arena to the backward kernel: every view's parameter
    gradients are written / added in place (no .grad tensors, no zero + add passes), then one all-reduce."""
    from diff_LangSurf_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    feeder = HostFeeder(views, grads)
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    params = dict(means3D=leaf(scene.means3D), shs=leaf(scene.shs), lang=leaf(scene.language_feature),
                  inst=leaf(scene.instance_feature), opac=leaf(scene.opacities), scales=leaf(scene.scales),
                  rots=leaf(scene.rotations))
    ams = [leaf(vw["am"]) for vw in views]   # all_map is a per-view derived input (built by the render wrapper)
    m2 = torch.zeros_like(scene.means3D, requires_grad=True)
    m2a = torch.zeros_like(scene.means3D, requires_grad=True)
    c0 = views[0]["cam"]
    sink = arena.grad_buffers()

    def step():
        for p in ams + [m2, m2a]:
            p.grad = None
        acc = None
        for i, am in enumerate(ams):
            d_cam, d_gcol = feeder.get()
            s = GaussianRasterizationSettings(c0.H, c0.W, c0.tanfovx, c0.tanfovy, bg, 1.0, d_cam[:16].view(4, 4),
                                              d_cam[16:32].view(4, 4), 3, d_cam[32:35], False, True, False, True)
            out = GaussianRasterizer(s, grad_buffers=sink, accumulate=i > 0)(
                means3D=params["means3D"], means2D=m2, means2D_abs=m2a, opacities=params["opac"], shs=params["shs"],
                language_feature_precomp=params["lang"], language_feature_instance_precomp=params["inst"],
                scales=params["scales"], rotations=params["rots"], all_map=am)
            color, lf, li, _, _, amap, depth = out
            torch.autograd.backward([color, lf, li, amap, depth],
                                    [d_gcol, grads["language_feature"], grads["instance_feature"], grads["all_map"],
                                     grads["plane_depth"]])
            feeder.release()
            acc = color.sum() if acc is None else acc + color.sum()
        if world > 1:
            arena.all_reduce()
        return torch.stack([acc, arena.views["means3D"].sum()]).to("cpu")
    return step, feeder.h2d_bytes, 8


def algorithmic_bytes(P, P_vis, R, W, H, M, F, Fi):
    """Compulsory HBM traffic per fwd+bwd iteration, SURVEY.md §8d."""
    Ct = 3 + F + Fi + 5
    fwd = P * (12 + 12 + 16 + 4 + 12 * M + 4 * (F + Fi + 5)) + P_vis * 75 + 8 * P + R * 12 + R * 24 * 2 + R * 8 + \
        R * (28 + 4 * Ct) + W * H * 4 * (Ct + 1 + 2) + 4 * P * 2
    bwd = R * (28 + 4 * Ct) + W * H * 4 * ((Ct + 1) + 5 + 2) + P_vis * (236 + 4 * (F + Fi + 5)) + \
        P * 4 * (3 + 3 + 3 + 3 + F + Fi + 5 + 4 + 1 + 6 + 3 * M + 3 + 4)
    return fwd, bwd


def run_cpu_sample(threads=None):
    """The reference's CPU-executable path on a bounded sample (BASELINE north_star): the PyTorch SH -> RGB and
    world-covariance preprocess the reference can run instead of the CUDA one (pipe.convert_SHs_python / compute_cov3D_python,
    gaussian_renderer/__init__.py:101-121; restated in oracle/preprocess_torch_oracle.py and pinned to the reference's own
    functions) with all host threads, forward AND autograd backward, feeding the per-pixel compositing restatement
    (oracle/lsx_oracle.c, OpenMP over tiles) forward + backward.  Sample: the C3 generator scaled by 1/4 in pixels and
    Gaussians (960x540, 250 000 Gaussians, same splat size in pixels => same per-pixel list statistics); best of 2."""
    from oracle import oracle as orc
    from oracle import preprocess_torch_oracle as pto
    from lsx_b200.synthetic import make_all_map, make_camera, make_scene, make_upstream_grads
    P, W, H, F = 250_000, 960, 540, 16
    if threads:
        orc.set_num_threads(threads)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    scene, cam = make_scene(P, W, H, F=F, seed=0), make_camera(W, H)
    am, g = make_all_map(scene, cam), make_upstream_grads(W, H, F)
    n = lambda t: t.detach().numpy()
    best, parts = 1e30, None
    for _ in range(2):
        t0 = time.perf_counter()
        shs, scales, rots = (t.clone().requires_grad_(True) for t in (scene.shs, scene.scales, scene.rotations))
        rgb = pto.sh_to_rgb(3, shs, scene.means3D, cam.campos)
        cov = pto.world_covariance(scales, 1.0, rots)
        t1 = time.perf_counter()
        o = orc.rasterize_forward(n(scene.means3D), n(scene.opacities), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos),
                                  W, H, cam.tanfovx, cam.tanfovy, [0, 0, 0], colors_precomp=n(rgb), cov3D_precomp=n(cov),
                                  language_feature=n(scene.language_feature), instance_feature=n(scene.instance_feature),
                                  all_map=n(am))
        gb = orc.rasterize_backward(o, n(g["color"]), n(g["language_feature"]), n(g["instance_feature"]), n(g["all_map"]),
                                    n(g["plane_depth"]))
        t2 = time.perf_counter()
        torch.autograd.backward([rgb, cov], [torch.from_numpy(gb["colors"]), torch.from_numpy(gb["cov3D"])])
        t3 = time.perf_counter()
        if t3 - t0 < best:
            best, parts = t3 - t0, (t1 - t0, t2 - t1, t3 - t2)
    return {"value": W * H / best / 1e6, "unit": "MPix/s", "cores": max(orc.num_threads(), torch.get_num_threads()), "kind": "port",
            "sample": f"fwd+bwd, best of 2, on the C3 generator scaled 1/4 ({P} Gaussians, {W}x{H}, F=16 + 3 instance + 5 map "
                      f"channels, R={o['num_rendered']}): torch-CPU eval_sh + covariance (the reference's Python preprocess path, "
                      f"{torch.get_num_threads()} threads) {parts[0]:.2f} s, compositing restatement oracle/lsx_oracle.c fwd+bwd "
                      f"({orc.num_threads()} OpenMP threads) {parts[1]:.2f} s, torch autograd through the preprocess {parts[2]:.2f} s; "
                      f"{best:.2f} s/iter; host has {os.cpu_count()} logical cores; not bit-comparable with the CUDA path"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="new", choices=["new", "reference", "reference-cpu"])
    ap.add_argument("--config", default="C3", help="C1..C5 (rasterizer fwd+bwd; C3 = headline) | C4-loop | C5-stress (bench_loop.py)")
    ap.add_argument("--views-per-gpu", type=int, default=None,
                    help="views rendered by every GPU per step (default 8: 8 x 8 GPUs = the 64-view batch of BASELINE config 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="loop configs: per-group async all-reduces under the last view's backward "
                                                         "instead of one blocking collective (measured slower at N = 2)")
    ap.add_argument("--no-parity", action="store_true", help="loop configs, N > 1: skip the untimed all-reduce parity check")
    ap.add_argument("--knn", action="store_true", help="loop configs: also time distCUDA2 over the scene's points")
    args = ap.parse_args()
    args.views_per_gpu_set = args.views_per_gpu is not None
    V = max(1, args.views_per_gpu if args.views_per_gpu_set else 8)
    # Untimed warm-up: at least 3 steps and at least 16 single-view passes, so that torch's caching allocator has
    # created every segment it needs (cudaMalloc costs 5-40 ms) before the timed region starts.
    warmup = max(args.warmup, 3, -(-16 // V))

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # stdout carries exactly one JSON line: libraries (e.g. NCCL's version banner) get stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    if args.impl == "reference-cpu":
        if rank == 0:
            cb = run_cpu_sample()
            emit({"impl": "reference-cpu", "metric": METRIC, "value": cb["value"], "unit": "MPix/s",
                  "n_gpus": 0, "steps": 1, "warmup": 1, "higher_is_better": True, "cpu_baseline": cb,
                  "e2e": {"value": cb["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    if args.config in ("C4-loop", "C5-stress"):
        import bench_loop
        bench_loop.run(args, emit, rank, local_rank, world, device, time_loop, ClockSampler)
        if world > 1:
            dist.destroy_process_group()
        return

    import harness as hz
    from lsx_b200.multiview import GradArena     # pure torch; the reference arm never loads liblsx_b200.so
    new = args.impl == "new"
    if new:
        from lsx_b200 import _lib, ops

    hbm_peak, peak_src, peaks_json = load_peaks()
    c, scene, grads, bg, views = build_views(args.config, device, rank, world, V)
    P, W, H, F = c["P"], c["W"], c["H"], c["F"]
    M, Fi = 16, 3
    Ct = 3 + F + Fi + 5

    if not new:
        mod = hz.ref_rast_for(F)
        if mod is None:
            if rank == 0:
                emit({"impl": "reference", "unavailable": "oracle/_ref/*.so not built (reference tree was not mounted at build time)"})
            return
    else:
        mod = ops
    arena = GradArena.allocate(P, M, F, Fi, device)

    sampler = ClockSampler(local_rank) if rank == 0 else None   # runs through both timed regions
    # ---- per-view statistics (untimed): R, P_vis, and S / B / visits counted on the device ------------------
    stats_views = []
    for i, vw in enumerate(views):
        fwd = dict(zip(hz.FWD_NAMES, mod.rasterize_gaussians(*vw["fargs"])))
        torch.cuda.synchronize()
        st = {"yaw": round(vw["yaw"], 3), "R": int(fwd["num_rendered"]), "P_vis": int((fwd["radii"] > 0).sum())}
        if new:
            rs = ops.render_stats(fwd["num_rendered"], fwd["geom"], fwd["binning"], fwd["img"], P, H, W, Ct)
            st.update({k: rs[k] for k in ("S", "B", "V", "Vb", "L")})
        elif i == 0:
            nbuf = hz.parse_ref_buffers(fwd["geom"], fwd["binning"], fwd["img"], P, st["R"], W, H)
            st["S"] = int(nbuf["n_contrib"].long().sum())
            del nbuf
        stats_views.append(st)
        del fwd
    R, P_vis, S = stats_views[0]["R"], stats_views[0]["P_vis"], stats_views[0]["S"]

    # ---- device-resident timing ----------------------------------------------------------------------
    step = batch_stepper(mod, views, grads, arena, world)
    launches0 = _lib.kernel_launch_count() if new else 0
    step_detail = {}
    ms_total = time_loop(step, args.steps, warmup, world, step_detail)
    launches = (_lib.kernel_launch_count() - launches0) if new else 0
    ms_step = ms_total / args.steps
    mpix = world * V * W * H / (ms_step * 1e-3) / 1e6

    # ---- per-stage device times (separate short run so the event hooks do not touch the headline number) ----
    stages = {}
    if new:
        _lib.profile_enable(True)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        stages = {k: v / (2 * V) for k, v in _lib.profile_read().items() if v > 0}   # ms per view
        _lib.profile_enable(False)

    # ---- end-to-end through the public API with host buffers ---------------------------------------------------
    if new:
        estep, h2d, d2h = module_e2e_stepper(scene, views, grads, bg, F, arena, world)
    else:
        estep, h2d, d2h = e2e_stepper(mod, views, grads, arena, world)
    ms_e2e = time_loop(estep, args.steps, max(3, warmup // 2), world) / args.steps
    clocks = sampler.stop() if sampler else None
    collective = time_collective(arena.flat, world)
    e2e = {"value": world * V * W * H / (ms_e2e * 1e-3) / 1e6, "unit": "MPix/s", "ms_per_step": ms_e2e,
           "ms_per_iter": ms_e2e / V, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "diff_LangSurf_rasterization.GaussianRasterizer + autograd" if new
                  else "reference _C.rasterize_gaussians/_backward (pybind)",
           "d2h_note": "training-shaped step: the rendered images are consumed on the device (upstream gradients come from the "
                       "host instead of a loss), so only two scalars — a checksum of the images and of the position gradients — "
                       "leave the GPU each step"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fwd_b = bwd_b = 0
    for st in stats_views:
        fb, bb = algorithmic_bytes(P, st["P_vis"], st["R"], W, H, M, F, Fi)
        fwd_b, bwd_b = fwd_b + fb, bwd_b + bb
    out = {
        "metric": METRIC, "value": mpix, "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_step, "ms_per_iter": ms_step / V, "ms_per_step_median": step_detail.get("median_ms"),
        "ms_per_step_max": step_detail.get("max_ms"),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": args.impl,
        "config": {"workload": f"{args.config}: {P} Gaussians, {W}x{H}, SH degree 3, F={F} language + {Fi} instance + 5 map "
                               f"channels + plane depth; one step = fwd+bwd of {V} views per GPU ({world * V} views in all, "
                               f"yaw -15..15 deg), gradients summed into one flat arena"
                               + (", + ONE NCCL all-reduce of the arena" if world > 1 else ""),
                   "views_per_gpu": V, "ms_per_iter_is": "ms_per_step / views_per_gpu = one single-view fwd+bwd (BASELINE's ms/iter)",
                   "l2": "per-view working set (~1.6 GB of inputs, records, lists, images, gradients) exceeds the 126 MB L2; no flush needed",
                   "upstream_grads": "fixed N(0,1)/(W*H) tensors, no loss inside the timed region"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "collective": collective,
        "stats": {"P_vis": P_vis, "R": R, "S": S, "B": stats_views[0].get("B"), "Ct": Ct, "per_view": stats_views,
                  "what": "S = (pixel, entry) tests per render pass of the reference's loops = sum of n_contrib; B = (pixel, entry) "
                          "blends; V = (8x4 block, entry) visits of this library's backward; Vb = visits in which a pixel blends; "
                          "L = total length of the per-block lists; counted on the device by lsx_render_stats"},
        "stages_ms": stages,
    }
    # roofline of the dominant kernel (live stage times, per launch = per view; means over this rank's views)
    if stages:
        dom = max(("render_fwd", "render_bwd"), key=lambda k: stages.get(k, 0.0))
        mean = lambda k: sum(st[k] for st in stats_views) / len(stats_views)
        Rm, Pm, Sm, Bm = mean("R"), mean("P_vis"), mean("S"), mean("B")
        peaks = {}
        try:
            import ctypes
            scratch = torch.empty(64 << 20, dtype=torch.uint8, device=device)
            for kind, name in ((0, "fp32_ffma_tflops"), (1, "mufu_ex2_gops"), (2, "red_add_f32_gops")):
                val = ctypes.c_double(0)
                _lib.check(_lib.load().lsx_microbench(kind, scratch.data_ptr(), scratch.numel(), ctypes.byref(val),
                                                      torch.cuda.current_stream().cuda_stream), "microbench")
                peaks[name] = val.value
            del scratch
        except Exception as e:  # noqa: BLE001
            peaks = {"error": str(e)[:200]}
        out["peaks"] = peaks
        fp32_peak = peaks.get("fp32_ffma_tflops") or 2 * 128 * 148 * 1.965e-3
        fp32_src = ("measured live: lsx_microbench FFMA (MEASURED_PEAKS.json has no fp32 figure)" if "fp32_ffma_tflops" in peaks
                    else "theoretical 2 x 128 x 148 SMs x 1.965 GHz")
        # SURVEY.md 8d flop model: per (pixel, entry) test + per blend
        flops = {"render_fwd": 14 * Sm + (3 + 2 * Ct) * Bm, "render_bwd": 16 * Sm + (52 + 8 * Ct) * Bm}
        ncu, ncu_src = ncu_record()
        for name in ("render_fwd", "render_bwd"):
            if name not in stages:
                continue
            tf = flops[name] / (stages[name] * 1e-3) / 1e12
            obj = {"kernel": name + "_kernel", "bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s",
                   "frac": tf / fp32_peak, "traffic": (ncu or {}).get(name, {}).get("dram_bytes"),
                   "issue_slots_pct": (ncu or {}).get(name, {}).get("issue_active_pct"), "ncu_capture": ncu_src,
                   "peak_source": fp32_src, "launch_ms": stages[name], "flops_per_launch": flops[name],
                   "flop_model": ("14*S + (3 + 2*Ct)*B" if name == "render_fwd" else "16*S + (52 + 8*Ct)*B") +
                                 f" (SURVEY.md 8d) with S={Sm:.0f}, B={Bm:.0f}, Ct={Ct}: means over this rank's views"}
            out["roofline" if name == dom else "roofline_" + name] = obj
        rec = Rm * (28 + 4 * Ct)
        kbytes = rec + W * H * 4 * (Ct + 1 + 2) if dom == "render_fwd" else rec + W * H * 4 * ((Ct + 1) + 5 + 2) + Pm * 4 * (Ct + 8)
        ach = kbytes / (stages[dom] * 1e-3) / 1e9
        out["roofline_hbm"] = {"kernel": dom + "_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                               "frac": ach / hbm_peak, "peak_source": peak_src, "algorithmic_bytes": kbytes,
                               "note": "the same kernel against the HBM roof: R*(28+4*Ct) + pixel planes (+ P_vis*(Ct+8)*4 for bwd); "
                                       "it is FP32-issue bound (DESIGN.md 3), this view only shows how far from the memory roof it is"}
        it_bytes = fwd_b + bwd_b
        out["roofline_iteration"] = {"bound": "hbm", "algorithmic_bytes": it_bytes,
                                     "achieved": it_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": it_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak}
    if new and world == 1:
        ref = hz.ref_rast_for(F)
        if ref is not None:
            rstep = batch_stepper(ref, views, grads, arena, 1)
            rsteps = max(2, args.steps // 4)
            rms = time_loop(rstep, rsteps, 1, 1) / rsteps
            out["ref_cuda"] = {"ms_per_step": rms, "ms_per_iter": rms / V, "value": V * W * H / (rms * 1e-3) / 1e6, "unit": "MPix/s",
                               "what": f"reference CUDA rasterizer (oracle/_ref/ref_rast_f{16 if F == 16 else 3}.so, sm_100a recompile), same step, same process",
                               "speedup_device_resident": rms / ms_step}
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = run_cpu_sample()
            except Exception as e:
                out["cpu_baseline"] = {"value": None, "unit": "MPix/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if args.impl == "reference":
        out["cpu_baseline"] = {"value": out["value"], "unit": "MPix/s", "cores": 0, "kind": "reference",
                               "sample": "this arm runs the reference's own implementation of the path — its CUDA rasterizer, unmodified "
                                         "sources recompiled for sm_100a (oracle/_ref) — on the GPU; the reference has no CPU "
                                         "implementation of this path.  The CPU restatement is timed by --impl reference-cpu"}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
