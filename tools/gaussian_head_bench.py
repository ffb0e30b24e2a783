"""Per-Gaussian render-wrapper head (activations + plane normal + all_map), forward + backward: fused vs torch ops."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.render_utils import gaussian_head  # noqa: E402
from lsx_b200.synthetic import make_camera  # noqa: E402
from oracle import gaussian_head_oracle as orc  # noqa: E402  (tool, not product)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cam = make_camera(1920, 1080, yaw_deg=5.0)
view, campos = cam.viewmatrix.cuda(), cam.campos.cuda()
for P in (1_000_000, 500_000):
    raw = [torch.randn(P, k, device="cuda").requires_grad_(True) for k in (3, 3, 4, 1)]
    ups = [torch.randn(P, k, device="cuda") for k in (3, 4, 1, 5)]

    def run(fn):
        def f():
            for r in raw:
                r.grad = None
            outs = fn(*raw, view, campos)
            torch.autograd.backward(outs, ups)
        return f

    tf, tt = timeit(run(gaussian_head)), timeit(run(orc.gaussian_head))
    print(json.dumps({"P": P, "fused_fwd_bwd_ms": tf, "torch_ops_fwd_bwd_ms": tt, "speedup": tt / tf,
                      "algorithmic_GBps": P * (44 + 52 + 96 + 44) / tf / 1e6}))
