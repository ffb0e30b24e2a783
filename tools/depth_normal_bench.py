"""depth -> normal (+ alpha) forward + backward: fused CUDA kernels vs the torch-op formulation the reference uses."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.render_utils import depth_to_normal  # noqa: E402
from oracle import depth_normal_oracle as orc  # noqa: E402  (tool, not product)


def timeit(fn, n=50):
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


for (W, H) in ((1920, 1080), (720, 480)):
    depth = (4.0 + torch.rand(H, W, device="cuda")).requires_grad_(True)
    alpha = torch.rand(H, W, device="cuda")
    up = torch.randn(3, H, W, device="cuda")
    fx = fy = 0.866 * W
    cx, cy = W / 2, H / 2

    def fused():
        depth.grad = None
        depth_to_normal(depth, fx, fy, cx, cy, alpha=alpha).backward(up)

    def torch_ops():
        depth.grad = None
        orc.depth_to_normal(depth, fx, fy, cx, cy, alpha).backward(up)

    tf, tt = timeit(fused), timeit(torch_ops)
    bytes_alg = W * H * 4 * (1 + 1 + 3) + W * H * 4 * (1 + 1 + 3 + 1)   # fwd: depth, alpha in, normal out; bwd: + grad in, g_depth out
    print(json.dumps({"W": W, "H": H, "fused_fwd_bwd_ms": tf, "torch_ops_fwd_bwd_ms": tt, "speedup": tt / tf,
                      "algorithmic_GBps": bytes_alg / tf / 1e6}))
