"""(1-l) L1 + l (1 - SSIM), forward + backward: fused CUDA kernels vs the torch-op formulation the reference uses."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "langscene-x_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from lsx_b200.loss import image_loss  # noqa: E402
from oracle import image_loss_oracle as orc  # noqa: E402  (tool, not product)


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


for (W, H) in ((1920, 1080), (720, 480)):
    gt = torch.rand(3, H, W, device="cuda")
    img = (gt + 0.1 * torch.randn(3, H, W, device="cuda")).clamp(0, 1).requires_grad_(True)

    def fused():
        img.grad = None
        image_loss(img, gt, 0.2)[0].backward()

    def torch_ops():
        img.grad = None
        orc.image_loss(img, gt, 0.2).backward()

    tf, tt = timeit(fused), timeit(torch_ops)
    print(json.dumps({"W": W, "H": H, "fused_fwd_bwd_ms": tf, "torch_ops_fwd_bwd_ms": tt, "speedup": tt / tf}))
