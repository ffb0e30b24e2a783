"""One-kernel Adam over the flat parameter arena.

`ArenaAdam(params, lrs)` replaces `torch.optim.Adam(groups, lr=0.0, eps=1e-15)` of the reference
(field_construction/scene/gaussian_model.py:313-328): same update rule, one learning rate per arena group, but parameters,
the (all-reduced) gradients and both moments live in flat fp32 arenas with the layout of lsx_b200.multiview.GradArena, so
`step(grads)` is a single streaming CUDA kernel.  No CPU path.
"""
import ctypes

import torch

from . import _lib
from .multiview import GradArena


class ArenaAdam:
    def __init__(self, params: GradArena, lrs, betas=(0.9, 0.999), eps=1e-15):
        if not params.flat.is_cuda:
            raise RuntimeError("ArenaAdam needs CUDA arenas (this operator has no CPU path)")
        self.betas, self.eps, self.step_count = betas, eps, 0
        self.lrs = dict(lrs)
        self.rebind(params, torch.zeros_like(params.flat), torch.zeros_like(params.flat))

    def rebind(self, params: GradArena, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor):
        """Point the optimiser at new arenas (after lsx_b200.densify.densify_and_prune: pass result.params and the `.flat`
        tensors of result.exp_avg / result.exp_avg_sq).  The step count is kept, as torch.optim.Adam keeps its per-parameter
        `step` through the reference's cat / prune surgery (gaussian_model.py:520-581)."""
        if exp_avg.numel() != params.flat.numel() or exp_avg_sq.numel() != params.flat.numel():
            raise RuntimeError("moment arenas and parameter arena have different layouts")
        self.params, self.exp_avg, self.exp_avg_sq = params, exp_avg, exp_avg_sq
        # groups in arena order (dicts keep insertion order): lsx_b200.multiview.PARAM_GROUPS for a GradArena-shaped arena,
        # the reference's own optimizer groups for a lsx_b200.densify.ParamArena
        names = [n for n in params.offsets if params.offsets[n][1] > 0]
        self._names = names
        begins = [params.offsets[n][0] for n in names] + [params.flat.numel()]
        self._begin = (ctypes.c_int64 * len(begins))(*begins)
        return self

    def set_lr(self, name, lr):
        """e.g. the exponential xyz schedule of GaussianModel.update_learning_rate"""
        self.lrs[name] = float(lr)

    def step(self, grads: GradArena):
        if grads.flat.numel() != self.params.flat.numel():
            raise RuntimeError("gradient arena and parameter arena have different layouts")
        self.step_count += 1
        lr = (ctypes.c_float * len(self._names))(*[float(self.lrs.get(n, 0.0)) for n in self._names])
        dev = self.params.flat.device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().lsx_arena_adam_step(
                self.params.flat.numel(), len(self._names), self._begin, lr, self.step_count, self.betas[0], self.betas[1],
                self.eps, self.params.flat.data_ptr(), grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "arena_adam_step")
