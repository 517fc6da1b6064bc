"""Fused tail of the training step: ``clip_grad_norm_`` + ``torch.optim.Adam`` (L2 weight decay) as
two kernel launches over flat buffers -- the pair of calls at train.py:273-275 of the reference.

The parameters are re-pointed into one flat fp32 buffer laid out exactly like ``FlatGradSync``'s gradient buffer
(16-byte aligned slices), so the update is one element-wise pass; learning rate and step count are device scalars,
which keeps the step capturable in a CUDA graph while a scheduler changes the rate between replays.
CUDA only: there is no CPU fallback (the reference's own torch optimiser is the CPU path)."""
import ctypes as C

import torch

from . import _lib
from .dist import FlatGradSync


class FusedClipAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0,
                 grad_sync=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise RuntimeError("FusedClipAdam needs float32 CUDA parameters on one device (no CPU fallback)")
        self.sync = grad_sync if grad_sync is not None else FlatGradSync(self.params, align=4)
        if [id(p) for p in self.sync.params] != [id(p) for p in self.params]:
            raise ValueError("grad_sync was built over a different parameter list")
        if any(o % 4 for o in self.sync.offsets):
            raise ValueError("grad_sync must be built with align=4 (16-byte aligned parameter slices)")
        n = self.sync.flat.numel()
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.sync.offsets):
                view = self.flat[off: off + p.numel()].view_as(p)
                view.copy_(p.detach())
                p.data = view                                   # same Parameter objects, storage now inside the flat buffer
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.lr = torch.tensor([float(lr)], device=dev, dtype=torch.float32)
        self.total_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.betas, self.eps, self.weight_decay, self.max_grad_norm = betas, float(eps), float(weight_decay), float(max_grad_norm)
        self._ws_bytes = _lib.lib().dcgru_clip_adam_workspace(n)
        self._ws = torch.empty(max(self._ws_bytes, 16), device=dev, dtype=torch.uint8)

    def set_lr(self, lr):
        """what a scheduler calls between steps (device-side: valid inside a captured graph too)"""
        self.lr.fill_(float(lr))

    def zero_grad(self):
        self.sync.zero()

    def step(self):
        """clip the (already all-reduced) flat gradient to ``max_grad_norm`` and apply one Adam update; returns the
        pre-clip global norm as a device scalar (no host sync)"""
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().dcgru_clip_adam_step(
            p(self.flat), p(self.sync.flat), p(self.exp_avg), p(self.exp_avg_sq), self.flat.numel(), p(self.lr),
            p(self.step_count), self.betas[0], self.betas[1], self.eps, self.weight_decay, self.max_grad_norm,
            p(self.total_norm), p(self._ws), self._ws_bytes,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "clip_adam_step")
        return self.total_norm
