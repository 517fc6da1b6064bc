"""Fused tail of the training step: ``clip_grad_norm_`` + ``torch.optim.Adam`` (L2 weight decay) as
two kernel launches over flat buffers -- the pair of calls at train.py:273-275 of the reference.

``FusedClipAdam`` is a ``torch.optim.Optimizer``: it has ``param_groups`` (``CosineAnnealingLR`` and the
``optimizer.param_groups[0]['lr']`` reads of train.py:224,282 work unchanged), ``state_dict()`` /
``load_state_dict()`` in torch.optim.Adam's per-parameter layout (``step``, ``exp_avg``, ``exp_avg_sq``), so
``utils.CheckpointSaver.save`` / ``utils.load_model_checkpoint`` (utils.py:141,160) round-trip it and a checkpoint
written by either optimiser loads into the other.

The parameters are re-pointed into one flat fp32 buffer laid out exactly like ``FlatGradSync``'s gradient buffer
(16-byte aligned slices), so the update is one element-wise pass; learning rate and step count are device scalars,
which keeps the step capturable in a CUDA graph while a scheduler changes the rate between replays.
CUDA only: there is no CPU fallback (the reference's own torch optimiser is the CPU path)."""
import ctypes as C

import torch

from . import _lib
from .dist import FlatGradSync


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0,
                 grad_sync=None):
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("no trainable parameters")
        dev = plist[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in plist):
            raise RuntimeError("FusedClipAdam needs float32 CUDA parameters on one device (no CPU fallback)")
        self.sync = grad_sync if grad_sync is not None else FlatGradSync(plist, align=4)
        if [id(p) for p in self.sync.params] != [id(p) for p in plist]:
            raise ValueError("grad_sync was built over a different parameter list")
        if any(o % 4 for o in self.sync.offsets):
            raise ValueError("grad_sync must be built with align=4 (16-byte aligned parameter slices)")
        n = self.sync.flat.numel()
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(plist, self.sync.offsets):
                view = self.flat[off: off + p.numel()].view_as(p)
                view.copy_(p.detach())
                p.data = view                                   # same Parameter objects, storage now inside the flat buffer
        super().__init__(plist, dict(lr=float(lr), betas=tuple(betas), eps=float(eps), weight_decay=float(weight_decay),
                                     max_grad_norm=float(max_grad_norm)))
        self.params = plist
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.lr = torch.tensor([float(lr)], device=dev, dtype=torch.float32)
        self._lr_host = float(lr)
        self.total_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self._ws_bytes = _lib.lib().dcgru_clip_adam_workspace(n)
        self._ws = torch.empty(max(self._ws_bytes, 16), device=dev, dtype=torch.uint8)
        self._bind_state()

    # ---- state in torch.optim.Adam's layout: per-parameter views into the flat moment buffers -----------------
    def _bind_state(self):
        for p, off in zip(self.params, self.sync.offsets):
            self.state[p] = {"step": self.step_count.to(torch.float32).reshape(()),   # refreshed by state_dict()
                             "exp_avg": self.exp_avg[off: off + p.numel()].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off: off + p.numel()].view_as(p)}

    def state_dict(self):
        step = self.step_count.to(torch.float32).reshape(()).clone()      # one D2H-free device copy, shared
        for p in self.params:
            self.state[p]["step"] = step
        self.param_groups[0]["lr"] = self.param_groups[0].get("lr", self._lr_host)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        before = dict(self.param_groups[0])
        super().load_state_dict(state_dict)                               # casts to the parameters' device / dtype
        # a checkpoint written by torch.optim.Adam has no max_grad_norm (clipping is a separate call there, train.py:273):
        # options the loaded group does not carry keep this optimiser's own values
        for k, v in before.items():
            if k != "params":
                self.param_groups[0].setdefault(k, v)
        steps = set()
        with torch.no_grad():
            for p, off in zip(self.params, self.sync.offsets):
                st = self.state.get(p, {})
                if "exp_avg" in st:
                    self.exp_avg[off: off + p.numel()].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off: off + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(float(st["step"])))
                else:
                    self.exp_avg[off: off + p.numel()].zero_()
                    self.exp_avg_sq[off: off + p.numel()].zero_()
                    steps.add(0)
        if len(steps) > 1:
            raise ValueError(f"FusedClipAdam keeps one step count for all parameters; the checkpoint has {sorted(steps)}")
        self.step_count.fill_(steps.pop() if steps else 0)
        self._bind_state()
        self._sync_lr()

    # ---- learning rate: param_groups[0]['lr'] (what schedulers write) mirrored into the device scalar -------------
    def set_lr(self, lr):
        """what a scheduler does between steps (device-side: valid between replays of a captured graph too)"""
        self.param_groups[0]["lr"] = float(lr)
        self._sync_lr()

    def _sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr.fill_(lr)
            self._lr_host = lr

    def zero_grad(self, set_to_none=False):
        """keeps the ``.grad`` views into the flat gradient buffer alive (``set_to_none`` would detach them)"""
        self.sync.zero()

    def _check_bound(self):
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * 4
        glo, ghi = self.sync.flat.data_ptr(), self.sync.flat.data_ptr() + self.sync.flat.numel() * 4
        for p in self.params:
            if not (lo <= p.data_ptr() < hi):
                raise RuntimeError("FusedClipAdam: a parameter no longer lives in the flat buffer (model.to()/.cuda() "
                                   "after constructing the optimiser?) -- build the optimiser after moving the model")
            if p.grad is None or not (glo <= p.grad.data_ptr() < ghi):
                raise RuntimeError("FusedClipAdam: a .grad no longer aliases the flat gradient buffer "
                                   "(zero_grad(set_to_none=True) on the module?) -- use optimizer.zero_grad()")

    @torch.no_grad()
    def step(self, closure=None):
        """clip the (already all-reduced) flat gradient to ``max_grad_norm`` and apply one Adam update; returns the
        pre-clip global norm as a device scalar (no host sync).  ``grad_sync.world > 1`` with ``average_in_optimizer``
        folds the 1/world scale of the data-parallel average into this pass."""
        if closure is not None:
            raise RuntimeError("FusedClipAdam.step() does not take a closure")
        self._check_bound()
        self._sync_lr()
        g = self.param_groups[0]
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().dcgru_clip_adam_step(
            p(self.flat), p(self.sync.flat), p(self.exp_avg), p(self.exp_avg_sq), self.flat.numel(), p(self.lr),
            p(self.step_count), g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], g["max_grad_norm"],
            float(self.sync.pending_scale()), p(self.total_norm), p(self._ws), self._ws_bytes,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "clip_adam_step")
        return self.total_norm
