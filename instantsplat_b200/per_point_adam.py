"""Drop-in for `scene.per_point_adam.PerPointAdam` (boundary B4):
/root/reference/scene/per_point_adam.py:17-98.  Same constructor, param-group keys
(`params, lr, betas, eps, weight_decay, per_point_lr, name`), state layout
(`state[p] = {step, exp_avg, exp_avg_sq}`) and error behaviour, but `step()` is ONE launch of the
fused sm_100a kernel (gs_adam.cu) over every tensor, with no host synchronisation."""
from __future__ import annotations

import ctypes

import torch
from torch.optim import Optimizer

from . import _lib
from ._lib import ADAM_MAX_TENSORS, GsbAdamTensor, check


def launch_adam(entries, flags):
    """entries: list of dicts(param, grad, exp_avg, exp_avg_sq, per_point_lr, step_size, beta1, beta2,
    eps, weight_decay, grad_scale, row_len)."""
    L = _lib.lib()
    for i in range(0, len(entries), ADAM_MAX_TENSORS):
        chunk = entries[i:i + ADAM_MAX_TENSORS]
        arr = (GsbAdamTensor * len(chunk))()
        for t, e in zip(arr, chunk):
            t.param, t.grad = e["param"].data_ptr(), e["grad"].data_ptr()
            t.exp_avg, t.exp_avg_sq = e["exp_avg"].data_ptr(), e["exp_avg_sq"].data_ptr()
            ppl = e.get("per_point_lr")
            t.per_point_lr = None if ppl is None else ppl.data_ptr()
            t.numel, t.row_len = e["param"].numel(), int(e.get("row_len", 1))
            t.grad_scale = float(e.get("grad_scale", 1.0))
            t.step_size, t.beta1, t.beta2 = float(e["step_size"]), float(e["beta1"]), float(e["beta2"])
            t.eps, t.weight_decay = float(e["eps"]), float(e["weight_decay"])
        check(L.gsb_adam_step(len(chunk), arr, flags.data_ptr(), _lib.stream_ptr()), "gsb_adam_step")


class PerPointAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if not all(0.0 <= x for x in [lr, eps, weight_decay]):
            raise ValueError(f"Invalid learning parameters: lr={lr}, eps={eps}, weight_decay={weight_decay}")
        if not all(0.0 <= beta < 1.0 for beta in betas):
            raise ValueError(f"Invalid beta parameters: {betas}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, per_point_lr=None)
        super().__init__(params, defaults)
        self._flags = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = []
        for group in self.param_groups:
            per_point_lr = group.get("per_point_lr")
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad
                if grad.is_sparse:
                    raise RuntimeError("PerPointAdam does not support sparse gradients")
                if not p.is_cuda:
                    raise _lib.GsbError("PerPointAdam (B200) needs CUDA parameters; there is no CPU path")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p)
                    state["exp_avg_sq"] = torch.zeros_like(p)
                state["step"] += 1
                bc1 = 1 - beta1 ** state["step"]
                bc2 = 1 - beta2 ** state["step"]
                step_size = group["lr"] * (bc2 ** 0.5 / bc1)
                row_len = 1
                ppl = None
                if per_point_lr is not None:
                    if not isinstance(per_point_lr, torch.Tensor):
                        raise TypeError("per_point_lr must be a torch.Tensor")
                    if per_point_lr.device != p.device:
                        raise ValueError("per_point_lr must be on the same device as parameter")
                    expected_shape = p.shape[:1] + (1,) * (p.dim() - 1)
                    if per_point_lr.shape != expected_shape:
                        raise ValueError(f"Invalid per_point_lr shape. Expected {expected_shape}, got {per_point_lr.shape}")
                    ppl = per_point_lr.float().contiguous()
                    row_len = p.numel() // max(1, p.shape[0])
                if not (p.is_contiguous() and grad.is_contiguous() and p.dtype == torch.float32):
                    raise _lib.GsbError("PerPointAdam (B200) needs contiguous fp32 parameters and gradients")
                entries.append(dict(param=p, grad=grad, exp_avg=state["exp_avg"], exp_avg_sq=state["exp_avg_sq"],
                                    per_point_lr=ppl, step_size=step_size, beta1=beta1, beta2=beta2,
                                    eps=group["eps"], weight_decay=group["weight_decay"], row_len=row_len))
        if entries:
            dev = entries[0]["param"].device
            if self._flags is None or self._flags.device != dev:
                self._flags = torch.zeros(ADAM_MAX_TENSORS, dtype=torch.int32, device=dev)
            launch_adam(entries, self._flags)
        return loss
