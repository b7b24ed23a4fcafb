"""Flat-buffer optimiser state: one fused Adam + mean-teacher EMA kernel per step (csrc/elementwise.cu adam_ema_kernel).

Replaces the 100 `mul_/add_` launches of SEDTask4.update_ema (recipes/dcase2023_task4_baseline/local/sed_trainer.py:187-199)
and torch.optim.Adam (train_sed.py:199-201) with a single pass over contiguous parameter / gradient / moment buffers.
`FusedAdam` is a torch.optim.Optimizer (param_groups carry lr/betas/eps so desed_task's ExponentialWarmup can drive it).
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


def flat_layout(sizes):
    """Offsets of tensors of `sizes` elements inside a flat fp32 buffer: parameters() order, every tensor starting on a
    16-byte boundary (TMA / 128-bit loads read parameters and gradients straight out of the flat buffers).  Returns
    (offsets, total).  The parameter, gradient, Adam-moment and teacher buffers all use this ONE layout."""
    offs, off = [], 0
    for n in sizes:
        offs.append(off)
        off += (int(n) + 3) // 4 * 4
    return offs, off


def flatten_parameters(module):
    """Re-point every parameter of `module` at a view of ONE contiguous fp32 buffer (parameters() order, flat_layout).
    Returns the flat buffer.  Values, names, shapes and state_dict are unchanged."""
    params = list(module.parameters())
    if getattr(module, "_sedk_flat", None) is not None:
        flat = module._sedk_flat
        if all(p.data_ptr() == flat.data_ptr() + off * 4 for p, off in zip(params, module._sedk_offsets)):
            return flat
    offs, total = flat_layout([p.numel() for p in params])
    dev = params[0].device
    flat = torch.zeros(total, device=dev, dtype=torch.float32)
    for p, off in zip(params, offs):
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
    module._sedk_flat = flat
    module._sedk_offsets = offs
    return flat


def ema_alpha(alpha, global_step):
    """sed_trainer.py:196-197: use the true average until the exponential average is more correct."""
    return min(1 - 1 / (global_step + 1), alpha)


def update_ema(alpha, global_step, model, ema_model):
    """Drop-in for SEDTask4.update_ema (sed_trainer.py:187-199) - one kernel over the flat buffers."""
    a = ema_alpha(alpha, global_step)
    p = flatten_parameters(model)
    e = flatten_parameters(ema_model)
    check(lib().sedk_adam_ema(ptr(p), None, None, None, ptr(e), p.numel(), 0, 0.0, 0.0, 0.0, 0.0, 0, a, 1.0,
                              stream_ptr()), "sedk_adam_ema")
    return a


class FusedAdam(torch.optim.Optimizer):
    """Adam(lr, betas, eps) without weight decay / amsgrad - the configuration of train_sed.py:199-201 - on flat buffers.

    `step()` reads gradients from p.grad (generic path); `step_flat(gflat, ...)` consumes a flat gradient buffer in
    parameters() order and can fold the EMA teacher update into the same kernel."""

    def __init__(self, module_or_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if isinstance(module_or_params, torch.nn.Module):
            self.module = module_or_params
            params = list(module_or_params.parameters())
        else:
            self.module = None
            params = list(module_or_params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.flat = None
        self.m = self.v = None
        self.step_count = 0

    def _ensure(self):
        if self.flat is None:
            if self.module is None:
                raise RuntimeError("FusedAdam needs the nn.Module (to flatten its parameters) for the fused path")
            self.flat = flatten_parameters(self.module)
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)

    def step_flat(self, gflat, ema_flat=None, ema_a=0.0, grad_scale=1.0, hyper_dev=None):
        """One fused EMA + Adam pass.  With `hyper_dev` (device fp32[4]) the per-step scalars come from device memory
        (CUDA-graph replay); the caller then owns the step counter / bias corrections."""
        self._ensure()
        g = self.param_groups[0]
        n = self.flat.numel()
        if hyper_dev is not None:
            check(lib().sedk_adam_ema_dev(ptr(self.flat), ptr(gflat), ptr(self.m), ptr(self.v), ptr(ema_flat), n, 1,
                                          g["betas"][0], g["betas"][1], g["eps"], ptr(hyper_dev), stream_ptr()),
                  "sedk_adam_ema_dev")
            return
        self.step_count += 1
        check(lib().sedk_adam_ema(ptr(self.flat), ptr(gflat), ptr(self.m), ptr(self.v), ptr(ema_flat), n, 1, g["lr"],
                                  g["betas"][0], g["betas"][1], g["eps"], self.step_count, ema_a, grad_scale,
                                  stream_ptr()), "sedk_adam_ema")

    def hyper(self, step, ema_a=0.0, grad_scale=1.0):
        """The four per-step scalars of sedk_adam_ema_dev for optimiser step `step` (1-based)."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        return [g["lr"] / (1.0 - b1 ** step), 1.0 / (1.0 - b2 ** step) ** 0.5, ema_a, grad_scale]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._ensure()
        params = [p for p in self.param_groups[0]["params"]]
        offs, total = flat_layout([p.numel() for p in params])
        gflat = torch.zeros(total, device=self.flat.device, dtype=torch.float32)
        for p, off in zip(params, offs):
            if p.grad is not None:
                gflat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        self.step_flat(gflat)
        return loss

    def state_dict(self):
        sd = super().state_dict()
        sd["sedk"] = dict(step_count=self.step_count, m=None if self.m is None else self.m.clone(),
                          v=None if self.v is None else self.v.clone())
        return sd

    def load_state_dict(self, state_dict):
        extra = state_dict.pop("sedk", None)
        super().load_state_dict(state_dict)
        if extra is not None:
            self.step_count = extra["step_count"]
            if extra["m"] is not None:
                self._ensure()
                self.m.copy_(extra["m"])
                self.v.copy_(extra["v"])
