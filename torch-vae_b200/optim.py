"""Fused AdamW for the drop-in VAE.

The reference builds ``torch.optim.AdamW`` over two parameter groups, ``model.encoder`` and
``model.decoder`` (train.py:210-228), and steps it once per batch (train.py:656).  On a GPU the
stock foreach implementation costs ~14 launches of ~20 us per step for 28 small tensors; this
subclass keeps torch's constructor, ``param_groups``, ``state`` layout (``step`` as a device
float scalar, ``exp_avg``, ``exp_avg_sq``) and ``state_dict()`` -- so schedulers
(``OneCycleLR``, train.py:232-238) and checkpoints (train.py:445-460) work unchanged -- and
replaces ``step()`` by one C-ABI call (``vae_adamw_step``).  CUDA only, fp32 parameters.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops


class FusedAdamW(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, **kw):
        if amsgrad:
            raise NotImplementedError("FusedAdamW: amsgrad is not supported")
        for bad in ("maximize", "differentiable"):
            if kw.get(bad):
                raise NotImplementedError(f"FusedAdamW: {bad} is not supported")
        kw.pop("foreach", None)
        kw.pop("fused", None)
        kw["capturable"] = True  # step counters live on the device
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, **kw)

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        by_betas = {}
        keep = []
        for group in self.param_groups:
            lr = group["lr"]
            if isinstance(lr, torch.Tensor):
                lr = float(lr)
            key = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                st = self._init_state(p)
                by_betas.setdefault(key, []).append(
                    _lib.AdamWTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                     st["step"].data_ptr(), p.numel(), float(lr), float(group["weight_decay"])))
        lib = ops._L()
        for (b1, b2, eps), entries in by_betas.items():
            arr = (_lib.AdamWTensor * len(entries))(*entries)
            rc = lib.vae_adamw_step(C.cast(arr, C.c_void_p), len(entries), b1, b2, eps, ops._stream())
            _lib.check(rc, "vae_adamw_step")
        return loss
