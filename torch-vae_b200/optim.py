"""Fused AdamW + OneCycle schedule for the drop-in VAE.

The reference builds ``torch.optim.AdamW`` over two parameter groups, ``model.encoder`` and
``model.decoder`` (train.py:210-228), a ``OneCycleLR`` over them (train.py:232-238), and steps both
once per batch (train.py:656,659).  On a GPU the stock foreach optimizer costs ~14 launches of
~20 us per step for 28 small tensors.

``FusedAdamW`` keeps torch's constructor, ``param_groups``, ``state`` layout (``step`` as a device
float scalar, ``exp_avg``, ``exp_avg_sq``) and ``state_dict()`` -- so schedulers and checkpoints
(train.py:445-460) work unchanged -- and replaces ``step()`` by one C-ABI call.  The
hyper-parameters a scheduler changes (lr, betas, weight decay) live in a small DEVICE table that
the kernel reads, so a captured CUDA graph follows the schedule instead of replaying the values of
capture time:

  * with a host scheduler (``torch.optim.lr_scheduler.OneCycleLR``) ``step()`` re-uploads the table
    whenever ``param_groups`` changed (eager mode; changing them during graph capture raises);
  * ``FusedOneCycleLR`` advances the schedule ON the device (``vae_onecycle_step``: one tiny launch,
    capturable), mirroring OneCycleLR's constructor and arithmetic.

CUDA only, fp32 parameters, no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import math

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
        self._hyper = None          # device [n_groups, 8] fp32: lr, beta1, beta2, weight_decay, 1-beta1, 1-beta2, -, -
        self._hyper_host = None     # the host values last uploaded
        self._device_schedule = False  # True once a FusedOneCycleLR owns lr / beta1 on the device

    # ------------------------------------------------------------------ state
    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def load_state_dict(self, state_dict):
        """Accepts checkpoints written by this class or by the reference's stock ``torch.optim.AdamW``
        (train.py:445-460: ``capturable=False``, ``step`` a CPU tensor or a float): every ``step`` is moved
        to its parameter's device as a 0-dim fp32 tensor and ``capturable`` stays True."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            group["capturable"] = True
            group["foreach"] = None
            group["fused"] = None
            for p in group["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                step = st.get("step", 0.0)
                if not isinstance(step, torch.Tensor):
                    step = torch.tensor(float(step))
                st["step"] = step.detach().to(device=p.device, dtype=torch.float32).reshape(())
                for k in ("exp_avg", "exp_avg_sq"):
                    st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
        self._hyper_host = None  # re-upload the table at the next step

    # -------------------------------------------------------- hyper-parameters
    def _host_hyper(self):
        rows = []
        for group in self.param_groups:
            lr = group["lr"]
            lr = float(lr) if not isinstance(lr, torch.Tensor) else float(lr.item())
            b1, b2 = float(group["betas"][0]), float(group["betas"][1])
            rows.append((lr, b1, b2, float(group["weight_decay"]), 1.0 - b1, 1.0 - b2, 0.0, 0.0))
        return rows

    def hyper_table(self, device) -> torch.Tensor:
        """The device table the kernel reads; created on first use."""
        device = torch.device(device)
        if device.index is None:
            device = torch.device(device.type, torch.cuda.current_device())
        if self._hyper is None or self._hyper.device != device or self._hyper.shape[0] != len(self.param_groups):
            self._hyper = torch.zeros((len(self.param_groups), 8), dtype=torch.float32, device=device)
            self._hyper_host = None
        return self._hyper

    def _sync_hyper(self, device):
        table = self.hyper_table(device)
        if self._device_schedule and self._hyper_host is not None:
            return table  # a device-side scheduler owns lr / beta1
        rows = self._host_hyper()
        if rows != self._hyper_host:
            if torch.cuda.is_current_stream_capturing():
                if self._hyper_host is not None:
                    raise RuntimeError(
                        "FusedAdamW: param_groups changed on the host during CUDA-graph capture; a captured step "
                        "cannot follow a host-side scheduler -- use torch_vae_b200.optim.FusedOneCycleLR")
                raise RuntimeError("FusedAdamW: run one eager step() before capturing (the hyper-parameter table is not on the device yet)")
            table.copy_(torch.tensor(rows, dtype=torch.float32))
            self._hyper_host = rows
        return table

    # -------------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries, group_of, keep = [], [], []
        eps = None
        device = None
        for gi, group in enumerate(self.param_groups):
            if eps is None:
                eps = float(group["eps"])
            elif float(group["eps"]) != eps:
                raise NotImplementedError("FusedAdamW: all parameter groups must share eps")
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                if device is None:
                    device = p.device
                elif p.device != device:
                    raise RuntimeError("FusedAdamW: all parameters must live on one device")
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                st = self._init_state(p)
                if not st["step"].is_cuda:
                    raise RuntimeError("FusedAdamW: optimizer state 'step' is not on the device (load the checkpoint "
                                       "through FusedAdamW.load_state_dict)")
                entries.append(_lib.AdamWTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                                st["step"].data_ptr(), p.numel(), 0.0, 0.0))
                group_of.append(gi)
        if not entries:
            return loss
        with torch.cuda.device(device):
            table = self._sync_hyper(device)
            arr = (_lib.AdamWTensor * len(entries))(*entries)
            grp = (C.c_int * len(group_of))(*group_of)
            rc = ops._L().vae_adamw_step_dev(C.cast(arr, C.c_void_p), C.cast(grp, C.c_void_p), len(entries), table.data_ptr(),
                                             len(self.param_groups), eps, ops._stream())
        _lib.check(rc, "vae_adamw_step_dev")
        return loss


class FusedOneCycleLR:
    """``torch.optim.lr_scheduler.OneCycleLR`` (the scheduler of train.py:232-238, stepped per batch at
    train.py:659) advanced on the DEVICE: ``step()`` is one tiny capturable launch that bumps a device
    counter and rewrites lr / beta1 in the ``FusedAdamW`` hyper-parameter table.

    Same constructor arguments and arithmetic as the torch class (two phases, cos or linear annealing,
    momentum cycled inversely through ``betas[0]``); ``three_phase`` is not supported."""

    def __init__(self, optimizer: FusedAdamW, max_lr, total_steps=None, epochs=None, steps_per_epoch=None, pct_start=0.3,
                 anneal_strategy="cos", cycle_momentum=True, base_momentum=0.85, max_momentum=0.95, div_factor=25.0,
                 final_div_factor=1e4, three_phase=False, last_epoch=-1):
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("FusedOneCycleLR drives a torch_vae_b200.FusedAdamW")
        if three_phase:
            raise NotImplementedError("FusedOneCycleLR: three_phase is not supported")
        if last_epoch != -1:
            raise NotImplementedError("FusedOneCycleLR: resume through load_state_dict")
        if total_steps is None:
            if epochs is None or steps_per_epoch is None:
                raise ValueError("You must define either total_steps OR (epochs AND steps_per_epoch)")
            total_steps = int(epochs) * int(steps_per_epoch)
        if total_steps <= 1 or not 0.0 < pct_start < 1.0 or pct_start * total_steps <= 1.0:
            raise ValueError(f"Expected total_steps > 1 and 1/total_steps < pct_start < 1, got {total_steps}, {pct_start}")
        if anneal_strategy not in ("cos", "linear"):
            raise ValueError(f"anneal_strategy must be one of 'cos' or 'linear', instead got {anneal_strategy}")
        self.optimizer = optimizer
        self.total_steps, self.pct_start = int(total_steps), float(pct_start)
        self.anneal, self.cycle_momentum = (0 if anneal_strategy == "cos" else 1), bool(cycle_momentum)
        n = len(optimizer.param_groups)
        if n > 8:
            raise NotImplementedError("FusedOneCycleLR: at most 8 parameter groups")

        def per_group(v):
            v = list(v) if isinstance(v, (list, tuple)) else [v] * n
            if len(v) != n:
                raise ValueError(f"Expected {n} values, got {len(v)}")
            return [float(x) for x in v]

        max_lrs, base_ms, max_ms = per_group(max_lr), per_group(base_momentum), per_group(max_momentum)
        self._rows = []
        for g, group in enumerate(optimizer.param_groups):
            initial = max_lrs[g] / div_factor
            group["initial_lr"], group["max_lr"], group["min_lr"] = initial, max_lrs[g], initial / final_div_factor
            group["lr"] = initial
            if self.cycle_momentum:
                group["max_momentum"], group["base_momentum"] = max_ms[g], base_ms[g]
                group["betas"] = (max_ms[g], *group["betas"][1:])
            self._rows += [initial, max_lrs[g], initial / final_div_factor, base_ms[g], max_ms[g]]
        self._rows_c = (C.c_double * len(self._rows))(*self._rows)
        self._counter = None
        self.last_epoch = 0  # host mirror, exact only while step() is called eagerly (not under graph replay)
        optimizer._hyper_host = None  # the initial lr / beta1 set above reach the device at the next sync

    def _ensure(self):
        opt = self.optimizer
        dev = next(p.device for g in opt.param_groups for p in g["params"])
        if self._counter is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedOneCycleLR: call step() once eagerly before capturing")
            self._counter = torch.zeros(1, dtype=torch.int64, device=dev)
        opt._device_schedule = False
        table = opt._sync_hyper(dev)   # uploads the step-0 values once
        opt._device_schedule = True
        return dev, table

    def step(self):
        dev, table = self._ensure()
        with torch.cuda.device(dev):
            rc = ops._L().vae_onecycle_step(table.data_ptr(), self._counter.data_ptr(), C.cast(self._rows_c, C.c_void_p),
                                            len(self.optimizer.param_groups), self.total_steps, self.pct_start, self.anneal,
                                            1 if self.cycle_momentum else 0, ops._stream())
        _lib.check(rc, "vae_onecycle_step")
        self.last_epoch += 1

    def get_last_lr(self):
        """Reads the device table (one small D2H copy)."""
        _, table = self._ensure()
        return [float(v) for v in table[:, 0].cpu()]

    def device_step_count(self) -> int:
        return 0 if self._counter is None else int(self._counter.item())

    def state_dict(self):
        return {"total_steps": self.total_steps, "pct_start": self.pct_start, "anneal": self.anneal,
                "cycle_momentum": self.cycle_momentum, "last_epoch": self.device_step_count(), "rows": list(self._rows)}

    def load_state_dict(self, sd):
        self.total_steps, self.pct_start = int(sd["total_steps"]), float(sd["pct_start"])
        self.anneal, self.cycle_momentum = int(sd["anneal"]), bool(sd["cycle_momentum"])
        self._rows = [float(x) for x in sd["rows"]]
        self._rows_c = (C.c_double * len(self._rows))(*self._rows)
        dev, table = self._ensure()
        k = int(sd["last_epoch"])
        self._counter.fill_(max(k - 1, 0) if k > 0 else 0)
        self.last_epoch = k
        if k > 0:  # replay the last step so that the table holds the lr of step k
            self.step()
            self.last_epoch = k


def onecycle_reference_value(step: int, total_steps: int, initial, peak, final, pct_start=0.3, anneal="cos"):
    """Closed form of one OneCycleLR quantity at ``last_epoch == step`` (used by tests and docs)."""
    end0, end1 = pct_start * total_steps - 1.0, total_steps - 1.0

    def ann(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1.0) if anneal == "cos" else (b - a) * pct + a

    if step <= end0:
        return ann(initial, peak, step / end0)
    return ann(peak, final, (step - end0) / (end1 - end0))
