"""Data-parallel training of the VAE step: bucketed gradient all-reduce over
NCCL / NVLink 5, overlapped with backward  (north_star item 4).

The reference has no data parallelism at all (``WORLD_SIZE`` only scales batch
size and LR, train.py:165-166,201), so the semantics are defined
here, DDP-style:

  * one process per GPU; every rank holds a full replica and takes its own
    ``batch_size_per_gpu`` samples (global batch = per-GPU batch x world);
  * BatchNorm uses PER-REPLICA batch statistics (no SyncBN); running buffers
    stay rank-local and can be synchronised with ``broadcast_buffers()``;
  * parameter gradients are averaged over ranks: N-rank result == mean over ranks
    of the single-rank gradients on each shard;
  * reparameterisation noise is seeded ``seed + rank`` (train.py:356-363).

Mechanics: parameters are packed, in REVERSE execution order (final_layer ->
decoder -> decoder_input -> fc_var/fc_mu -> encoder), into flat fp32 buckets.
The backward schedules in ``engine.py`` hand each layer's gradients to
``push()`` the moment they are produced; when the last gradient of a bucket
arrives its all-reduce is enqueued on a side stream (ordered after the producing
kernels by an event), so NVLink traffic runs under the remaining backward
kernels.  ``finish()`` joins the side stream and points every ``p.grad`` at its
reduced bucket view.  The path shards with no other data-path collective.

Ownership: while a sink is attached the backward schedules write each gradient
straight into its bucket slice (no gather copy) and return ``None`` to autograd
for the parameters, so ``AccumulateGrad`` never touches a bucket that is being
reduced.  Consequently gradient ACCUMULATION over several backward passes is not
supported with a sink: every backward overwrites the buckets (``begin()`` may be
preceded by ``zero_grad`` with either ``set_to_none`` setting).
"""

from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def execution_order_reversed(model) -> List[torch.nn.Parameter]:
    """Parameters in the order their gradients are produced by backward."""
    order: List[torch.nn.Parameter] = []
    last = model.final_layer
    order += [last[3].weight, last[3].bias]
    for blk in [model.final_layer] + list(model.decoder)[::-1]:
        order += [blk[0].weight, blk[0].bias, blk[1].weight, blk[1].bias]
    order += [model.decoder_input.weight, model.decoder_input.bias]
    order += [model.fc_mu.weight, model.fc_mu.bias, model.fc_var.weight, model.fc_var.bias]
    for blk in list(model.encoder)[::-1]:
        order += [blk[0].weight, blk[0].bias, blk[1].weight, blk[1].bias]
    return order


class GradientBuckets:
    def __init__(self, model, process_group=None, bucket_cap_mb: float = 25.0,
                 first_bucket_cap_mb: float = 1.0, average: bool = True):
        # Without an initialised process group the sink is LOCAL (one process, one GPU): the backward schedules still
        # write every gradient straight into its bucket slice and autograd gets None, so the step runs the same
        # gradient path as the N-GPU job, minus the collective (no per-parameter AccumulateGrad launches).
        self.local = not dist.is_initialized()
        self.model = model
        self.group = process_group
        self.world = 1 if self.local else dist.get_world_size(process_group)
        self.average = average
        params = [p for p in execution_order_reversed(model) if p.requires_grad]
        missing = set(id(p) for p in model.parameters() if p.requires_grad) - set(id(p) for p in params)
        if missing:
            raise RuntimeError("model has trainable parameters outside the known VanillaVAE layout")
        self.device = params[0].device
        self.is_cuda = self.device.type == "cuda"
        # greedy packing in gradient-production order; a small first bucket starts the
        # pipeline early (the last layers' gradients are tiny)
        caps = [int(first_bucket_cap_mb * 2**20 / 4), int(bucket_cap_mb * 2**20 / 4)]
        groups: List[List[torch.nn.Parameter]] = [[]]
        size = 0
        for p in params:
            cap = caps[0] if len(groups) == 1 else caps[1]
            if groups[-1] and size + p.numel() > cap:
                groups.append([])
                size = 0
            groups[-1].append(p)
            size += p.numel()
        self.buckets: List[torch.Tensor] = []
        self.slot: Dict[int, tuple] = {}
        self.pending: List[int] = []
        for bi, grp in enumerate(groups):
            n = sum(p.numel() for p in grp)
            flat = torch.zeros(n, dtype=torch.float32, device=self.device)
            off = 0
            for p in grp:
                self.slot[p.data_ptr()] = (bi, off, p)
                off += p.numel()
            self.buckets.append(flat)
            self.pending.append(len(grp))
        self._count = list(self.pending)
        self._staged = [([], []) for _ in self.buckets]
        self._nccl_avg = (not self.local) and self.is_cuda and dist.get_backend(process_group) == "nccl"
        self._works: List[Optional[object]] = [None] * len(self.buckets)
        self.comm_stream = torch.cuda.Stream(device=self.device) if (self.is_cuda and not self.local) else None
        self.bytes_per_step = 4 * sum(b.numel() for b in self.buckets)
        model._grad_sink = self

    # ------------------------------------------------------------------
    def view(self, p) -> torch.Tensor:
        bi, off, _ = self.slot[p.data_ptr()]
        return self.buckets[bi][off : off + p.numel()].view_as(p)

    def begin(self) -> None:
        """Call before backward of every step."""
        self._count = list(self.pending)
        self._works = [None] * len(self.buckets)
        self._staged = [([], []) for _ in self.buckets]

    def push(self, p, grad: torch.Tensor) -> None:
        """Gradient of ``p`` is ready (called from engine.py backward schedules)."""
        ent = self.slot.get(p.data_ptr())
        if ent is None:
            return
        bi = ent[0]
        view = self.view(p)
        if grad.data_ptr() != view.data_ptr():  # kernels normally write straight into the bucket slice (engine._dst)
            dst, src = self._staged[bi]
            dst.append(view)
            src.append(grad.reshape(p.shape))
        self._count[bi] -= 1
        if self._count[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        flat = self.buckets[bi]
        dst, src = self._staged[bi]
        if dst:  # gradients that were produced elsewhere: one multi-tensor copy, not one copy each
            torch._foreach_copy_(dst, src)
        self._staged[bi] = ([], [])
        if self.local:
            return
        op = dist.ReduceOp.SUM
        if self.is_cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
            self.comm_stream.wait_event(ready)
            with torch.cuda.stream(self.comm_stream):
                if self.average and self.world > 1:
                    if self._nccl_avg:
                        op = dist.ReduceOp.AVG  # averaged inside the collective: no pre-scale launch
                    else:
                        flat.mul_(1.0 / self.world)  # pre-scale: sum of pre-scaled == mean
                self._works[bi] = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        else:
            if self.average and self.world > 1:
                flat.mul_(1.0 / self.world)
            self._works[bi] = dist.all_reduce(flat, op=op, group=self.group, async_op=True)

    def finish(self) -> None:
        """Join the reductions and expose the averaged gradients as ``p.grad``."""
        for bi, c in enumerate(self._count):
            if c != 0:
                raise RuntimeError(f"gradient bucket {bi} incomplete: {c} gradients were never produced")
        for w in self._works:
            if w is not None:
                w.wait()
        if self.is_cuda and self.comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        for _, (bi, off, p) in self.slot.items():
            p.grad = self.view(p)

    def broadcast_buffers(self, src: int = 0) -> None:
        """DDP's ``broadcast_buffers`` equivalent: running stats follow rank ``src``."""
        if self.local:
            return
        for b in self.model.buffers():
            dist.broadcast(b, src=src, group=self.group)

    def broadcast_parameters(self, src: int = 0) -> None:
        if self.local:
            return
        for p in self.model.parameters():
            dist.broadcast(p.data, src=src, group=self.group)
