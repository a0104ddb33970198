"""The train step of the reference loop (train.py:620-659) as one replayable object, and a piano-roll
dataset staged ONCE in HBM.

``TrainStep`` runs, in the reference's order, forward (train.py:635), ``zero_grad`` (:642), loss (:644),
backward (:650) [+ the bucketed NCCL gradient all-reduce of north_star item 4], ``optimizer.step`` (:656) and
``scheduler.step`` (:659); after three eager steps it captures the sequence in a CUDA graph and replays it, so
a step costs one launch on the host.  The three loss scalars stay on the device until somebody reads them
(the reference's three ``.item()`` syncs per step, :672-674, become one copy per log interval).

``DeviceRollDataset`` replaces ``ImageFolder`` + DataLoader workers + per-sample torchvision transforms
(datasets.py:97-121, data_transformations.py:48-72) for piano-rolls: the uint8 rolls live in device memory,
an epoch is a device permutation, a batch is one gather + one ``RollTransform`` kernel; nothing crosses PCIe
per step (train.py:625).
"""

from __future__ import annotations

from typing import Dict, Iterator, Optional

import torch

from . import ops
from .data_transformations import RollTransform
from .optim import FusedAdamW, FusedOneCycleLR
from .parallel import GradientBuckets


class TrainStep:
    def __init__(self, model, optimizer, scheduler=None, buckets: Optional[GradientBuckets] = None, use_graph: bool = True,
                 warm_steps: int = 3, zero_arena: bool = True):
        self.model, self.opt, self.sched, self.buckets = model, optimizer, scheduler, buckets
        # statistics / split-K scratch of the step from one arena cleared once per step (ops.ZeroArena)
        self._arena_dev = next(model.parameters()).device if zero_arena else None
        if self._arena_dev is not None:
            ops.enable_zero_arena(self._arena_dev)
        self.use_graph = use_graph
        self.warm_steps = max(1, int(warm_steps))   # >= 1: the optimizer's device tables must exist before capture
        self._graph = None
        self._static_x = None
        self._static_losses = None
        self._eager_steps = 0
        # warm-up steps and the capture run on a private (non-default) stream: autograd's AccumulateGrad nodes and
        # the NCCL / weight-packing side streams must not carry dependencies on default-stream work into the capture
        self._side = torch.cuda.Stream() if use_graph else None

    # train.py:635-659
    def _step(self, x) -> Dict[str, torch.Tensor]:
        if self._arena_dev is not None:
            ops.zero_arena_reset(self._arena_dev)
        out = self.model.forward(x)
        self.opt.zero_grad(set_to_none=True)
        losses = self.model.loss(out)
        if self.buckets is not None:
            self.buckets.begin()
        losses["loss"].backward()
        if self.buckets is not None:
            self.buckets.finish()
        self.opt.step()
        if self.sched is not None:
            self.sched.step()
        # detached: a caller holding the loss must not keep the autograd graph -- and with it the parameters'
        # AccumulateGrad nodes, which remember the stream they were created on -- alive into the next step (a stale
        # node from a warm-up stream makes the capture fail with cudaErrorStreamCaptureIsolation)
        return {k: v.detach() for k, v in losses.items()}

    def _capture(self, x) -> None:
        self._static_x = x.clone()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            self._static_losses = self._step(self._static_x)
        torch.cuda.synchronize()
        self._graph = graph

    def _eager_on_side_stream(self, x):
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            losses = self._step(x)
        cur.wait_stream(self._side)
        return losses

    def __call__(self, x) -> Dict[str, torch.Tensor]:
        """One optimisation step on the batch ``x`` ([B, 1, H, W] float32, cuda).  The returned loss tensors
        are device-resident and are overwritten by the next call when the step runs as a graph."""
        if not self.use_graph:
            return self._step(x)
        if self._graph is None:
            if self._eager_steps < self.warm_steps:
                self._eager_steps += 1
                return self._eager_on_side_stream(x)
            if not isinstance(self.opt, FusedAdamW) or (self.sched is not None and not isinstance(self.sched, FusedOneCycleLR)):
                raise RuntimeError("TrainStep: graph replay needs FusedAdamW (+ FusedOneCycleLR): a host-side optimizer or "
                                   "scheduler would be frozen at its capture-time values; pass use_graph=False")
            self._capture(x)   # capture does not execute: fall through to the first replay
        if x.shape != self._static_x.shape:
            raise RuntimeError(f"TrainStep: batch shape changed from {tuple(self._static_x.shape)} to {tuple(x.shape)}; "
                               "drop the ragged last batch (drop_last) or pass use_graph=False")
        self._static_x.copy_(x, non_blocking=True)
        self._graph.replay()
        return self._static_losses


class DeviceRollDataset:
    """uint8 piano-rolls ``[N, Hs, Ws]`` (pitch x time, values 0 / 255 or grey levels) resident in HBM."""

    def __init__(self, rolls: torch.Tensor, transform: RollTransform, batch_size: int, rank: int = 0, world: int = 1,
                 shuffle: bool = True, drop_last: bool = True, seed: int = 0):
        if rolls.dtype != torch.uint8 or rolls.dim() != 3:
            raise ValueError("rolls must be a uint8 tensor [N, pitches, steps]")
        if not rolls.is_cuda:
            raise RuntimeError("DeviceRollDataset keeps the rolls on the GPU: move them with .cuda() once")
        self.rolls, self.transform = rolls.contiguous(), transform
        self.batch_size, self.rank, self.world = int(batch_size), int(rank), int(world)
        self.shuffle, self.drop_last, self.seed = shuffle, drop_last, int(seed)
        self._gen = torch.Generator(device=rolls.device)
        per_rank = rolls.shape[0] // self.world
        self.steps_per_epoch = per_rank // self.batch_size if drop_last else -(-per_rank // self.batch_size)
        if self.steps_per_epoch == 0:
            raise ValueError(f"{rolls.shape[0]} rolls cannot fill one batch of {batch_size} on {world} rank(s)")

    def __len__(self) -> int:
        return self.steps_per_epoch

    def epoch(self, epoch: int) -> Iterator[torch.Tensor]:
        """Batches of one epoch: every rank draws the SAME permutation (seed + epoch, the DistributedSampler
        convention of train.py:356-363) and takes its own strided share."""
        n = self.rolls.shape[0]
        if self.shuffle:
            self._gen.manual_seed(self.seed + epoch)
            perm = torch.randperm(n, device=self.rolls.device, generator=self._gen)
        else:
            perm = torch.arange(n, device=self.rolls.device)
        mine = perm[self.rank::self.world]
        for i in range(self.steps_per_epoch):
            idx = mine[i * self.batch_size:(i + 1) * self.batch_size]
            yield self.transform(self.rolls.index_select(0, idx))   # one gather + one transform kernel


def synthetic_rolls(n: int, pitches: int, steps: int, device, density: float = 0.10, seed: int = 1234) -> torch.Tensor:
    """Random {0, 255} piano-rolls with ~``density`` of the cells on, notes held for a few steps."""
    g = torch.Generator(device=device).manual_seed(seed)
    onsets = torch.rand((n, pitches, steps), device=device, generator=g) < density / 3.0
    held = onsets.clone()
    held[:, :, 1:] |= onsets[:, :, :-1]
    held[:, :, 2:] |= onsets[:, :, :-2]
    return held.to(torch.uint8) * 255
