"""Multi-GPU data parallelism for WDF training: shard the independent sequences, all-reduce once.

Sequences never interact in the forward pass (every window restarts from zero state,
clipper_pot.py:110-124), so the batch axis shards trivially: rank r owns rows
``[r*B/G, (r+1)*B/G)`` and keeps its x / target / y shards in its own HBM. The only exchange is one
all-reduce (sum) per step of the DWDF_OUT_LEN raw gradient/loss sums — 192 bytes, pure latency —
issued on the compute stream right after the adjoint's fixed-order reduction; every rank then runs
the identical finalize (chain rule, loss) and optimizer update, so parameters stay bit-identical
across ranks without a broadcast. One process per GPU (torchrun); NCCL over NVLink for the real
thing, gloo in the CPU tests of the host logic.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(B: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of the batch axis: rows [start, stop) of rank ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return B * rank // world_size, B * (rank + 1) // world_size


class DataParallelTrainer:
    """Drives one training step per call on this rank's shard.

    ``local_raw(x, target) -> tensor[OUT_LEN] (float64)`` computes this rank's RAW sums (on the GPU:
    ``CompiledCircuit.forward`` + ``backward(target=..., raw=True)`` or ``train_pass(raw=True)``);
    ``finalize(raw) -> dict`` turns the all-reduced sums into gradients and loss
    (``CompiledCircuit.finalize``); ``apply()`` is the optimizer step. The trainer only owns the
    collective — it is the same code under NCCL and gloo.
    """

    def __init__(self, local_raw: Callable, finalize: Callable, apply: Optional[Callable] = None, group=None):
        self.local_raw = local_raw
        self.finalize = finalize
        self.apply = apply
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def step(self, x, target):
        raw = self.local_raw(x, target)
        if self.world_size > 1:
            dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=self.group)  # the single collective of the step
        res = self.finalize(raw)
        if self.apply is not None:
            self.apply()
        return res


def clipper_trainer(circuit, optimizer=None, loss="mse", skip=0, fused=False, group=None) -> DataParallelTrainer:
    """The standard wiring for a CompiledCircuit on this rank's GPU."""

    def local_raw(x, target):
        if fused:
            circuit.train_pass(x, target, skip=skip, raw=True)
        else:
            circuit.forward(x)
            circuit.backward(target=target, skip=skip, raw=True)
        return circuit.out

    def finalize(raw):
        return circuit.finalize(target=True, loss=loss)

    return DataParallelTrainer(local_raw, finalize, (lambda: optimizer.apply()) if optimizer is not None else None, group)
