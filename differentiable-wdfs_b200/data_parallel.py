"""Multi-GPU data parallelism for WDF training: shard the independent sequences, all-reduce once.

Sequences never interact in the forward pass (every window restarts from zero state,
clipper_pot.py:110-124), so the batch axis shards trivially: rank r owns rows
``[r*B/G, (r+1)*B/G)`` and keeps its x / target / y shards in its own HBM. The only exchange is one
all-reduce (sum) per step of the DWDF_OUT_LEN raw gradient/loss sums — 192 bytes, pure latency —
issued on the compute stream right after the adjoint's fixed-order reduction; every rank then runs
the identical finalize (chain rule, loss) and optimizer update, so parameters stay bit-identical
across ranks without a broadcast. One process per GPU (torchrun).

Two ways to run the exchange:
* ``PeerComm`` — the engine's own: the step's reduction kernel exchanges the sums itself over NVLink peer memory
  (``dwdf_train_step_dp`` / ``dwdf_allreduce_sum``, include/dwdf.h: mailboxes mapped with CUDA IPC). torch.distributed
  only carries the 64-byte IPC handles once, at start-up (any backend: nccl, or gloo in the tests).
* ``DataParallelTrainer`` with ``torch.distributed.all_reduce`` between ``backward(raw=True)`` and ``finalize`` — the
  library-collective version, kept as the reference to compare against (and what the gloo CPU tests of the host logic run).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib as L


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


class PeerComm:
    """This rank's end of the peer-memory exchange (``dwdf_comm``): allocates the mailbox on ``device``, gathers the
    CUDA IPC handles of all ranks through ``torch.distributed`` (``group``; any backend) and maps the peers' mailboxes.
    ``world_size == 1`` (or no process group) gives a trivial communicator. Raises ``DwdfError`` if CUDA IPC / peer
    access is unavailable — callers may then fall back to ``torch.distributed.all_reduce``."""

    def __init__(self, device, group=None, timeout_s: float = 20.0):
        self.lib = L.lib()
        self.device = torch.device(device)
        active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if active else 0
        self.world_size = dist.get_world_size(group) if active else 1
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.dwdf_comm_create(self.rank, self.world_size, C.byref(self.handle)))
            L.check(self.lib.dwdf_comm_set_timeout(self.handle, float(timeout_s)))
            if self.world_size > 1:
                n = int(self.lib.dwdf_comm_handle_bytes())
                mine = C.create_string_buffer(n)
                L.check(self.lib.dwdf_comm_get_handle(self.handle, mine))
                gathered = [None] * self.world_size
                dist.all_gather_object(gathered, bytes(mine.raw), group=group)
                blob = C.create_string_buffer(b"".join(gathered), n * self.world_size)
                L.check(self.lib.dwdf_comm_connect(self.handle, blob))
                dist.barrier(group=group)  # every mailbox is zeroed and mapped before anyone's first exchange

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks of a float64 device vector (at most 2047 elements), on the current stream."""
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device == self.device):
            raise ValueError("all_reduce_ takes a contiguous float64 tensor on the communicator's device")
        with torch.cuda.device(self.device):
            L.check(self.lib.dwdf_allreduce_sum(self.handle, C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return t

    def close(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.dwdf_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
