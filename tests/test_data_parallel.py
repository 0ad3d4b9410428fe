"""Host logic of the multi-GPU path on CPU: world_size 2 over gloo.

The GPU kernels are not involved here (no GPU in this container): each rank computes the RAW sums of
its shard with the CPU oracle, the product's DataParallelTrainer owns the sharding and the single
all-reduce, and the result must equal the oracle on the unsharded batch.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, T, q):
    import importlib

    from conftest import make_inputs
    from oracle.cpu import ClipperParams, Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = importlib.import_module("differentiable-wdfs_b200.data_parallel")
    orc, p = Oracle(), ClipperParams()
    x = make_inputs(B, T, seed=5)  # generated globally, then sliced: identical at every world size
    target = orc.clipper_forward(x, ClipperParams(R=p.R * 1.1, Is=p.Is * 2), exact=True)
    lo, hi = dp.shard_rows(B, world, rank)
    calls = {"apply": 0}

    def local_raw(xs, ts):
        g = orc.clipper_grad(xs, ts, p, exact=False, mode="target", loss="mse", dtype=np.float64)
        n = xs.size
        raw = np.zeros(24)
        raw[:4] = g["grads"] * n / 2.0  # undo the 2/N of the MSE: plain sums over this shard
        raw[4] = g["loss"] * n  # sum of squared errors
        raw[23] = n
        return torch.from_numpy(raw)

    def finalize(raw):
        n = float(raw[23])
        return {"grads": raw[:4] * 2.0 / n, "loss": raw[4] / n}

    tr = dp.DataParallelTrainer(local_raw, finalize, apply=lambda: calls.__setitem__("apply", calls["apply"] + 1))
    assert tr.world_size == world
    res = tr.step(x[lo:hi], target[lo:hi])
    full = orc.clipper_grad(x, target, p, exact=False, mode="target", loss="mse", dtype=np.float64)
    q.put((rank, lo, hi, res["grads"].numpy().tolist(), float(res["loss"]), full["grads"].tolist(), float(full["loss"]), calls["apply"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 7])
def test_two_rank_step_equals_unsharded(B):
    world, T = 2, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    rows = [(o[1], o[2]) for o in out]
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == B  # contiguous, complete, disjoint
    for o in out:
        np.testing.assert_allclose(o[3], o[5], rtol=1e-10)  # sharded == unsharded gradient
        assert abs(o[4] / o[6] - 1) < 1e-12
        assert o[7] == 1
    assert out[0][3] == out[1][3]  # bit-identical on both ranks: parameters stay in lock-step without a broadcast


def test_shard_rows_partition():
    import importlib

    dp = importlib.import_module("differentiable-wdfs_b200.data_parallel")
    for B in (0, 1, 7, 64, 65536):
        for w in (1, 2, 3, 8):
            cuts = [dp.shard_rows(B, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    with pytest.raises(ValueError):
        dp.shard_rows(8, 2, 2)
