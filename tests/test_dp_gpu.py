"""Multi-GPU data parallelism on real devices: a sharded training step equals the unsharded one.

Two processes (one per rank, as torchrun would start them), each on its own GPU when the box has two — on a one-GPU
box both ranks share the device (two contexts, time-sliced; CUDA IPC between processes works on one device too), which
exercises the same code: mailbox creation, IPC handle exchange, the reduction kernel's exchange over peer memory, the
identical update on every rank. torch.distributed (gloo) only carries the IPC handles and the final comparison.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dwdf, device, mode="approx"):
    Vs = dwdf.ResistiveVoltageSource(47000.0, True)
    Cc = dwdf.Capacitor(2.2e-9, 48000.0, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
    circ = dwdf.compile_circuit(dp, probe=Cc, ordering="python", device=device)
    opt = dwdf.Adam(circ, lr={s: 1e-3 * float(circ.params[s]) for s in range(circ.n_params)}, beta_1=0.5)
    return circ, opt


def _worker(rank, world, port, B, T, steps, loss, q):
    import importlib

    import torch.distributed as dist

    from conftest import make_inputs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dwdf = importlib.import_module("differentiable-wdfs_b200")
    device = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(device)
    x = make_inputs(B, T, seed=77)  # generated globally, then sliced: identical at every world size
    target = 0.8 * np.tanh(2.0 * x).astype(np.float32) * 0.3
    lo, hi = dwdf.shard_rows(B, world, rank)
    xs, ts = torch.from_numpy(x[lo:hi]).to(device), torch.from_numpy(target[lo:hi]).to(device)
    comm = dwdf.PeerComm(device)
    circ, opt = _build(dwdf, device)
    hist = []
    for _ in range(steps):
        res = circ.train_step(xs, ts, opt, loss=loss, skip=5, comm=comm)
        hist.append((res["out"].cpu().numpy().copy(), circ.params.cpu().numpy().copy()))
    # the generic exchange on a longer vector (what the neural root's weight gradients use)
    v = torch.arange(700, dtype=torch.float64, device=device) * (rank + 1)
    comm.all_reduce_(v)
    torch.cuda.synchronize()
    ref = None
    if rank == 0:  # the same steps on the whole batch, one process, no communicator
        c1, o1 = _build(dwdf, device)
        xa, ta = torch.from_numpy(x).to(device), torch.from_numpy(target).to(device)
        ref = []
        for _ in range(steps):
            r1 = c1.train_step(xa, ta, o1, loss=loss, skip=5)
            ref.append((r1["out"].cpu().numpy().copy(), c1.params.cpu().numpy().copy()))
    q.put((rank, hist, ref, v.cpu().numpy()))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,T,loss", [(8192, 512, "mse"), (333, 1024, "mse+esr"), (200, 512, "mse+esr_as_called")])
def test_two_rank_sharded_step_equals_unsharded(B, T, loss):
    import torch.multiprocessing as mp

    world, steps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, T, steps, loss, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, h0, ref, v0), (_, h1, _, v1) = out
    for s in range(steps):
        # both ranks hold the same bits: same sums added in the same (rank) order, same update
        np.testing.assert_array_equal(h0[s][0], h1[s][0])
        np.testing.assert_array_equal(h0[s][1], h1[s][1])
        # sharded == unsharded: gradients and loss to the summation order of the fp64 partials (and of the adjoint's time
        # chunks, whose count depends on the per-rank batch), parameters after Adam to fp32 round-off
        np.testing.assert_allclose(h0[s][0][:4], ref[s][0][:4], rtol=3e-5)
        np.testing.assert_allclose(h0[s][0][16:19], ref[s][0][16:19], rtol=1e-6)
        np.testing.assert_allclose(h0[s][1], ref[s][1], rtol=2e-6)
    want = np.arange(700, dtype=np.float64) * 3
    np.testing.assert_array_equal(v0, want)
    np.testing.assert_array_equal(v1, want)


def _worker_nn(rank, world, port, B, T, steps, loss, q):
    import importlib
    import os as _os

    import torch.distributed as dist

    from conftest import GOLDEN, make_inputs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dwdf = importlib.import_module("differentiable-wdfs_b200")
    device = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(device)
    nnv = np.load(_os.path.join(GOLDEN, "nn_vectors.npz"))
    mj = dwdf.model_io.json_from_weights(nnv["2x8_weights"], [int(v) for v in nnv["2x8_sizes"]])

    def build():
        Vs = dwdf.ResistiveVoltageSource(47000.0)
        Cc = dwdf.Capacitor(2.2e-9, 48000.0)
        circ = dwdf.compile_circuit(dwdf.DenseRootModel(mj), tree=dwdf.Parallel(Vs, Cc), probe=Cc, ordering="python", device=device)
        return circ, dwdf.AdamWeights(circ, lr=1e-3, beta_1=0.5)

    x = make_inputs(B, T, seed=78)
    target = (0.5 * np.tanh(2.0 * x)).astype(np.float32)
    lo, hi = dwdf.shard_rows(B, world, rank)
    xs, ts = torch.from_numpy(x[lo:hi]).to(device), torch.from_numpy(target[lo:hi]).to(device)
    comm = dwdf.PeerComm(device)
    circ, opt = build()
    hist = []
    for _ in range(steps):
        res = circ.train_step(xs, ts, opt, loss=loss, skip=20, comm=comm)
        hist.append((res["out"][16:19].cpu().numpy().copy(), res["grads"].cpu().numpy().copy(), circ.weights.cpu().numpy().copy()))
    torch.cuda.synchronize()
    ref = None
    if rank == 0:
        c1, o1 = build()
        xa, ta = torch.from_numpy(x).to(device), torch.from_numpy(target).to(device)
        ref = []
        for _ in range(steps):
            r1 = c1.train_step(xa, ta, o1, loss=loss, skip=20)
            ref.append((r1["out"][16:19].cpu().numpy().copy(), r1["grads"].cpu().numpy().copy(), c1.weights.cpu().numpy().copy()))
    q.put((rank, hist, ref))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("loss", ["mse+esr", "mse+esr_as_called"])
def test_two_rank_neural_root_step_equals_unsharded(loss):
    """The root the reference actually trains (clipper_pot.py:246-269: the network's weights), sharded over two ranks: raw
    weight-gradient sums and loss sums exchanged over peer memory inside dwdf_train_step_neural."""
    import torch.multiprocessing as mp

    world, steps, B, T = 2, 3, 50, 400
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_nn, args=(r, world, port, B, T, steps, loss, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, h0, ref), (_, h1, _) = out
    for s in range(steps):
        for k in range(3):
            np.testing.assert_array_equal(h0[s][k], h1[s][k])  # identical bits on both ranks
        np.testing.assert_allclose(h0[s][0], ref[s][0], rtol=1e-6)
        scale = np.max(np.abs(ref[s][1]))
        assert np.max(np.abs(h0[s][1] - ref[s][1])) < 3e-5 * scale
        np.testing.assert_allclose(h0[s][2], ref[s][2], rtol=2e-5, atol=1e-7)
