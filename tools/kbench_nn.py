#!/usr/bin/env python
"""Kernel micro-benchmark of the neural-root clipper forward (development tool).
python tools/kbench_nn.py [--B 65536] [--T 4096] [--models 2x4,2x8,2x16,4x4,4x8] [--r]"""
import argparse, ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bench import synth_inputs, FS

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--T", type=int, default=4096)
ap.add_argument("--models", default="2x4,2x8,2x16,4x4,4x8")
ap.add_argument("--r", action="store_true")
ap.add_argument("--cpu-rows", type=int, default=64)
ap.add_argument("--opt", type=int, default=0, help="dwdf_set_option bits (8: no time-parallel kernels, 16: force them)")
a = ap.parse_args()
dwdf = importlib.import_module("differentiable-wdfs_b200")
nnv = np.load(os.path.join(ROOT, "tests", "golden", "nn_vectors.npz"))
dev = torch.device("cuda", 0)
if a.opt:
    dwdf.set_option(a.opt)
bench.T = a.T
x = synth_inputs(torch, a.B, 1, dev)
r = torch.full_like(x, 47000.0) if a.r else None
y = torch.empty_like(x)
for name in a.models.split(","):
    mj = dwdf.model_io.json_from_weights(nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]])
    Vs = dwdf.ResistiveVoltageSource(47000.0); Cc = dwdf.Capacitor(2.2e-9, FS); P1 = dwdf.Parallel(Vs, Cc)
    circ = dwdf.compile_circuit(dwdf.DenseRootModel(mj), tree=P1, probe=Cc, ordering="plugin", r_element=Vs if a.r else None, device=dev)
    for _ in range(2): circ.forward(x, r=r, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): circ.forward(x, r=r, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    redone = dwdf.time_parallel_redone()
    n_w = circ.weights.numel()
    line = f"neural root {name} ({n_w} weights) B={a.B} T={a.T} r={a.r}: forward {ms:.3f} ms [redone chunks so far {redone}]  {a.B*a.T/ms/1e6:.2f} Gsamples/s  ({2*n_w*a.B*a.T/ms/1e9:.1f} TFLOP/s in the network)"
    # CPU: oracle (numpy) is not a fair baseline; the reference's own RTNeural path lives in oracle/_ref (not on the GPU box unless built here)
    ref = os.path.join(ROOT, "oracle", "_ref", "libdwdf_ref_nn.so")
    print(line, flush=True)
    tgt = (0.9 * y).clone()
    circ.forward(x, r=r)
    for _ in range(1): circ.backward(target=tgt)
    torch.cuda.synchronize(); e0.record(); circ.backward(target=tgt); e1.record(); torch.cuda.synchronize()
    msb = e0.elapsed_time(e1)
    print(f"            adjoint {msb:.3f} ms  fwd+bwd {a.B*a.T/(ms+msb)/1e6:.2f} Gsamples/s", flush=True)
